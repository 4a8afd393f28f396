"""nerf_signature_b200 — B200 (sm_100a) native implementation of NeRF_Signature's ray-batch
render/train hot path behind the reference's own Python API.

    from nerf_signature_b200 import raymarching                     # drop-in for `import raymarching`
    from nerf_signature_b200.hash_encoding import HashEmbedder       # hash_encoding.py
    from nerf_signature_b200.hash_encoding_wtmk_bit import HashEmbedder as HashEmbedder_msg
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork   # watermark field + renderer
    from nerf_signature_b200.nerf.network_hash import NeRFNetwork as CleanNeRFNetwork

All compute goes through libnsig_b200.so (C ABI in include/nsig.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
__all__ = ["raymarching", "hash_encoding", "hash_encoding_wtmk_bit", "activation", "nerf"]


def __getattr__(name):  # lazy sub-module import keeps `import nerf_signature_b200` cheap
    if name in __all__:
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
