"""CPU tests (-m "not gpu"): the HOST side of the package's hash encoders (module construction, level resolutions, pointer /
resolution arrays, autograd glue, which tables receive a gradient) against the REFERENCE MODULES themselves, imported
unmodified (hash_encoding.py, hash_encoding_wtmk_bit.py), with the device faked: `_lib.call` dispatches to the C oracle,
whose entry points have the argument order of the nsig_* functions, and nsig_msg_table_sum is emulated from its contract.
The kernels' arithmetic is tested on the GPU (tests/test_hash_gpu.py); here the numbers are the oracle's and what is
checked is everything the Python layer adds around them."""
import ctypes
import os

import numpy as np
import pytest
import torch

from make_golden_hash import load_ref_module, make_tables
from nerf_signature_b200 import _lib
from nerf_signature_b200 import hash_encoding as he
from nerf_signature_b200 import hash_encoding_wtmk_bit as hm

REF = os.environ.get("NSIG_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "hash_encoding_wtmk_bit.py")),
                                reason="needs the reference sources (build container only)")
LOG2_T = 11


class _HostPtr:
    def __init__(self, tensor):
        self.tensor = tensor
        self._as_parameter_ = ctypes.c_void_p(tensor.data_ptr())


@pytest.fixture
def fake_device(monkeypatch, oracle_cpu):
    olib = oracle_cpu.lib()
    names = {"nsig_hash_encode_forward": "oracle_hash_encode_forward", "nsig_hash_encode_backward": "oracle_hash_encode_backward",
             "nsig_msg_encode_forward_perbit": "oracle_msg_encode_forward"}

    def fake_call(name, *args):
        if name == "nsig_msg_table_sum":         # S[lo:lo+n] = sum_i tables[2i + bit_i][lo:lo+n]; n == 0: the whole table
            tables, md, msg, log2_T, S, lo, n = args
            assert (lo, n) == (0, 0) and len(tables._keep) >= 2 * md
            bits = msg.tensor.tolist()
            acc = torch.zeros_like(S.tensor)
            for i in range(md):
                acc += tables._keep[2 * i + int(bits[i])].detach()
            S.tensor.copy_(acc)
            return
        sig = list(_lib._SIGNATURES[name][0])[:-1]
        assert len(args) == len(sig), (name, len(args), len(sig))
        fn = getattr(olib, names[name])
        fn.restype = None
        fn(*[a if t is ctypes.c_void_p else t(a) for a, t in zip(args, sig)])

    ptr = lambda t: None if t is None else _HostPtr(t)                # noqa: E731
    monkeypatch.setattr(_lib, "call", fake_call)
    monkeypatch.setattr(_lib, "ptr", ptr)
    monkeypatch.setattr(he, "_P", ptr)
    monkeypatch.setattr(hm, "_P", ptr)
    return oracle_cpu


def _points(n, seed):
    rs = np.random.RandomState(seed)
    x = rs.uniform(0, 1, size=(n, 3)).astype(np.float32)
    x[:4] = [[0, 0, 0], [1, 1, 1], [0.5, 0.25, 0.125], [1, 0, 0.5]]
    return torch.from_numpy(x)


def test_base_encoder_module_against_the_reference_module(fake_device):
    ref_he = load_ref_module("hash_encoding.py", "ref_hash_encoding")
    kw = dict(bounding_box=(0, 1), n_levels=16, n_features_per_level=2, log2_hashmap_size=LOG2_T, base_resolution=16,
              finest_resolution=2048)
    ours, ref = he.HashEmbedder(**kw), ref_he.HashEmbedder(**kw)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())          # checkpoints load both ways
    assert [tuple(v.shape) for v in ours.state_dict().values()] == [tuple(v.shape) for v in ref.state_dict().values()]
    assert ours.out_dim == ref.out_dim and torch.equal(ours.b, ref.b)
    assert ours.resolutions == [float(torch.floor(ref.base_resolution * ref.b ** i)) for i in range(16)]
    tabs = make_tables(5, 16, LOG2_T)
    with torch.no_grad():
        for eo, er, t in zip(ours.embeddings, ref.embeddings, tabs):
            eo.weight.copy_(torch.from_numpy(t) * 1e3)
            er.weight.copy_(torch.from_numpy(t) * 1e3)
    x = _points(257, 6)
    fo, fr = ours(x), ref(x)
    assert fo.shape == fr.shape == (257, 32) and torch.equal(fo, fr)                # features: bit-exact
    g = torch.randn(257, 32, generator=torch.Generator().manual_seed(1))
    fo.backward(g)
    fr.backward(g)
    for eo, er in zip(ours.embeddings, ref.embeddings):
        assert eo.weight.grad.shape == er.weight.grad.shape
        assert (eo.weight.grad - er.weight.grad).abs().max() <= 1e-6 * er.weight.grad.abs().max()
    # frozen tables (watermark training freezes the base encoder): no gradient buffers are produced
    for e in ours.embeddings:
        e.weight.requires_grad_(False)
        e.weight.grad = None
    assert not ours(x).requires_grad
    with pytest.raises(NotImplementedError):
        he.HashEmbedder(bounding_box=(-1, 1))                                        # every reference call site passes (0, 1)


@pytest.mark.parametrize("md", [4, 32])
def test_message_encoder_module_against_the_reference_module(fake_device, md):
    ref_hm = load_ref_module("hash_encoding_wtmk_bit.py", "ref_hash_encoding_wtmk_bit")
    kw = dict(bounding_box=(0, 1), n_levels=2 * md, n_features_per_level=2, log2_hashmap_size=LOG2_T, base_resolution=2048,
              finest_resolution=2048, message_dim=md)
    ours, ref = hm.HashEmbedder(**kw), ref_hm.HashEmbedder(**kw)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    assert ours.resolution == 2048.0 and ours.out_dim == ref.out_dim
    tabs = make_tables(8, 2 * md, LOG2_T)
    with torch.no_grad():
        for eo, er, t in zip(ours.embeddings, ref.embeddings, tabs):
            eo.weight.copy_(torch.from_numpy(t) * 1e3)
            er.weight.copy_(torch.from_numpy(t) * 1e3)
    x = _points(199, 9)
    message = torch.from_numpy(np.random.RandomState(3).randint(0, 2, size=md).astype(np.float32))
    fo, fr = ours(x, message), ref(x, message)
    assert fo.shape == fr.shape == (199, 2)
    assert (fo - fr).abs().max() <= 2e-6 * md ** 0.5 * fr.abs().max() + 1e-7         # pre-summed form: summation order
    # literal per-bit order: what is left is the order in which torch sums the stacked per-bit features (one ulp)
    assert (ours.forward_perbit(x, message) - fr.detach()).abs().max() <= 1e-6 * fr.abs().max()
    g = torch.randn(199, 2, generator=torch.Generator().manual_seed(2))
    fo.backward(g)
    fr.backward(g)
    selected = {2 * i + int(b) for i, b in enumerate(message.tolist())}
    first = None
    for k, (eo, er) in enumerate(zip(ours.embeddings, ref.embeddings)):
        if k in selected:                                                            # the same dL/dS for every selected table
            assert eo.weight.grad is not None and er.weight.grad is not None
            assert (eo.weight.grad - er.weight.grad).abs().max() <= 1e-6 * er.weight.grad.abs().max()
            first = eo.weight.grad if first is None else first
            assert torch.equal(eo.weight.grad, first)
        else:                                                                        # unselected tables: no gradient at all
            assert eo.weight.grad is None
            assert er.weight.grad is None or not er.weight.grad.any()
    with pytest.raises(ValueError):
        ours(x, message[:-1])                                                        # wrong number of bits
