"""CPU tests (-m "not gpu") of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU,
exports exactly what include/nsig.h declares, validates arguments before touching CUDA, and the Python
layer refuses to run without CUDA (no CPU fallback, no route through oracle/)."""
import ast
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nsig.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nsig_[A-Za-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from nerf_signature_b200 import _build, _lib
    _build.build_library()
    return _lib


def test_library_exports_every_declared_symbol(lib):
    handle = lib.load()
    decl = declared_symbols()
    assert len(decl) >= 21
    for name in decl:
        assert hasattr(handle, name), f"{name} declared in include/nsig.h but not exported"
    # and the binding table covers the header (nothing declared is unreachable from Python)
    assert sorted(lib.EXPORTED_SYMBOLS) == decl
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "nsig_" in l}
    assert exported == set(decl), exported ^ set(decl)  # C linkage, no stray mangled nsig entry points


_C_TYPES = {"uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "int": ctypes.c_int, "float": ctypes.c_float,
            "double": ctypes.c_double, "uint64_t": ctypes.c_uint64, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t,
            "nsig_stream_t": ctypes.c_void_p}


def declared_prototypes():
    """{name: (return type, [ctypes type of every parameter])} parsed from include/nsig.h."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for ret, name, params in re.findall(r"\b([A-Za-z_][A-Za-z0-9_ \*]*?)\s*\b(nsig_[A-Za-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        types = []
        for prm in params.split(","):
            prm = prm.strip()
            if prm in ("", "void"):
                continue
            if "*" in prm:
                types.append(ctypes.c_void_p)
            else:
                toks = prm.replace("const", "").split()
                assert len(toks) == 2, f"{name}: cannot parse parameter '{prm}'"
                types.append(_C_TYPES[toks[0]])
        protos[name] = (" ".join(ret.split()), types)
    return protos


def test_ctypes_signatures_match_the_header_parameter_by_parameter(lib):
    """The binding passes scalars by ctypes argtypes: a count or type drift between include/nsig.h (which the .cu files
    compile against through nsig_common.cuh) and _lib._SIGNATURES would silently shift every later argument."""
    protos = declared_prototypes()
    assert sorted(protos) == declared_symbols()
    handle = lib.load()
    for name, (ret, types) in protos.items():
        fn = getattr(handle, name)
        assert list(fn.argtypes or []) == types, f"{name}: argtypes differ from the header"
        want_ret = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "uint32_t": ctypes.c_uint32,
                    "const char*": ctypes.c_char_p}[ret.replace(" *", "*")]
        assert fn.restype is want_ret, f"{name}: restype {fn.restype} vs header '{ret}'"
        if name in lib._SIGNATURES:   # every status-returning entry point ends with the stream it runs on
            assert ret == "int" and types[-1] is ctypes.c_void_p


def test_library_is_sm100a_only_and_has_no_torch_dependency(lib):
    assert lib.version().startswith("nsig_b200") and "sm_100a" in lib.version()
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    dyn = subprocess.run(["readelf", "-d", lib.LIB_PATH], capture_output=True, text=True).stdout
    needed = "\n".join(l for l in dyn.splitlines() if "(NEEDED)" in l)   # only the library names (addresses may spell "c10")
    assert needed and "torch" not in needed and "c10" not in needed and "python" not in needed


def test_argument_validation_happens_before_any_cuda_call(lib):
    h = lib.load()
    V = ctypes.c_void_p
    # empty inputs are a no-op success
    assert h.nsig_near_far_from_aabb(None, None, None, 0, 0.2, None, None, None) == 0
    assert h.nsig_packbits(None, 0, 0.5, None, None) == 0
    assert h.nsig_hash_encode_forward(None, 0, None, None, 16, 19, None, None, None) == 0
    # null pointers / impossible shapes are NSIG_EINVAL (-1), not a crash
    assert h.nsig_near_far_from_aabb(None, None, None, 4, 0.2, None, None, None) == -1
    assert h.nsig_morton3D(None, 8, None, None) == -1
    assert h.nsig_composite_rays_train_forward(None, None, None, None, 8, 8, 1e-4, None, None, None, None) == -1
    dummy = (ctypes.c_float * 8)()
    p = ctypes.cast(dummy, V)
    tabs = (V * 16)(*[p] * 16)
    res = (ctypes.c_float * 16)(*[16.0] * 16)
    assert h.nsig_hash_encode_forward(p, 1, tabs, res, 17, 19, p, None, None) == -1   # > NSIG_MAX_LEVELS
    assert h.nsig_hash_encode_forward(p, 1, tabs, res, 16, 31, p, None, None) == -1   # log2_T out of range
    assert h.nsig_march_rays_train(p, p, p, 1.0, 0.0, 1024, 4, 0, 128, 64, p, p, p, p, p, p, p, None, p, None) == -1  # C=0
    assert h.nsig_march_rays_train_scratch_bytes(1000) == 8000
    # entry points added for the fused step: loss head, composite epilogue, decoder
    assert h.nsig_split_clamp_forward(None, 0, 0, None, None, None) == 0
    assert h.nsig_split_clamp_forward(p, 9, 6, p, p, None) == -1                       # more block floats than floats
    assert h.nsig_split_clamp_backward(None, None, None, 3, 6, p, None) == -1
    assert h.nsig_wtmk_loss_forward(p, p, 6, p, p, 2, 0.005, 1.0, 10.0, None, p, p, None) == -1   # no output
    assert h.nsig_wtmk_loss_backward(p, p, 6, 2, None, p, p, None) == -1                   # no incoming gradient
    assert h.nsig_composite_rays_train_blend_forward(p, p, p, p, 8, 8, 1e-4, 1.0, None, None, p, p, p, p, p, None) == -1
    assert h.nsig_composite_rays_train_blend_backward(None, None, p, p, p, p, p, p, 8, 8, 1e-4, 1.0, p, p, None) == -1
    assert h.nsig_decoder_workspace_bytes(32, 12, 12, 0) == 0 and h.nsig_decoder_workspace_bytes(32, 12, 12, 8) > 0
    assert h.nsig_decoder_forward(p, 2, 4, 4, 8, 9, 1, tabs, p, p, None, None) == -1       # num_bits*redundancy > 8
    assert h.nsig_decoder_weights_bytes(0) == 0 and h.nsig_decoder_weights_bytes(8) > 2 * 7 * 64 * 9 * 64 * 2
    assert h.nsig_decoder_prepare_weights(tabs, 8, 1, 1, None, None) == -1                  # no output buffer
    assert h.nsig_decoder_prepare_weights(tabs, 0, 1, 1, p, None) == -1                     # no conv block
    # the fused field kernels stage the fp16 MLP weights with 16-byte copies: misaligned weight pointers are rejected
    buf = (ctypes.c_float * 64)()
    base = ctypes.addressof(buf)
    aligned, odd = V(base + (-base) % 16), V(base + (-base) % 16 + 2)
    # look-ahead table sum: needs both messages and an output; ranges must be multiples of 4 floats inside the table
    assert h.nsig_msg_adam_lookahead_sum(p, 8, 4, p, None, aligned, p, p, None, None, 1e-2, 0.9, 0.99, 1e-15, 12, None, 0, 0, aligned, None) == -1
    assert h.nsig_msg_adam_lookahead_sum(p, 8, 4, p, p, aligned, p, p, None, None, 1e-2, 0.9, 0.99, 1e-15, 12, None, 0, 0, None, None) == -1
    assert h.nsig_msg_adam_lookahead_sum(p, 8, 4, p, p, aligned, p, p, None, None, 1e-2, 0.9, 0.99, 1e-15, 12, None, 2, 8, aligned, None) == -1
    assert h.nsig_msg_adam_lookahead_sum(p, 8, 4, p, p, aligned, p, p, None, None, 1e-2, 0.9, 0.99, 1e-15, 12, None, 0, 1 << 20, aligned, None) == -1
    assert h.nsig_msg_adam_step(p, 8, 4, p, aligned, p, p, None, None, 1e-2, 0.9, 0.99, 1e-15, 12, None, 0, 6, 0, None) == -1   # ragged range
    # update + next message's table sum in one pass: same contract as the look-ahead sum
    assert h.nsig_msg_adam_step_sum(p, 8, 4, p, None, aligned, p, p, None, None, 1e-2, 0.9, 0.99, 1e-15, 12, None, 0, 0, aligned, None) == -1
    assert h.nsig_msg_adam_step_sum(p, 8, 4, p, p, aligned, p, p, None, None, 1e-2, 0.9, 0.99, 1e-15, 12, None, 0, 0, odd, None) == -1
    assert h.nsig_msg_adam_step_sum(p, 8, 4, p, p, aligned, p, p, None, None, 1e-2, 0.9, 0.99, 1e-15, 12, None, 2, 8, aligned, None) == -1
    # deferred decoder tail: a process-wide switch (no CUDA call), restored at once
    assert h.nsig_decoder_defer_weight_grads(1) == 0 and h.nsig_decoder_defer_weight_grads(0) == 0
    assert h.nsig_decoder_gelu_probe(None, 0, None, None, None) == 0 and h.nsig_decoder_gelu_probe(p, 8, None, p, None) == -1
    # grid-limited march: the unlimited entry's validation
    assert h.nsig_march_rays_train_limited(p, p, p, 1.0, 0.0, 1024, 4, 0, 128, 64, p, p, p, p, p, p, p, None, p, 296, None) == -1  # C=0
    assert h.nsig_march_rays_train_limited(p, p, p, 1.0, 0.0, 1024, 0, 1, 128, 64, p, p, p, p, p, p, p, None, p, 296, None) == 0   # no rays
    assert h.nsig_field_backward(p, p, 4, 1.0, p, p, p, odd, aligned, 1.0, None, 2048.0, 19, p, None, None, None, None) == -1
    # round-2 entry points: tcgen05 backward, fused-slot probe, one-kernel GradScaler, flat Adam
    assert h.nsig_field_backward_tc(None, None, 0, 1.0, None, None, None, None, None, 1.0, None, 2048.0, 19, None, None) == 0
    assert h.nsig_field_backward_tc(p, p, 4, 1.0, aligned, p, p, aligned, aligned, 1.0, None, 2048.0, 19, None, None) == -1   # no G
    assert h.nsig_field_backward_tc(p, p, 4, 1.0, odd, p, p, aligned, aligned, 1.0, None, 2048.0, 19, p, None) == -1       # feat alignment
    assert h.nsig_field_backward_tc(p, p, 4, 1.0, aligned, p, p, aligned, aligned, 1.0, None, 0.0, 19, p, None) == -1      # msg resolution
    assert h.nsig_field_backward_masks(None, 0, 1.0, None, None, None, None, None, None, None, 1.0, None, 2048.0, 19, None, None) == 0
    assert h.nsig_field_backward_masks(p, 4, 1.0, aligned, p, p, p, p, odd, aligned, 1.0, None, 2048.0, 19, p, None) == -1   # weights
    assert h.nsig_field_backward_tc_masks(p, 4, 1.0, odd, p, p, p, p, aligned, aligned, 1.0, None, 2048.0, 19, p, None) == -1  # masks
    assert h.nsig_field_backward_tc_masks(p, 4, 1.0, aligned, p, p, p, p, aligned, aligned, 0.0, None, 2048.0, 19, p, None) == -1
    assert h.nsig_fused_hash_slots(None, 0, None, 16, 19, None, None, None) == 0
    assert h.nsig_fused_hash_slots(p, 1, res, 17, 19, p, None, None) == -1
    assert h.nsig_grad_check_update_scale(None, 8, p, p, 2.0, 0.5, 2000, p, p, None, p, None, None) == -1
    assert h.nsig_grad_check_update_scale(aligned, 8, p, p, 2.0, 0.5, 0, p, p, None, p, None, None) == -1                  # growth_interval
    assert h.nsig_flat_adam_step(None, None, None, None, 0, None, None, None, 1e-2, None, 0.9, 0.99, 1e-15, None) == 0
    assert h.nsig_flat_adam_step(p, p, p, p, 8, None, None, None, 1e-2, None, 0.9, 0.99, 1e-15, None) == -1                # no step counter


def test_python_layer_has_no_cpu_path(lib):
    from nerf_signature_b200.hash_encoding import HashEmbedder
    enc = HashEmbedder(bounding_box=(0, 1), n_levels=16, log2_hashmap_size=10, base_resolution=16, finest_resolution=2048)
    with pytest.raises(lib.NsigError):
        enc(torch.rand(4, 3))  # CPU tensor: refuse, never fall back
    with pytest.raises(lib.NsigError):
        lib.ptr(torch.zeros(3))


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from nerf_signature_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libnsig_b200.so"))
    with pytest.raises(_lib.NsigError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under nerf_signature_b200/ may import or reference it."""
    pkg = os.path.join(ROOT, "nerf_signature_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                for m in mods:
                    assert not (m == "oracle" or m.startswith("oracle.")), f"{f} imports {m}"
    for dirpath, _, files in os.walk(os.path.join(pkg, "csrc")):
        for f in files:
            assert "oracle" not in open(os.path.join(dirpath, f)).read(), f


def test_every_entry_point_validates_before_touching_cuda(lib):
    """Error behaviour of the boundary, for ALL status-returning entry points at once: NULL pointers with non-zero sizes
    are NSIG_EINVAL (-1); all-zero sizes are a no-op (0) or, where even an empty call needs a valid configuration,
    NSIG_EINVAL.  This container has no GPU: a positive status (a cudaError_t such as cudaErrorNoDevice) would mean the
    entry point reached the CUDA runtime before checking its arguments - and none may crash."""
    h = lib.load()
    for size in (4, 0):
        for name, (argtypes, _) in lib._SIGNATURES.items():
            args = [None if t is ctypes.c_void_p else 1.0 if t in (ctypes.c_float, ctypes.c_double) else size for t in argtypes]
            rc = getattr(h, name)(*args)
            assert rc == -1 if size else rc in (0, -1), (name, size, rc)
