"""CPU tests (-m "not gpu") of nerf_signature_b200/raymarching/backend.py - the drop-in for the reference's
`raymarching/backend.py` (the pybind11 call surface of raymarching/src/raymarching.h:7-18 over the C ABI): the ten names,
their parameter lists, and that every argument reaches the C entry point in the position include/nsig.h gives the parameter
of the same name.  No kernel runs: `_lib.call` is replaced by a recorder."""
import ctypes
import inspect
import os
import re
import sys
import types

import numpy as np
import pytest
import torch

from nerf_signature_b200 import _lib
from nerf_signature_b200.raymarching import backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NSIG_REFERENCE", "/root/reference")
NAMES = ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
         "composite_rays_train_forward", "composite_rays_train_backward", "march_rays", "composite_rays"]


def _c_params(header_text, fname):
    """[(type, name)] of a C declaration `... fname(type name, ...)`."""
    m = re.search(r"\b%s\s*\(([^)]*)\)" % re.escape(fname), header_text)
    assert m, fname
    out = []
    for prm in m.group(1).split(","):
        toks = prm.replace("*", " * ").split()
        out.append((" ".join(toks[:-1]), toks[-1]))
    return out


def _header():
    return re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "nsig.h")).read(), flags=re.S)


def test_backend_has_the_ten_functions_with_the_c_abi_parameter_lists():
    h = _header()
    assert sorted(n for n in dir(backend._backend) if not n.startswith("_")) == sorted(NAMES)
    for name in NAMES:
        want = [p for _, p in _c_params(h, "nsig_" + name) if p not in ("scratch", "stream")]
        got = list(inspect.signature(getattr(backend._backend, name)).parameters)
        assert got == want, (name, got, want)


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "raymarching", "src", "raymarching.h")),
                    reason="needs the reference sources (build container only)")
def test_backend_parameter_lists_are_the_reference_header_s():
    """raymarching/src/raymarching.h:7-18: same function names, same parameter names in the same order - which is what lets
    the reference's raymarching.py call `_backend.<fn>(...)` positionally without an edit."""
    ref = open(os.path.join(REF, "raymarching", "src", "raymarching.h")).read()
    for name in NAMES:
        want = [p for _, p in _c_params(ref, name)]
        got = list(inspect.signature(getattr(backend._backend, name)).parameters)
        assert got == want, (name, got, want)
    # and the reference wrapper calls nothing else on its backend
    used = set(re.findall(r"_backend\.(\w+)\(", open(os.path.join(REF, "raymarching", "raymarching.py")).read()))
    assert used == set(NAMES)


def test_every_argument_reaches_the_c_parameter_of_the_same_name(monkeypatch):
    h = _header()
    calls = []
    monkeypatch.setattr(_lib, "call", lambda fname, *args: calls.append((fname, args)))
    monkeypatch.setattr(backend, "_P", lambda t: None if t is None else ("ptr", id(t)))
    for name in NAMES:
        cparams = _c_params(h, "nsig_" + name)
        kwargs, scalars = {}, {}
        for i, (ctype, pname) in enumerate(cparams):
            if pname in ("scratch", "stream"):
                continue
            if "*" in ctype:
                kwargs[pname] = torch.zeros(4)                      # a distinct tensor object per pointer parameter
            else:
                scalars[pname] = kwargs[pname] = (i + 2) if "int" in ctype else (i + 0.5)
        calls.clear()
        getattr(backend._backend, name)(**kwargs)
        assert len(calls) == 1 and calls[0][0] == "nsig_" + name
        args = calls[0][1]
        expect_n = len([1 for _, p in cparams if p != "stream"])    # _lib.call appends the stream itself
        assert len(args) == expect_n, (name, len(args), expect_n)
        for (ctype, pname), a in zip(cparams, args):
            if pname == "scratch":
                assert a is not None
            elif "*" in ctype:
                assert a == ("ptr", id(kwargs[pname])), (name, pname)
            else:
                assert a == scalars[pname] and type(a) is (int if "int" in ctype else float), (name, pname, a)


def test_backend_has_no_cpu_path():
    with pytest.raises(_lib.NsigError):
        backend._backend.morton3D(torch.zeros(4, 3, dtype=torch.int32), 4, torch.zeros(4, dtype=torch.int32))
    with pytest.raises(_lib.NsigError):
        backend._backend.packbits(torch.zeros(4, 16)[:, ::2], 1, 0.5, torch.zeros(1, dtype=torch.uint8))   # strided view


# ---------------------------------------------------------------------------------------------------------------------
# the reference's UNMODIFIED raymarching/raymarching.py on top of the shim
# ---------------------------------------------------------------------------------------------------------------------
class _HostPtr:
    """What `_P(tensor)` returns on the faked device: the raw host address for ctypes, plus the tensor itself for the two
    entry points that have no C-oracle counterpart and are emulated in Python."""

    def __init__(self, tensor):
        self.tensor = tensor
        self._as_parameter_ = ctypes.c_void_p(tensor.data_ptr())


@pytest.fixture
def fake_device(monkeypatch, oracle_cpu):
    """`.cuda()` is the identity, `_lib.call` dispatches to the C oracle (whose functions have the argument order of
    raymarching.h) and `_P` hands out host addresses: the Python layers above the C ABI run on a box without a GPU, the
    NUMBERS are the oracle's.  nsig_zero_sample_padding (no oracle counterpart) is emulated from its contract in nsig.h."""
    from nerf_signature_b200.raymarching import raymarching as pkg_rm
    olib = oracle_cpu.lib()

    def fake_call(name, *args):
        args = list(args)
        if name == "nsig_zero_sample_padding":
            xyzs, dirs, deltas, counter, align, M = [a.tensor if isinstance(a, _HostPtr) else a for a in args]
            m = int(counter[0])
            end = M if align == 0 else min(m + align - m % align, M)
            for buf in (xyzs, dirs, deltas):
                buf.view(-1, buf.shape[-1])[m:end] = 0
            return
        sig = list(_lib._SIGNATURES[name][0])[:-1]                    # without the stream
        if name == "nsig_march_rays_train":                           # the oracle needs no scratch
            args, sig = args[:-1], sig[:-1]
        assert len(args) == len(sig), (name, len(args), len(sig))
        fn = getattr(olib, "oracle_" + name[len("nsig_"):])
        fn.restype = None
        fn(*[a if t is ctypes.c_void_p else t(a) for a, t in zip(args, sig)])
        if name == "nsig_march_rays_train":
            # contract of nsig.h the reference kernel does not have (its wrapper zero-fills everything beforehand): when
            # rays are dropped for lack of room, the rows after the last kept ray are cleared by the kernel itself
            M, xyzs, dirs, deltas, rays, counter = args[9], args[12].tensor, args[13].tensor, args[14].tensor, args[15].tensor, args[16].tensor
            if int(counter[0]) > M:
                ends = (rays[:, 1] + rays[:, 2]).long()
                kept = ends[ends <= M]
                last = int(kept.max()) if kept.numel() else 0
                for buf in (xyzs, dirs, deltas):
                    buf[last:M] = 0

    ptr = lambda t: None if t is None else _HostPtr(t)                # noqa: E731
    monkeypatch.setattr(_lib, "call", fake_call)
    monkeypatch.setattr(backend, "_P", ptr)
    monkeypatch.setattr(pkg_rm, "_P", ptr)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    return oracle_cpu


def _load_reference_raymarching(monkeypatch):
    """raymarching/raymarching.py of the reference, unmodified, with `.backend` resolved to this package's shim."""
    pkg = types.ModuleType("ref_raymarching_pkg")
    pkg.__path__ = []
    monkeypatch.setitem(sys.modules, "ref_raymarching_pkg", pkg)
    monkeypatch.setitem(sys.modules, "ref_raymarching_pkg.backend", backend)
    monkeypatch.setitem(sys.modules, "_raymarching", None)            # the compiled extension is not importable
    rm = types.ModuleType("ref_raymarching_pkg.raymarching")
    rm.__package__ = "ref_raymarching_pkg"
    path = os.path.join(REF, "raymarching", "raymarching.py")
    exec(compile(open(path).read(), path, "exec"), rm.__dict__)
    assert rm._backend is backend._backend
    return rm


needs_reference = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "raymarching", "raymarching.py")),
                                     reason="needs the reference sources (build container only)")


@needs_reference
def test_reference_raymarching_module_runs_unmodified_on_the_backend(monkeypatch, fake_device):
    """The drop-in claim end to end, as far as a GPU-less box allows: the reference's own raymarching.py (autograd
    Functions, buffer allocation, `.item()` on the counter, alignment padding, in-place alive-ray state) is loaded
    unmodified with `.backend` resolved to this package's shim, and every call it makes lands in the C ABI with arguments
    the entry points accept.  The device is faked (fixture above), so the NUMBERS are the oracle's; what is tested is the
    calling convention between the real caller and the shim - shapes, dtypes, argument order, outputs written in place."""
    from nerf_signature_b200 import synthetic as syn
    oracle_cpu = fake_device
    rm = _load_reference_raymarching(monkeypatch)

    C, H, bound, max_steps, N = 2, 128, 2.0, 256, 200
    rng = np.random.default_rng(3)
    grid = syn.sphere_grid(C).astype(np.float32)
    grid *= (rng.uniform(size=grid.shape) < 0.9)
    rays_o, rays_d = syn.blender_rays(N, seed=3)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    t = torch.from_numpy
    # ---- utilities ----
    nears, fars = rm.near_far_from_aabb(t(rays_o), t(rays_d), t(aabb), 0.2)
    on, of = oracle_cpu.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    assert np.array_equal(nears.numpy(), on) and np.array_equal(fars.numpy(), of)
    coords = rng.integers(0, 128, size=(1000, 3)).astype(np.int32)
    idx = rm.morton3D(t(coords))
    assert np.array_equal(idx.numpy(), oracle_cpu.morton3D(coords))
    assert np.array_equal(rm.morton3D_invert(idx).numpy(), coords)
    bitfield = rm.packbits(t(grid), 0.5)
    assert bitfield.dtype == torch.uint8 and np.array_equal(bitfield.numpy(), oracle_cpu.packbits(grid.reshape(-1), 0.5))
    # ---- training: march (force_all_rays, align 128) + differentiable composite ----
    counter = torch.zeros(2, dtype=torch.int32)
    xyzs, dirs, deltas, rays = rm.march_rays_train(t(rays_o), t(rays_d), bound, bitfield, C, H, nears, fars, counter, -1, False, 128,
                                                   True, 0, max_steps)
    ox, od, odl, orays, ocnt = oracle_cpu.march_rays_train(rays_o, rays_d, bound, bitfield.numpy(), C, H, on, of, max_steps=max_steps)
    m = int(ocnt[0])
    assert np.array_equal(counter.numpy(), ocnt) and np.array_equal(rays.numpy(), orays)
    assert xyzs.shape[0] == m + 128 - m % 128 and np.array_equal(xyzs[:m].numpy(), ox[:m]) and not xyzs[m:].any()
    sig = torch.rand(xyzs.shape[0], requires_grad=True)
    rgb = torch.rand(xyzs.shape[0], 3, requires_grad=True)
    ws, depth, image = rm.composite_rays_train(sig, rgb, deltas, rays, 1e-4)
    ows, odepth, oimg = oracle_cpu.composite_rays_train_forward(sig.detach().numpy(), rgb.detach().numpy(), deltas.numpy(), orays, 1e-4)
    assert np.array_equal(ws.detach().numpy(), ows) and np.array_equal(image.detach().numpy(), oimg)
    g_ws, g_img = torch.randn(N), torch.randn(N, 3)
    (ws * g_ws).sum().add((image * g_img).sum()).backward()
    ogs, ogc = oracle_cpu.composite_rays_train_backward(g_ws.numpy(), g_img.numpy(), sig.detach().numpy(), rgb.detach().numpy(),
                                                        deltas.numpy(), orays, ows, oimg, 1e-4)
    assert np.array_equal(sig.grad.numpy(), ogs) and np.array_equal(rgb.grad.numpy(), ogc)
    # ---- inference: one iteration of the alive-ray loop, in-place state ----
    n_step = 4
    alive = torch.arange(N, dtype=torch.int32)
    rays_t = nears.clone()
    x2, d2, dl2 = rm.march_rays(N, n_step, alive, rays_t, t(rays_o), t(rays_d), bound, bitfield, C, H, nears, fars, 128, False, 0,
                                max_steps)
    assert x2.shape[0] == N * n_step + 128 - (N * n_step) % 128
    sg = torch.rand(x2.shape[0], n_step).T.reshape(-1)[:x2.shape[0]] * 20
    sg[n_step:2 * n_step] = 500.0                                     # ray 1 is opaque: killed by this call
    cl = torch.rand(x2.shape[0], 3)
    wsum, dep, img = torch.zeros(N), torch.zeros(N), torch.zeros(N, 3)
    o_alive, o_t = alive.numpy().copy(), rays_t.numpy().copy()
    o_ws, o_dep, o_img = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    rm.composite_rays(N, n_step, alive, rays_t, sg, cl, dl2, wsum, dep, img, 1e-2)
    oracle_cpu.composite_rays(N, n_step, o_alive, o_t, sg.numpy(), cl.numpy(), dl2.numpy(), o_ws, o_dep, o_img, 1e-2)
    assert np.array_equal(alive.numpy(), o_alive) and alive[1] == -1 and alive[0] == 0
    assert np.array_equal(wsum.numpy(), o_ws) and np.array_equal(img.numpy(), o_img) and np.array_equal(rays_t.numpy(), o_t)


@needs_reference
def test_package_wrappers_behave_like_the_reference_wrappers(monkeypatch, fake_device):
    """nerf_signature_b200.raymarching (the package's own autograd shells: no worst-case zero fill, only the padding rows
    cleared, no empty_cache) against the reference's raymarching.py on the same faked device: same shapes, same alignment
    padding rule (a full `align` is added when the count is already aligned), same zero rows, same counter updates, same
    in-place semantics, same gradients - in force_all_rays mode, in mean_count mode (fixed-size buffers, all rows returned)
    and when the buffers are too small and rays are dropped."""
    from nerf_signature_b200 import synthetic as syn
    from nerf_signature_b200 import raymarching as pkg
    ref = _load_reference_raymarching(monkeypatch)
    C, H, bound, max_steps, N = 2, 128, 2.0, 128, 150
    rng = np.random.default_rng(11)
    grid = syn.sphere_grid(C).astype(np.float32)
    grid *= (rng.uniform(size=grid.shape) < 0.9)
    rays_o, rays_d = syn.blender_rays(N, seed=11)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    t = torch.from_numpy

    def same(a, b):
        assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b)

    for fn_args in ((t(rays_o), t(rays_d), t(aabb), 0.2), (t(rays_o), t(rays_d), t(aabb))):
        for a, b in zip(pkg.near_far_from_aabb(*fn_args), ref.near_far_from_aabb(*fn_args)):
            same(a, b)
    nears, fars = ref.near_far_from_aabb(t(rays_o), t(rays_d), t(aabb), 0.2)
    same(pkg.sph_from_ray(t(rays_o), t(rays_d), 3.0), ref.sph_from_ray(t(rays_o), t(rays_d), 3.0))
    coords = t(rng.integers(0, 128, size=(777, 3)).astype(np.int32))
    same(pkg.morton3D(coords), ref.morton3D(coords))
    same(pkg.morton3D_invert(ref.morton3D(coords)), ref.morton3D_invert(ref.morton3D(coords)))
    bitfield = ref.packbits(t(grid), 0.5)
    same(pkg.packbits(t(grid), 0.5), bitfield)
    reuse_a, reuse_b = torch.full_like(bitfield, 255), torch.full_like(bitfield, 255)
    out_a, out_b = pkg.packbits(t(grid), 0.5, reuse_a), ref.packbits(t(grid), 0.5, reuse_b)        # caller's buffer, in place
    assert out_a.data_ptr() == reuse_a.data_ptr() and out_b.data_ptr() == reuse_b.data_ptr()
    same(reuse_a, reuse_b)

    def march(mod, mean_count, force_all_rays, align):
        counter = torch.zeros(2, dtype=torch.int32)
        out = mod.march_rays_train(t(rays_o), t(rays_d), bound, bitfield, C, H, nears, fars, counter, mean_count, False, align,
                                   force_all_rays, 0, max_steps)
        return out, counter

    total = int(march(ref, -1, True, -1)[1][0])
    assert total > 128
    for mean_count, force, align in ((-1, True, 128), (-1, True, -1), (0, False, 128), (total + 300, False, 128),
                                     (total, False, 128), (total // 2, False, 128), (total + 5, True, 64)):
        (pa, ca), (pb, cb) = march(pkg, mean_count, force, align), march(ref, mean_count, force, align)
        same(ca, cb)
        for a, b in zip(pa, pb):                                   # xyzs, dirs, deltas (padding rows included), rays
            same(a, b)
    # differentiable composite on the aligned buffers
    (xyzs, dirs, deltas, rays), _ = march(ref, -1, True, 128)
    base_s, base_c = torch.rand(xyzs.shape[0]) * 5, torch.rand(xyzs.shape[0], 3)
    g_ws, g_img = torch.randn(N), torch.randn(N, 3)
    res = []
    for mod in (pkg, ref):
        s, c = base_s.clone().requires_grad_(True), base_c.clone().requires_grad_(True)
        ws, depth, image = mod.composite_rays_train(s, c, deltas, rays, 1e-4)
        ((ws * g_ws).sum() + (image * g_img).sum()).backward()
        res.append((ws.detach(), depth.detach(), image.detach(), s.grad, c.grad))
    for a, b in zip(*res):
        same(a, b)
    # inference: three iterations of the alive-ray loop, state carried in place
    state = []
    for mod in (pkg, ref):
        alive, rays_t = torch.arange(N, dtype=torch.int32), nears.clone()
        wsum, dep, img = torch.zeros(N), torch.zeros(N), torch.zeros(N, 3)
        gen = torch.Generator().manual_seed(4)
        shapes = []
        for it in range(3):
            n_alive, n_step = alive.shape[0], 2 + it
            x, d, dl = mod.march_rays(n_alive, n_step, alive, rays_t, t(rays_o), t(rays_d), bound, bitfield, C, H, nears, fars,
                                      128, False, 0, max_steps)
            shapes.append(tuple(x.shape))
            sg = torch.rand(x.shape[0], generator=gen) * 36
            cl = torch.rand(x.shape[0], 3, generator=gen)
            mod.composite_rays(n_alive, n_step, alive, rays_t, sg, cl, dl, wsum, dep, img, 1e-2)
            alive = alive[alive >= 0]
        state.append((alive, rays_t, wsum, dep, img, torch.tensor(shapes)))
    for a, b in zip(*state):
        same(a, b)
    assert 0 < state[0][0].shape[0] < N                               # some rays were killed, some are still alive
