"""-m gpu: nerf_signature_b200/raymarching/backend.py - the `_backend` object a reference maintainer drops in for the
pybind11 extension - driven the way the reference's raymarching.py drives its backend (caller-allocated, zero-filled
buffers; in-place alive-ray state) and compared with the C oracle: integers and marched floats bit-exact, composites 2e-5
(the tolerances of tests/test_raymarching_gpu.py, which tests the same entry points through the package's own wrappers).
(File name sorts last on purpose: added at the very end of round 2, after the GPU budget was spent - its plumbing was
dry-run on CPU against the oracle through a recorder, tests/test_backend_shim_cpu.py covers the argument routing.)"""
import numpy as np
import pytest
import torch

from nerf_signature_b200 import synthetic as syn
from nerf_signature_b200.raymarching.backend import _backend

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def dev(a):
    return torch.from_numpy(np.array(a, copy=True)).cuda()       # never aliases the host array (in-place kernels)


def host(t):
    return t.cpu().numpy()


def _scene(N=500, C=2, bound=2.0, seed=3):
    rng = np.random.default_rng(seed)
    grid = syn.sphere_grid(C)
    grid *= (rng.uniform(size=grid.shape) < 0.9)                                   # speckled ball per cascade
    rays_o, rays_d = syn.blender_rays(N, seed=seed)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    return grid.astype(np.float32), rays_o, rays_d, aabb


def test_utilities_through_the_backend(oracle_cpu):
    grid, rays_o, rays_d, aabb = _scene()
    N = rays_o.shape[0]
    nears, fars = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    _backend.near_far_from_aabb(dev(rays_o), dev(rays_d), dev(aabb), N, 0.2, nears, fars)
    on, of = oracle_cpu.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    assert np.array_equal(bits(host(nears)), bits(on)) and np.array_equal(bits(host(fars)), bits(of))
    coords = np.random.default_rng(0).integers(0, 128, size=(4099, 3)).astype(np.int32)
    idx = torch.empty(coords.shape[0], dtype=torch.int32, device="cuda")
    _backend.morton3D(dev(coords), coords.shape[0], idx)
    assert np.array_equal(host(idx), oracle_cpu.morton3D(coords))
    back = torch.empty(coords.shape[0], 3, dtype=torch.int32, device="cuda")
    _backend.morton3D_invert(idx, coords.shape[0], back)
    assert np.array_equal(host(back), coords)
    n_bytes = grid.size // 8
    bitfield = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    _backend.packbits(dev(grid.reshape(-1)), n_bytes, 0.5, bitfield)
    assert np.array_equal(host(bitfield), oracle_cpu.packbits(grid.reshape(-1), 0.5))
    sph = torch.empty(N, 2, device="cuda")
    _backend.sph_from_ray(dev(rays_o), dev(rays_d), 4.0, N, sph)
    np.testing.assert_allclose(host(sph), oracle_cpu.sph_from_ray(rays_o, rays_d, 4.0), rtol=1e-4, atol=1e-5)


def test_training_march_and_composite_through_the_backend(oracle_cpu):
    C, H, bound, max_steps = 2, 128, 2.0, 256
    grid, rays_o, rays_d, aabb = _scene(C=C, bound=bound)
    N = rays_o.shape[0]
    bitfield = oracle_cpu.packbits(grid.reshape(-1), 0.5)
    on, of = oracle_cpu.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    M = N * max_steps
    # raymarching.py:205-216 of the reference: zero-filled sample buffers, zero counter, zero noises (perturb off)
    xyzs, dirs = torch.zeros(M, 3, device="cuda"), torch.zeros(M, 3, device="cuda")
    deltas = torch.zeros(M, 2, device="cuda")
    rays = torch.empty(N, 3, dtype=torch.int32, device="cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    noises = torch.zeros(N, device="cuda")
    _backend.march_rays_train(dev(rays_o), dev(rays_d), dev(bitfield), bound, 0.0, max_steps, N, C, H, M, dev(on), dev(of),
                              xyzs, dirs, deltas, rays, counter, noises)
    ox, od, odl, orays, ocnt = oracle_cpu.march_rays_train(rays_o, rays_d, bound, bitfield, C, H, on, of, max_steps=max_steps)
    m = int(ocnt[0])
    assert m > N and np.array_equal(host(counter), ocnt)
    assert np.array_equal(host(rays), orays)                    # ray n owns row n, offsets in ray order: the oracle's form
    assert np.array_equal(bits(host(xyzs[:m])), bits(ox[:m])) and np.array_equal(bits(host(deltas[:m])), bits(odl[:m]))
    assert np.array_equal(bits(host(dirs[:m])), bits(od[:m])) and not bool(xyzs[m:].any())

    rng = np.random.default_rng(1)
    sig = rng.uniform(0, 8, m).astype(np.float32)
    rgb = rng.uniform(0, 1, (m, 3)).astype(np.float32)
    ws, depth, image = torch.empty(N, device="cuda"), torch.empty(N, device="cuda"), torch.empty(N, 3, device="cuda")
    d_sig, d_rgb, d_del = dev(sig), dev(rgb), deltas[:m].contiguous()
    _backend.composite_rays_train_forward(d_sig, d_rgb, d_del, rays, m, N, 1e-4, ws, depth, image)
    ows, odepth, oimg = oracle_cpu.composite_rays_train_forward(sig, rgb, odl[:m], orays, 1e-4)
    np.testing.assert_allclose(host(ws), ows, rtol=0, atol=2e-5)
    np.testing.assert_allclose(host(image), oimg, rtol=0, atol=2e-5)
    np.testing.assert_allclose(host(depth), odepth, rtol=0, atol=2e-4)
    g_ws = rng.normal(size=N).astype(np.float32)
    g_img = rng.normal(size=(N, 3)).astype(np.float32)
    g_sig, g_rgb = torch.zeros(m, device="cuda"), torch.zeros(m, 3, device="cuda")     # raymarching.py:283-284
    _backend.composite_rays_train_backward(dev(g_ws), dev(g_img), d_sig, d_rgb, d_del, rays, ws, image, m, N, 1e-4, g_sig, g_rgb)
    ogs, ogc = oracle_cpu.composite_rays_train_backward(g_ws, g_img, sig, rgb, odl[:m], orays, ows, oimg, 1e-4)
    assert np.abs(host(g_sig) - ogs).max() <= 1e-3 * np.abs(ogs).max()
    assert np.abs(host(g_rgb) - ogc).max() <= 1e-4 * max(1.0, np.abs(ogc).max())


def test_inference_march_and_composite_through_the_backend(oracle_cpu):
    C, H, bound, max_steps, n_step = 2, 128, 2.0, 256, 4
    grid, rays_o, rays_d, aabb = _scene(N=300, C=C, bound=bound, seed=8)
    N = rays_o.shape[0]
    bitfield = oracle_cpu.packbits(grid.reshape(-1), 0.5)
    on, of = oracle_cpu.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    alive = np.arange(N, dtype=np.int32)
    rays_t = on.copy()
    M = N * n_step
    xyzs, dirs = torch.zeros(M, 3, device="cuda"), torch.zeros(M, 3, device="cuda")       # raymarching.py:333-335
    deltas = torch.zeros(M, 2, device="cuda")
    d_alive, d_t = dev(alive), dev(rays_t)
    _backend.march_rays(N, n_step, d_alive, d_t, dev(rays_o), dev(rays_d), bound, 0.0, max_steps, C, H, dev(bitfield), dev(on),
                        dev(of), xyzs, dirs, deltas, torch.zeros(N, device="cuda"))
    ox, od, odl = oracle_cpu.march_rays(N, n_step, alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, on, of,
                                        max_steps=max_steps)
    assert np.array_equal(bits(host(xyzs)), bits(ox[:M])) and np.array_equal(bits(host(deltas)), bits(odl[:M]))
    assert np.array_equal(bits(host(dirs)), bits(od[:M])) and bool(deltas.any())

    rng = np.random.default_rng(2)
    sig = rng.uniform(0, 20, (N, n_step)).astype(np.float32)    # thin rays stay alive ...
    sig[1::2] = rng.uniform(300, 600, sig[1::2].shape)          # ... every other ray is opaque: killed at once (T < 1e-2)
    sig = sig.reshape(-1)                                       # samples of alive ray i are rows [i * n_step, (i + 1) * n_step)
    rgb = rng.uniform(0, 1, (M, 3)).astype(np.float32)
    ws, depth, image = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda"), torch.zeros(N, 3, device="cuda")
    _backend.composite_rays(N, n_step, 1e-2, d_alive, d_t, dev(sig), dev(rgb), deltas, ws, depth, image)
    o_alive, o_t = alive.copy(), rays_t.copy()
    o_ws, o_depth, o_img = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    oracle_cpu.composite_rays(N, n_step, o_alive, o_t, sig, rgb, odl[:M], o_ws, o_depth, o_img, 1e-2)
    assert np.array_equal(host(d_alive), o_alive) and (o_alive < 0).any() and (o_alive >= 0).any()    # kill flags: exact
    np.testing.assert_allclose(host(ws), o_ws, rtol=0, atol=1e-4)
    np.testing.assert_allclose(host(image), o_img, rtol=0, atol=1e-4)
    np.testing.assert_allclose(host(d_t), o_t, rtol=0, atol=1e-4)
