"""-m gpu parity tests of csrc/grid.cu (fused occupancy sweep, partial-update cell selection, packbits with a
device-side threshold, mark_untrained_grid, get_rays) and of the half2 shadow tables, against oracle/grid_oracle.py
(+ the hash/MLP oracles for the densities) and the reference-generated fixture tests/golden/grid_golden.npz.

Tolerances: cell positions feed the density through the fp16-operand MLP -> grid values 2e-3 relative (as in
test_field_gpu.py); everything integer (cells, bitfield given the grid and the threshold, untrained mask) exact."""
import os

import numpy as np
import pytest
import torch

from oracle import grid_oracle as go

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net(bound, md, H, seed=0, thresh=10.0):
    from test_field_gpu import _net as make
    net = make(bound, md, seed=seed)
    net.grid_size = H
    net.density_thresh = thresh
    C = net.cascade
    net.density_grid = torch.zeros(C, H ** 3, device="cuda")
    net.density_bitfield = torch.zeros(C * H ** 3 // 8, dtype=torch.uint8, device="cuda")
    return net


def _oracle_sigma(net, pos, msg, oracle_cpu):
    from test_field_gpu import _oracle_forward
    dirs = np.zeros_like(pos); dirs[:, 2] = 1.0
    _, _, sigma, _, _ = _oracle_forward(net, pos, dirs, msg, oracle_cpu, density_scale=float(net.density_scale))
    return sigma.detach().numpy()


def _grid0(C, H, seed):
    rs = np.random.RandomState(seed)
    g = (rs.uniform(size=(C, H ** 3)) * 2.0).astype(np.float32)
    g[:, ::7] = -1.0     # untrained cells stay untouched
    g[:, 1::5] = 0.0
    return g


def _check_update(net, g0, cells, noise, msg, oracle_cpu, thresh):
    C, H = net.cascade, net.grid_size
    sig = []
    for cas in range(C):
        coords = go.morton3D_invert(cells[cas])
        pos = go.cell_positions(coords, noise[cas], cas, H, net.bound)
        sig.append(_oracle_sigma(net, pos, msg, oracle_cpu))
    g, mean, th, bits = go.update_extra_state(g0, cells, sig, thresh)
    got = net.density_grid.cpu().numpy()
    assert np.array_equal(got[g0 < 0], g0[g0 < 0])                         # untrained cells are never touched
    np.testing.assert_allclose(got, g, rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(net.mean_density, mean, rtol=2e-3)
    stats = net._last_stats.cpu().numpy()
    assert stats[1] == np.float32(min(float(stats[0]), thresh))
    # bitfield: exact given the grid and the threshold the device used
    assert np.array_equal(net.density_bitfield.cpu().numpy(), go.packbits(got, stats[1]))
    # and equal to the oracle's except for cells within the density tolerance of the threshold
    diff = np.unpackbits(net.density_bitfield.cpu().numpy() ^ bits, bitorder="little").astype(bool)
    near = np.abs(g.reshape(-1) - th) <= 4e-3 * abs(th) + 1e-6
    assert not (diff & ~near).any()


@pytest.mark.parametrize("bound,md,H,half2", [(1.0, 8, 32, True), (2.0, 4, 32, True), (1.0, 4, 16, False)])
def test_full_update_vs_oracle(oracle_cpu, bound, md, H, half2):
    net = _net(bound, md, H)
    net.half2_tables = half2
    C = net.cascade
    g0 = _grid0(C, H, 1)
    net.density_grid.copy_(torch.from_numpy(g0))
    noise = np.random.RandomState(2).uniform(size=(C, H ** 3, 3)).astype(np.float32)
    msg = np.random.RandomState(3).randint(0, 2, size=md).astype(np.float32)
    net.local_step = 4
    net.step_counter[:, 0] = torch.arange(16, dtype=torch.int32, device="cuda") * 100 + 7
    net.update_extra_state(torch.from_numpy(msg).cuda(), noise=torch.from_numpy(noise).cuda())
    assert net.iter_density == 1 and net.local_step == 0
    assert net.mean_count == int(sum(i * 100 + 7 for i in range(4)) / 4)   # renderer_wtmk.py:532-534
    cells = [np.arange(H ** 3)] * C                                        # the full sweep visits cells in Morton order
    _check_update(net, g0, cells, noise, msg, oracle_cpu, 10.0)


@pytest.mark.parametrize("bound,md,H", [(1.0, 8, 32), (2.0, 4, 16)])
def test_partial_update_given_cells_vs_oracle(oracle_cpu, bound, md, H):
    net = _net(bound, md, H, thresh=0.9)                                    # density_thresh below the mean: it wins
    C = net.cascade
    g0 = _grid0(C, H, 4)
    net.density_grid.copy_(torch.from_numpy(g0))
    rs = np.random.RandomState(5)
    n = H ** 3 // 2
    cells = rs.randint(0, H ** 3, size=(C, n))                              # with duplicates
    noise = rs.uniform(size=(C, n, 3)).astype(np.float32)
    msg = rs.randint(0, 2, size=md).astype(np.float32)
    net.iter_density = 16
    net.update_extra_state(torch.from_numpy(msg).cuda(), cells=torch.from_numpy(cells.astype(np.int32)).cuda(),
                           noise=torch.from_numpy(noise).cuda())
    _check_update(net, g0, list(cells), noise, msg, oracle_cpu, 0.9)


def test_golden_orchestration_through_the_kernels():
    """The reference's own update (fixture) replayed through nsig_grid_finalize / nsig_grid_pack: same EMA, mean,
    threshold and bitfield from the same recorded densities."""
    from nerf_signature_b200 import _lib
    P = _lib.ptr
    gold = np.load(os.path.join(ROOT, "tests", "golden", "grid_golden.npz"))
    H = 16
    for tag in ("b1", "b2"):
        g0 = gold[f"{tag}_grid0"]
        C = g0.shape[0]
        ar = np.arange(H)
        xx, yy, zz = np.meshgrid(ar, ar, ar, indexing="ij")
        cells = go.morton3D(np.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1))
        tmp = -np.ones_like(g0)
        for cas in range(C):
            tmp[cas][cells] = gold[f"{tag}_full_sigma"][cas]
        grid = torch.from_numpy(g0.copy()).cuda()
        sums = torch.zeros(1, dtype=torch.float64, device="cuda")
        stats = torch.empty(2, device="cuda")
        bits = torch.zeros(C * H ** 3 // 8, dtype=torch.uint8, device="cuda")
        _lib.call("nsig_grid_finalize", P(grid), P(torch.from_numpy(tmp).cuda()), C * H ** 3, 0.95, P(sums))
        _lib.call("nsig_grid_pack", P(grid), C * H ** 3 // 8, P(sums), C * H ** 3, 10.0, P(bits), P(stats))
        assert np.array_equal(grid.cpu().numpy(), gold[f"{tag}_full_grid"])
        np.testing.assert_allclose(float(stats[0]), gold[f"{tag}_full_mean"], rtol=1e-6)
        diff = np.unpackbits(bits.cpu().numpy() ^ gold[f"{tag}_full_bitfield"], bitorder="little").astype(bool)
        near = np.abs(gold[f"{tag}_full_grid"].reshape(-1) - float(stats[1])) <= 4e-7 * float(stats[1])
        assert not (diff & ~near).any()


def test_sample_cells_and_philox_jitter():
    from nerf_signature_b200 import _lib
    P = _lib.ptr
    C, H = 2, 32
    H3 = H ** 3
    rs = np.random.RandomState(0)
    g = rs.uniform(-1, 1, size=(C, H3)).astype(np.float32)
    g[1, :] = -1.0
    g[1, 5:200:3] = 0.5                                                     # few occupied cells in cascade 1
    grid = torch.from_numpy(g).cuda()
    n = H3 // 4
    out = []
    for seed in (11, 11, 12):
        cells = torch.empty(C, 2 * n, dtype=torch.int32, device="cuda")
        scratch = torch.empty(_lib.load().nsig_grid_sample_cells_scratch_bytes(C, H), dtype=torch.uint8, device="cuda")
        _lib.call("nsig_grid_sample_cells", P(grid), C, H, n, n, seed, P(cells), P(scratch))
        out.append(cells.cpu().numpy())
    assert np.array_equal(out[0], out[1]) and not np.array_equal(out[0], out[2])
    c = out[0]
    assert c.min() >= 0 and c.max() < H3
    for cas in range(C):
        occ = set(np.nonzero(g[cas] > 0)[0].tolist())
        assert set(c[cas, n:].tolist()) <= occ                              # second half: occupied cells only
        assert len(set(c[cas, n:].tolist())) > 0.3 * min(len(occ), n)
    # uniform half: every octant of the grid gets its share (8 octants, n draws -> n/8 +- 5 sigma)
    co = go.morton3D_invert(c[0, :n])
    octant = (co[:, 0] >= H // 2) * 4 + (co[:, 1] >= H // 2) * 2 + (co[:, 2] >= H // 2)
    cnt = np.bincount(octant, minlength=8)
    assert np.all(np.abs(cnt - n / 8) < 5 * np.sqrt(n / 8))
    # cascade with no occupied cell at all falls back to uniform cells instead of faulting
    grid.fill_(-1.0)
    cells = torch.empty(C, 2 * n, dtype=torch.int32, device="cuda")
    _lib.call("nsig_grid_sample_cells", P(grid), C, H, n, n, 3, P(cells), P(scratch))
    assert 0 <= int(cells.min()) and int(cells.max()) < H3

    # in-kernel Philox jitter: reproducible per seed, different across seeds, and inside the cell
    net = _net(1.0, 4, 16)
    res = []
    for seed in (5, 5, 6):
        torch.manual_seed(seed)
        net.density_grid.zero_(); net.iter_density = 0
        net.update_extra_state(None)
        res.append(net.density_grid.cpu().numpy().copy())
    assert np.array_equal(res[0], res[1]) and not np.array_equal(res[0], res[2])
    assert np.isfinite(res[0]).all() and (res[0] > 0).all()


@pytest.mark.parametrize("tag,bound", [("b1", 1), ("b2", 2)])
def test_mark_untrained_grid(tag, bound):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "grid_golden.npz"))
    net = _net(float(bound), 4, 16)
    poses, intr = gold[f"{tag}_mark_poses"], tuple(float(v) for v in gold[f"{tag}_mark_intrinsic"])
    net.mark_untrained_grid(poses, intr)
    got = net.density_grid.cpu().numpy()
    ref = gold[f"{tag}_mark_grid"]
    assert ((got == -1) != (ref == -1)).sum() <= 2                          # reference itself (CPU fixture)
    seen = go.mark_untrained_grid(poses, intr, net.cascade, 16, bound)
    assert ((got != -1) != seen).sum() <= 2                                 # oracle, CUDA division form
    # production size, many cameras
    net = _net(float(bound), 4, 128)
    rs = np.random.RandomState(1)
    P = np.tile(np.eye(4, dtype=np.float32), (70, 1, 1))
    P[:, :3, 3] = rs.normal(size=(70, 3)) * bound
    q, _ = np.linalg.qr(rs.normal(size=(70, 3, 3)))
    P[:, :3, :3] = q
    net.mark_untrained_grid(torch.from_numpy(P), (300.0, 310.0, 64.0, 60.0))
    seen = go.mark_untrained_grid(P, (300.0, 310.0, 64.0, 60.0), net.cascade, 128, bound)
    got = net.density_grid.cpu().numpy() != -1
    assert 0 < got.sum() < got.size
    assert (got != seen).sum() <= 1e-5 * got.size


def test_get_rays():
    from nerf_signature_b200.nerf.rays import get_rays
    gold = np.load(os.path.join(ROOT, "tests", "golden", "grid_golden.npz"))
    Hh, Ww = (int(v) for v in gold["rays_HW"])
    intr = gold["rays_intr"]
    poses = torch.from_numpy(gold["rays_poses"]).cuda()
    # all pixels
    r = get_rays(poses, intr, Hh, Ww, -1)
    o, d = go.get_rays(gold["rays_poses"], tuple(float(v) for v in intr), Hh, Ww, None)
    np.testing.assert_allclose(r["rays_d"].cpu().numpy(), d, rtol=0, atol=2e-7)
    np.testing.assert_allclose(r["rays_d"].cpu().numpy(), gold["rays_all_d"], rtol=0, atol=1e-6)
    assert np.array_equal(r["rays_o"].cpu().numpy(), gold["rays_all_o"])
    assert r["inds"].shape == (3, Hh * Ww)
    # random pixels: same generator calls as the reference
    torch.manual_seed(0)
    r = get_rays(poses, intr, Hh, Ww, 64)
    inds = r["inds"].cpu().numpy()
    assert inds.shape == (3, 64) and inds.min() >= 0 and inds.max() < Hh * Ww
    o, d = go.get_rays(gold["rays_poses"], tuple(float(v) for v in intr), Hh, Ww, inds)
    np.testing.assert_allclose(r["rays_d"].cpu().numpy(), d, rtol=0, atol=2e-7)
    # patches: 4x4 blocks of neighbouring pixels
    r = get_rays(poses, intr, Hh, Ww, 64, None, 4)
    p = r["inds"][0].cpu().numpy().reshape(4, 4, 4)
    assert np.array_equal(p - p[:, :1, :1], np.broadcast_to((np.arange(4)[:, None] * Ww + np.arange(4)[None, :]), (4, 4, 4)))
    # error-map sampling
    em = torch.rand(3, 128 * 128)
    r = get_rays(poses, intr, Hh, Ww, 32, em)
    assert r["inds_coarse"].shape == (3, 32) and r["rays_d"].shape == (3, 32, 3)
    o, d = go.get_rays(gold["rays_poses"], tuple(float(v) for v in intr), Hh, Ww, r["inds"].cpu().numpy())
    np.testing.assert_allclose(r["rays_d"].cpu().numpy(), d, rtol=0, atol=2e-7)


def test_half2_shadow_tables():
    """nsig_tables_to_half2: power-of-two scaling, error <= 2^-11 of the level maximum, refresh on table change; the
    fused forward with shadow tables stays within the path's tolerance of the fp32-table forward."""
    from test_field_gpu import _net as make, _points
    net = make(1.0, 8)
    sh = net.encoder.half_tables()
    assert net.encoder.half_tables() is sh                                   # cached
    inv = sh.inv_scale.cpu().numpy()
    for l, (t, h) in enumerate(zip(net.encoder.tables(), sh.tables)):
        t = t.detach().cpu().numpy(); h = h.float().cpu().numpy()
        m = np.abs(t).max()
        assert np.log2(inv[l]) == np.round(np.log2(inv[l]))                  # exact power of two
        assert 2.0 ** 14 <= m / inv[l] < 2.0 ** 15
        assert np.abs(h * inv[l] - t).max() <= 2.0 ** -11 * m
    with torch.no_grad():
        net.encoder.embeddings[3].weight.mul_(4.0)                           # version bump -> rebuilt
    sh2 = net.encoder.half_tables()
    assert sh2.inv_scale.cpu().numpy()[3] == 4 * inv[3]
    x, dirs = _points(4000, 1.0, 7)
    xt, dt = torch.from_numpy(x).cuda(), torch.from_numpy(dirs).cuda()
    msg = torch.from_numpy(np.random.RandomState(1).randint(0, 2, size=8).astype(np.float32)).cuda()
    with torch.no_grad():
        net.half2_tables = True
        s1, c1 = net(xt, dt, msg)
        net.half2_tables = False
        s0, c0 = net(xt, dt, msg)
    np.testing.assert_allclose(s1.cpu().numpy(), s0.cpu().numpy(), rtol=1e-3)
    np.testing.assert_allclose(c1.cpu().numpy(), c0.cpu().numpy(), rtol=0, atol=1e-3)
    assert not torch.equal(s1, s0)                                           # the shadow path really ran
