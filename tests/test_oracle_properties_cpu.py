"""CPU tests (-m "not gpu"): size-independent properties of the oracle itself.

tests/test_oracle_cpu.py pins the oracle to what the REFERENCE produced on a handful of golden cases; the
full-size GPU tests (tests/test_full_size_gpu.py) then lean on properties instead of stored outputs.  This file
checks that the oracle has those properties on inputs the golden cases do not contain (other seeds, other cascades,
ragged and empty batches), so that a property violated on the GPU can be blamed on the kernel, not on the checker:

  * morton3D / morton3D_invert are mutually inverse on the whole 128^3 lattice (raymarching.cu:214-254);
  * packbits is the strict `>` of raymarching.cu:268-289, bit k of byte j <-> cell 8j+k;
  * march_rays_train (raymarching.cu:312-480): rows of `rays` are (ray, exclusive-scan offset, count); every emitted
    sample lies inside the box in an occupied cell, t grows along a ray, an empty grid emits nothing;
  * composite_rays_train (raymarching.cu:501-682): weights_sum in [0, 1], image linear in rgb, zero density -> nothing,
    and the backward is the gradient of the forward (central differences in fp64 over the fp32 oracle);
  * hash encoders (hash_encoding.py:96-111, hash_encoding_wtmk_bit.py:99-116): features are linear in the table
    values, the backward is the adjoint of the forward (<f(T), g> == <T, backward(g)>), and the message form equals the
    plain one-level encode of the sum of the selected tables.
"""
import numpy as np
import pytest

SEEDS = (0, 1, 2)


def _rays(rng, n, radius=2.2, inside=False):
    if inside:
        o = rng.uniform(-0.5, 0.5, (n, 3))
    else:
        o = rng.normal(size=(n, 3))
        o = radius * o / np.linalg.norm(o, axis=1, keepdims=True)
    tgt = rng.uniform(-0.6, 0.6, (n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)


def _sphere_grid(rng, C, H, fill=0.35):
    """[C, H^3] densities in MORTON order: a ball per cascade plus speckle, like a trained grid."""
    idx = np.arange(H ** 3, dtype=np.int32)
    return idx, rng.uniform(0.0, 1.0, (C, H ** 3)).astype(np.float32) * (rng.uniform(size=(C, H ** 3)) < fill)


# ---------------------------------------------------------------------------------------------------
def test_morton_is_a_bijection_of_the_128_cube(oracle_cpu):
    H = 128
    idx = np.arange(H ** 3, dtype=np.int32)
    xyz = oracle_cpu.morton3D_invert(idx)
    assert xyz.min() == 0 and xyz.max() == H - 1
    assert np.array_equal(oracle_cpu.morton3D(xyz), idx)
    # bit interleave: x owns bits 0,3,6..., y bits 1,4,7..., z bits 2,5,8...
    one = oracle_cpu.morton3D(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [2, 0, 0], [127, 127, 127]], np.int32))
    assert one.tolist() == [1, 2, 4, 8, H ** 3 - 1]


@pytest.mark.parametrize("seed", SEEDS)
def test_packbits_is_the_strict_threshold_bit_k_of_byte_j(oracle_cpu, seed):
    rng = np.random.default_rng(seed)
    grid = rng.uniform(0, 1, 8 * 4099).astype(np.float32)
    thresh = np.float32(0.5)
    grid[::7] = thresh                      # ties are NOT occupied (strict >)
    grid[3::11] = -1.0                      # untrained cells
    packed = oracle_cpu.packbits(grid, float(thresh))
    want = np.packbits((grid > thresh).reshape(-1, 8), axis=1, bitorder="little").reshape(-1)
    assert np.array_equal(packed, want)
    assert not oracle_cpu.packbits(np.full(64, thresh, np.float32), float(thresh)).any()


@pytest.mark.parametrize("seed", SEEDS)
def test_near_far_brackets_the_box(oracle_cpu, seed):
    rng = np.random.default_rng(seed)
    o, d = _rays(rng, 513)
    bound = 1.0
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = oracle_cpu.near_far_from_aabb(o, d, aabb, 0.2)
    hit = fars < 1e8
    assert hit.all()                        # every ray is aimed at the box
    assert (nears >= 0.2).all() and (fars > nears).all()
    for t, tol in ((nears, 1e-4), (fars, 1e-4)):
        p = o + d * t[:, None]
        assert (np.abs(p) <= bound + tol).all()
        assert (np.abs(np.abs(p).max(axis=1) - bound) < 1e-3).all()      # entry and exit points lie ON a face
    # rays that miss: near = far = the reference's sentinel (float max), never marched
    o2 = np.array([[3.0, 3.0, 3.0]], np.float32)
    d2 = np.array([[1.0, 0.0, 0.0]], np.float32)
    n2, f2 = oracle_cpu.near_far_from_aabb(o2, d2, aabb, 0.2)
    assert n2[0] == f2[0] and n2[0] > 1e30
    # origins inside the box: near is clamped to min_near
    oi, di = _rays(rng, 64, inside=True)
    ni, fi = oracle_cpu.near_far_from_aabb(oi, di, aabb, 0.2)
    assert np.array_equal(ni, np.full(64, 0.2, np.float32)) and (fi > ni).all()


@pytest.mark.parametrize("seed,C,bound,dt_gamma", [(0, 1, 1.0, 0.0), (1, 2, 2.0, 0.0), (2, 3, 4.0, 1.0 / 128), (3, 1, 1.0, 0.0)])
def test_march_rows_offsets_and_sample_positions(oracle_cpu, seed, C, bound, dt_gamma):
    rng = np.random.default_rng(seed)
    H, N, max_steps = 32, 257, 256
    _, dens = _sphere_grid(rng, C, H)
    bitfield = oracle_cpu.packbits(dens.reshape(-1), 0.5)
    o, d = _rays(rng, N, radius=1.9 * bound if bound > 1 else 2.2)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = oracle_cpu.near_far_from_aabb(o, d, aabb, 0.2)
    noises = rng.uniform(size=N).astype(np.float32) if seed != 3 else None
    xyzs, dirs, deltas, rays, counter = oracle_cpu.march_rays_train(o, d, bound, bitfield, C, H, nears, fars, noises,
                                                                    dt_gamma=dt_gamma, max_steps=max_steps)
    assert counter[1] == N
    # the oracle emits rows in ray order: (ray, exclusive scan of the counts, count)
    assert np.array_equal(rays[:, 0], np.arange(N))
    counts = rays[:, 2]
    assert (counts >= 0).all() and (counts <= max_steps).all() and counts.sum() == counter[0] > 0
    assert np.array_equal(rays[:, 1], np.concatenate([[0], np.cumsum(counts)[:-1]]))
    M = int(counter[0])
    assert (np.abs(xyzs[:M]) <= bound).all()
    assert not xyzs[M:].any() and not deltas[M:].any()                    # zero-filled tail, like the reference wrapper
    assert (deltas[:M, 0] > 0).all()
    for n in np.flatnonzero(counts)[:64]:
        lo, c = rays[n, 1], counts[n]
        assert np.allclose(dirs[lo:lo + c], d[n])
        t = ((xyzs[lo:lo + c] - o[n]) * d[n]).sum(axis=1)
        assert (np.diff(t) > 0).all() and t[0] >= nears[n] - 1e-4 and t[-1] <= fars[n] + 1e-4
        # deltas[:, 0] is the step dt taken AFTER the sample, deltas[:, 1] = (t + dt) - last_t with last_t the previous
        # sample's t + dt (raymarching.cu:463-470): equal to the sample spacing only while dt is constant
        end = t + deltas[lo:lo + c, 0]
        assert np.allclose(np.diff(end), deltas[lo + 1:lo + c, 1], atol=5e-5)
    # every emitted sample sits in an occupied cell of the cascade its position and step size select
    p = xyzs[:M]
    dt = deltas[:M, 0]
    mx = np.abs(p).max(axis=1)
    with np.errstate(divide="ignore"):
        lvl_pos = np.ceil(np.log2(np.maximum(mx, 1e-30)))
        lvl_dt = np.ceil(np.log2(np.maximum(dt * H * 0.5, 1e-30)))        # mip_from_dt: dt * H/2 (raymarching.cu:72-81)
    lvl = np.clip(np.maximum(lvl_pos, lvl_dt), 0, C - 1).astype(np.int64)
    mip_bound = np.minimum(2.0 ** lvl, bound)
    cell = np.clip((0.5 * (p / mip_bound[:, None] + 1) * H).astype(np.int64), 0, H - 1)
    lin = lvl * H ** 3 + oracle_cpu.morton3D(cell.astype(np.int32)).astype(np.int64)
    occupied = (bitfield[lin // 8] >> (lin % 8)) & 1
    assert occupied.all()
    # an empty grid emits nothing, whatever the rays
    _, _, _, rays0, counter0 = oracle_cpu.march_rays_train(o, d, bound, np.zeros_like(bitfield), C, H, nears, fars,
                                                           noises, dt_gamma=dt_gamma, max_steps=max_steps)
    assert counter0[0] == 0 and not rays0[:, 2].any() and not rays0[:, 1].any()


def test_march_is_a_function_of_each_ray_alone(oracle_cpu):
    """Shuffling the batch shuffles the per-ray sample runs and nothing else (what lets tests compare a
    deterministic-order kernel with the reference's atomicAdd order in canonical form)."""
    rng = np.random.default_rng(5)
    H, N, C, bound = 32, 200, 2, 2.0
    _, dens = _sphere_grid(rng, C, H)
    bitfield = oracle_cpu.packbits(dens.reshape(-1), 0.5)
    o, d = _rays(rng, N, radius=3.0)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = oracle_cpu.near_far_from_aabb(o, d, aabb, 0.2)
    noises = rng.uniform(size=N).astype(np.float32)
    a = oracle_cpu.march_rays_train(o, d, bound, bitfield, C, H, nears, fars, noises, max_steps=128)
    perm = rng.permutation(N)
    b = oracle_cpu.march_rays_train(o[perm], d[perm], bound, bitfield, C, H, nears[perm], fars[perm], noises[perm],
                                    max_steps=128)
    assert np.array_equal(a[3][perm, 2], b[3][:, 2])
    for j in range(0, N, 7):
        n = perm[j]
        la, lb, c = a[3][n, 1], b[3][j, 1], a[3][n, 2]
        assert np.array_equal(a[0][la:la + c].view(np.uint32), b[0][lb:lb + c].view(np.uint32))
        assert np.array_equal(a[2][la:la + c].view(np.uint32), b[2][lb:lb + c].view(np.uint32))


# ---------------------------------------------------------------------------------------------------
def _ragged(rng, N, max_c):
    counts = rng.integers(0, max_c, N).astype(np.int32)
    counts[rng.integers(0, N, 3)] = 0                                        # rays that hit nothing
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays = np.stack([np.arange(N, dtype=np.int32), offs, counts], axis=1)
    M = int(counts.sum())
    sig = rng.uniform(0, 6, M).astype(np.float32)
    rgb = rng.uniform(0, 1, (M, 3)).astype(np.float32)
    dl = np.stack([rng.uniform(0.01, 0.05, M), rng.uniform(0.01, 0.2, M)], axis=1).astype(np.float32)
    return rays, sig, rgb, dl


@pytest.mark.parametrize("seed", SEEDS)
def test_composite_forward_properties(oracle_cpu, seed):
    rng = np.random.default_rng(seed)
    rays, sig, rgb, dl = _ragged(rng, 97, 40)
    ws, depth, img = oracle_cpu.composite_rays_train_forward(sig, rgb, dl, rays)
    assert (ws >= 0).all() and (ws <= 1 + 1e-6).all()
    empty = rays[:, 2] == 0
    assert not ws[empty].any() and not img[empty].any() and not depth[empty].any()
    # closed form per ray, in fp64, with the reference's early termination (T < T_thresh stops AFTER the sample)
    for n in range(rays.shape[0]):
        lo, c = rays[n, 1], rays[n, 2]
        T, w_sum, acc = 1.0, 0.0, np.zeros(3)
        for s in range(lo, lo + c):
            alpha = 1.0 - np.exp(-float(sig[s]) * float(dl[s, 0]))
            w = alpha * T
            acc += w * rgb[s]
            w_sum += w
            T *= 1.0 - alpha
            if T < 1e-4:
                break
        assert abs(ws[n] - w_sum) < 2e-5 and np.abs(img[n] - acc).max() < 2e-5
    # linear in rgb, density-free samples contribute nothing
    ws2, _, img2 = oracle_cpu.composite_rays_train_forward(sig, 0.5 * rgb, dl, rays)
    assert np.array_equal(ws2, ws) and np.allclose(img2, 0.5 * img, atol=1e-6)
    ws0, d0, img0 = oracle_cpu.composite_rays_train_forward(np.zeros_like(sig), rgb, dl, rays)
    assert not ws0.any() and not img0.any() and not d0.any()
    # no rays / no samples: nothing to do, nothing written out of bounds
    e = oracle_cpu.composite_rays_train_forward(np.zeros(0), np.zeros((0, 3)), np.zeros((0, 2)), np.zeros((0, 3), np.int32))
    assert all(a.size == 0 for a in e)


def test_composite_backward_is_the_gradient_of_the_forward(oracle_cpu):
    rng = np.random.default_rng(11)
    rays, sig, rgb, dl = _ragged(rng, 12, 9)
    sig = (sig * 0.5).astype(np.float32)            # keep T above T_thresh: the truncation point is not differentiable
    g_ws = rng.normal(size=rays.shape[0]).astype(np.float32)
    g_img = rng.normal(size=(rays.shape[0], 3)).astype(np.float32)
    ws, _, img = oracle_cpu.composite_rays_train_forward(sig, rgb, dl, rays)
    gs, gc = oracle_cpu.composite_rays_train_backward(g_ws, g_img, sig, rgb, dl, rays, ws, img)

    def loss(s, c):                                  # fp64 restatement of the same sum (no truncation reached)
        total = 0.0
        for n in range(rays.shape[0]):
            lo, cnt = rays[n, 1], rays[n, 2]
            alpha = 1.0 - np.exp(-s[lo:lo + cnt] * dl[lo:lo + cnt, 0].astype(np.float64))
            T = np.concatenate([[1.0], np.cumprod(1.0 - alpha)[:-1]])
            w = alpha * T
            total += g_ws[n] * w.sum() + (g_img[n] * (w[:, None] * c[lo:lo + cnt]).sum(axis=0)).sum()
        return total

    s64, c64 = sig.astype(np.float64), rgb.astype(np.float64)
    eps = 1e-6
    num_s = np.zeros_like(s64)
    for i in range(s64.size):
        a, b = s64.copy(), s64.copy()
        a[i] += eps
        b[i] -= eps
        num_s[i] = (loss(a, c64) - loss(b, c64)) / (2 * eps)
    assert np.abs(gs - num_s).max() < 1e-3 * max(1.0, np.abs(num_s).max())
    num_c = np.zeros_like(c64)
    for i in range(0, c64.shape[0], 3):
        for k in range(3):
            a = c64.copy()
            a[i, k] += 1.0                           # linear in rgb: one-sided unit step is exact
            num_c[i, k] = loss(s64, a) - loss(s64, c64)
    sel = np.arange(0, c64.shape[0], 3)
    assert np.abs(gc[sel] - num_c[sel]).max() < 1e-4


# ---------------------------------------------------------------------------------------------------
def _resolutions(n_levels=16, base=16.0, finest=2048.0):
    b = np.exp((np.log(np.float32(finest)) - np.log(np.float32(base))) / np.float32(n_levels - 1)).astype(np.float32)
    return np.floor(np.float32(base) * b ** np.arange(n_levels, dtype=np.float32)).astype(np.float32)


@pytest.mark.parametrize("seed", SEEDS)
def test_hash_encode_is_linear_in_the_tables_and_backward_is_its_adjoint(oracle_cpu, seed):
    rng = np.random.default_rng(seed)
    L, log2_T, B = 16, 12, 300
    res = _resolutions(L)
    x = rng.uniform(0, 1, (B, 3)).astype(np.float32)                         # the encoders' box is (0, 1)
    x[:4] = [[0, 0, 0], [1, 1, 1], [0.5, 0.25, 0.125], [1, 0, 0.5]]          # box corners / cell boundaries
    T1 = [rng.normal(size=(1 << log2_T, 2)).astype(np.float32) for _ in range(L)]
    T2 = [rng.normal(size=(1 << log2_T, 2)).astype(np.float32) for _ in range(L)]
    f1, slots = oracle_cpu.hash_encode_forward(x, T1, res, log2_T, want_slots=True)
    f2 = oracle_cpu.hash_encode_forward(x, T2, res, log2_T)
    f12 = oracle_cpu.hash_encode_forward(x, [2 * a + b for a, b in zip(T1, T2)], res, log2_T)
    assert slots.min() >= 0 and slots.max() < (1 << log2_T)
    assert np.allclose(f12, 2 * f1 + f2, atol=2e-5)
    # constant tables interpolate to the constant (the 8 trilinear weights sum to 1)
    ones = oracle_cpu.hash_encode_forward(x, [np.ones((1 << log2_T, 2), np.float32)] * L, res, log2_T)
    assert np.abs(ones - 1).max() < 1e-5
    # adjoint: <forward(T), g> == <T, backward(g)>
    g = rng.normal(size=(B, 2 * L)).astype(np.float32)
    dT = oracle_cpu.hash_encode_backward(x, g, res, L, log2_T)
    lhs = float((f1.astype(np.float64) * g).sum())
    rhs = float(sum((t.astype(np.float64) * d).sum() for t, d in zip(T1, dT)))
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))
    # the gradient only touches slots the forward read
    for l in (0, 7, 15):
        touched = np.zeros(1 << log2_T, bool)
        touched[slots[:, l].reshape(-1)] = True
        assert not dT[l][~touched].any()


@pytest.mark.parametrize("seed,md", [(0, 8), (1, 32), (2, 48)])
def test_message_encode_is_the_one_level_encode_of_the_selected_tables_sum(oracle_cpu, seed, md):
    rng = np.random.default_rng(seed)
    log2_T, B, resolution = 12, 257, 2048.0
    x = rng.uniform(0, 1, (B, 3)).astype(np.float32)
    tables = [rng.normal(scale=1e-2, size=(1 << log2_T, 2)).astype(np.float32) for _ in range(2 * md)]
    message = rng.integers(0, 2, md).astype(np.float32)
    out = oracle_cpu.msg_encode_forward(x, tables, message, resolution, log2_T)
    S = np.zeros_like(tables[0], dtype=np.float64)
    for i in range(md):
        S += tables[2 * i + int(message[i])]
    pre = oracle_cpu.hash_encode_forward(x, [S.astype(np.float32)], np.array([resolution], np.float32), log2_T)
    assert np.abs(out - pre).max() < 1e-5 * md ** 0.5 + 2e-6                 # summation order differs
    # flipping one bit changes the result by exactly that pair's difference, interpolated
    flip = message.copy()
    flip[md // 2] = 1 - flip[md // 2]
    out2 = oracle_cpu.msg_encode_forward(x, tables, flip, resolution, log2_T)
    i = md // 2
    dlt = tables[2 * i + int(flip[i])].astype(np.float64) - tables[2 * i + int(message[i])]
    want = oracle_cpu.hash_encode_forward(x, [dlt.astype(np.float32)], np.array([resolution], np.float32), log2_T)
    assert np.abs((out2 - out) - want).max() < 1e-5
    # backward: every selected table receives the SAME gradient G; adjoint of the one-level encode
    g = rng.normal(size=(B, 2)).astype(np.float32)
    G = oracle_cpu.msg_encode_backward(x, g, resolution, log2_T)
    G1 = oracle_cpu.hash_encode_backward(x, g, np.array([resolution], np.float32), 1, log2_T)[0]
    assert np.array_equal(G.view(np.uint32), G1.view(np.uint32))
    lhs = float((out.astype(np.float64) * g).sum())
    rhs = float((S * G).sum())
    assert abs(lhs - rhs) < 1e-3 * max(1e-3, abs(lhs))
