"""-m gpu parity tests for the hash encoders through the drop-in modules
(nerf_signature_b200.hash_encoding[_wtmk_bit].HashEmbedder -> C ABI), against the golden vectors
generated from the reference modules (tests/golden/hash_golden.npz) and the C oracle.
Hash slots: bit-exact.  Base-encoder features: bit-exact (same fp32 op sequence).  Message features in
the pre-summed form: 1e-5 relative to the feature scale (summation order differs from torch.sum)."""
import numpy as np
import pytest
import torch

from make_golden_hash import make_tables

pytestmark = pytest.mark.gpu


def _load_base(tag, g):
    from nerf_signature_b200.hash_encoding import HashEmbedder
    log2_T = int(g[f"base_{tag}_log2T"])
    enc = HashEmbedder(bounding_box=(0, 1), n_levels=16, n_features_per_level=2, log2_hashmap_size=log2_T,
                       base_resolution=16, finest_resolution=2048)
    tabs = make_tables(int(g[f"base_{tag}_seed"]), 16, log2_T)
    with torch.no_grad():
        for i in range(16):
            enc.embeddings[i].weight.copy_(torch.from_numpy(tabs[i]))
    return enc.cuda(), tabs


@pytest.mark.parametrize("tag", ["small", "full"])
def test_base_encoder_golden(golden_hash, tag):
    g = golden_hash
    enc, _ = _load_base(tag, g)
    assert np.array_equal(np.asarray(enc.resolutions, np.float32), g[f"base_{tag}_res"])
    x = torch.from_numpy(g[f"base_{tag}_x"]).cuda()
    slots = enc.hashed_indices(x).cpu().numpy()
    assert np.array_equal(slots, g[f"base_{tag}_slots"])
    feat = enc(x).detach().cpu().numpy()
    assert np.array_equal(feat.view(np.uint32), g[f"base_{tag}_feat"].view(np.uint32))


def test_base_encoder_backward_golden(golden_hash):
    g = golden_hash
    enc, _ = _load_base("small", g)
    x = torch.from_numpy(g["base_small_x"]).cuda()
    out = enc(x)
    out.backward(torch.from_numpy(g["base_small_gout"]).cuda())
    got = np.stack([e.weight.grad.cpu().numpy() for e in enc.embeddings])
    ref = g["base_small_gtab"]
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())
    assert np.array_equal(got == 0, ref == 0)


def test_base_encoder_large_vs_oracle(oracle_cpu):
    """BASELINE-sized tables (2^19), 200k ray-coherent points: slots and features bit-exact vs the C oracle."""
    from nerf_signature_b200.hash_encoding import HashEmbedder
    torch.manual_seed(0)
    enc = HashEmbedder(bounding_box=(0, 1), n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                       base_resolution=16, finest_resolution=2048).cuda()
    rs = np.random.RandomState(1)
    o = rs.uniform(0.1, 0.9, size=(400, 1, 3)); d = rs.normal(size=(400, 1, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    t = np.arange(500).reshape(1, -1, 1) * 0.0017
    x = np.clip(o + d * t, 0, 1).reshape(-1, 3).astype(np.float32)
    tabs = [e.weight.detach().cpu().numpy() for e in enc.embeddings]
    want, wslots = oracle_cpu.hash_encode_forward(x, tabs, enc.resolutions, 19, want_slots=True)
    xt = torch.from_numpy(x).cuda()
    assert np.array_equal(enc.hashed_indices(xt).cpu().numpy(), wslots)
    assert np.array_equal(enc(xt).detach().cpu().numpy().view(np.uint32), want.view(np.uint32))
    # empty batch
    assert enc(torch.zeros(0, 3, device="cuda")).shape == (0, 32)


@pytest.mark.parametrize("tag", ["small", "md32", "md48"])
def test_msg_encoder_golden(golden_hash, tag):
    from nerf_signature_b200.hash_encoding_wtmk_bit import HashEmbedder
    g = golden_hash
    log2_T, md = int(g[f"msg_{tag}_log2T"]), int(g[f"msg_{tag}_md"])
    enc = HashEmbedder(bounding_box=(0, 1), n_levels=md * 2, n_features_per_level=2, log2_hashmap_size=log2_T,
                       base_resolution=2048, finest_resolution=2048, message_dim=md)
    tabs = make_tables(int(g[f"msg_{tag}_seed"]), 2 * md, log2_T)
    with torch.no_grad():
        for i in range(2 * md):
            enc.embeddings[i].weight.copy_(torch.from_numpy(tabs[i]))
    enc = enc.cuda()
    x = torch.from_numpy(g[f"msg_{tag}_x"]).cuda()
    msg = torch.from_numpy(g[f"msg_{tag}_message"]).cuda()
    ref = g[f"msg_{tag}_feat"]
    scale = np.abs(ref).max()
    pre = enc(x, msg)                      # pre-summed form (product path)
    per = enc.forward_perbit(x, msg)       # reference-form evaluation order
    np.testing.assert_allclose(per.detach().cpu().numpy(), ref, rtol=0, atol=2e-6 * scale)
    np.testing.assert_allclose(pre.detach().cpu().numpy(), ref, rtol=0, atol=1e-5 * scale)
    if tag == "small":
        pre.backward(torch.from_numpy(g["msg_small_gout"]).cuda())
        gref = g["msg_small_gtab"]
        bits = g["msg_small_message"].astype(int)
        for i in range(md):
            sel, uns = enc.embeddings[2 * i + bits[i]].weight, enc.embeddings[2 * i + 1 - bits[i]].weight
            assert uns.grad is None  # unselected tables stay grad-free, Adam skips them (SURVEY hard part 3)
            np.testing.assert_allclose(sel.grad.cpu().numpy(), gref[2 * i + bits[i]], rtol=1e-5,
                                       atol=1e-6 * np.abs(gref).max())
        # gradients of different tables must not alias (GradScaler.unscale_ works in place)
        ptrs = {enc.embeddings[2 * i + bits[i]].weight.grad.data_ptr() for i in range(md)}
        assert len(ptrs) == md


# ---------------------------------------------------------------------------------------------------------
# The FUSED kernels (k_field_fwd / k_render_rays / k_grid_sweep: the benchmarked path) derive the voxel index with
# one double multiply (hash_common.cuh locate_axis_fused) instead of the reference's fp32 division.  north_star wants
# hash slots bit-exact: compare that device function against the reference-order encoder (pinned to the reference
# module's goldens above) where a one-ulp difference of the quotient would flip the cell.
# ---------------------------------------------------------------------------------------------------------
def _boundary_points(res, ulps):
    """Every cell boundary k/res' (k = 0..res) of one level, stepped by -ulps..+ulps fp32 ulps, for both candidate
    boundary values: the fp32 product k*gs and the fp32 quotient k/res."""
    gs = np.float32(1.0) / np.float32(res)
    k = np.arange(0, int(res) + 2, dtype=np.float32)
    base = np.concatenate([k * gs, (k / np.float32(res)).astype(np.float32)])
    base = np.clip(base, 0.0, 1.0).astype(np.float32)
    pts = [base]
    up, dn = base.copy(), base.copy()
    for _ in range(ulps):
        up = np.nextafter(up, np.float32(2.0)); dn = np.nextafter(dn, np.float32(-1.0))
        pts += [up.copy(), dn.copy()]
    return np.unique(np.clip(np.concatenate(pts), 0.0, 1.0).astype(np.float32))


def test_fused_slots_equal_reference_order_slots():
    from nerf_signature_b200.hash_encoding import HashEmbedder
    enc = HashEmbedder(bounding_box=(0, 1), n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                       base_resolution=16, finest_resolution=2048).cuda()
    # (a) all 16 base resolutions + the message encoder's (2048): every boundary +-40 ulp, placed on each axis in turn
    #     against pseudo-random other coordinates
    rs = np.random.RandomState(0)
    total = 0
    for res in sorted(set(enc.resolutions + [2048.0])):
        b = _boundary_points(res, 40)
        for axis in range(3):
            x = rs.uniform(0, 1, size=(b.size, 3)).astype(np.float32)
            x[:, axis] = b
            xt = torch.from_numpy(x).cuda()
            want = enc.hashed_indices(xt)
            got = enc.fused_hashed_indices(xt)
            assert torch.equal(got, want), f"fused slot mismatch at a cell boundary of resolution {res}"
            total += x.shape[0] * 16
        # the level's own geometry, all three axes on boundaries at once
        x = np.stack([rs.choice(b, b.size), rs.choice(b, b.size), b], axis=1).astype(np.float32)
        xt = torch.from_numpy(x).cuda()
        assert torch.equal(enc.fused_hashed_indices(xt), enc.hashed_indices(xt))
    assert total > 10 ** 7
    # (b) 10^7 random points (uniform, plus a batch that includes values outside [0,1]: the clamp path)
    g = torch.Generator(device="cuda").manual_seed(1)
    for i in range(10):
        xt = torch.rand(10 ** 6, 3, device="cuda", generator=g)
        if i == 9:
            xt = xt * 1.2 - 0.1
        assert torch.equal(enc.fused_hashed_indices(xt), enc.hashed_indices(xt))
    # (c) interpolation weights of the fused path: within 4 fp32 ulp-of-one of the reference expression
    xt = torch.rand(200000, 3, device="cuda", generator=g)
    _, w = enc.fused_hashed_indices(xt, want_weights=True)
    res_t = torch.tensor(enc.resolutions, device="cuda")
    gs = (1.0 / res_t).view(1, -1, 1)
    xx = xt.view(-1, 1, 3)
    vmin = torch.floor(xx / gs) * gs
    w_ref = (xx - vmin) / ((vmin + gs) - vmin)
    assert float((w - w_ref).abs().max()) <= 5e-7 * 2048  # |dw| <= ~2 ulp(x) * res


def test_fused_slots_message_geometry():
    """The message encoder's single resolution through the fused geometry (FieldParams.msg_geom): 2048 exactly
    (base == finest, b = 1), whereas the base encoder's finest level is floor(16 * b**15) = 2047 in torch fp32."""
    from nerf_signature_b200.hash_encoding import HashEmbedder
    enc = HashEmbedder(bounding_box=(0, 1), n_levels=2, n_features_per_level=2, log2_hashmap_size=19,
                       base_resolution=2048, finest_resolution=2048).cuda()
    assert enc.resolutions == [2048.0, 2048.0]
    g = torch.Generator(device="cuda").manual_seed(2)
    xt = torch.rand(500000, 3, device="cuda", generator=g)
    b = torch.from_numpy(_boundary_points(2048.0, 8)).cuda()
    xb = torch.rand(b.numel(), 3, device="cuda", generator=g)
    xb[:, 1] = b
    for x in (xt, xb):
        got = enc.fused_hashed_indices(x, resolutions=[2048.0])
        assert torch.equal(got, enc.hashed_indices(x)[:, :1])
