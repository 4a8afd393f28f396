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
