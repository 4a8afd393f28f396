"""CPU tests (-m "not gpu"): the oracle is only trusted once it reproduces what the REFERENCE produced.

  * oracle/hash_oracle.c and oracle/torch_port.py against tests/golden/hash_golden.npz — outputs of the
    reference's own hash_encoding.py / hash_encoding_wtmk_bit.py modules (tests/golden/make_golden_hash.py);
  * oracle/raymarch_oracle.c against tests/golden/raymarch_golden.npz — outputs of the UNMODIFIED reference
    CUDA extension run on a B200 (tests/golden/make_golden_raymarch.py).
Integer outputs and every float the march emits: bit-exact.  Composite: 2e-5 (the device uses the
ex2.approx-based __expf, the C oracle expf).
"""
import json
import os

import numpy as np
import pytest
import torch

from make_golden_hash import make_tables
import make_golden_raymarch as mgr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---------------------------------------------------------------------------------------------------
# hash encoders
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["small", "full"])
def test_c_hash_oracle_matches_reference_module(golden_hash, oracle_cpu, tag):
    g = golden_hash
    log2_T = int(g[f"base_{tag}_log2T"])
    tabs = make_tables(int(g[f"base_{tag}_seed"]), 16, log2_T)
    feat, slots = oracle_cpu.hash_encode_forward(g[f"base_{tag}_x"], tabs, g[f"base_{tag}_res"], log2_T, want_slots=True)
    assert np.array_equal(slots, g[f"base_{tag}_slots"])                       # hash slots: bit-exact
    assert np.array_equal(bits(feat), bits(g[f"base_{tag}_feat"]))             # features: bit-exact


def test_level_resolutions_are_the_reference_fp32_values(golden_hash):
    # floor(16 * b**i) in torch fp32 ends at 2047, not 2048
    want = [16, 22, 30, 42, 58, 80, 111, 153, 212, 294, 406, 561, 776, 1072, 1482, 2047]
    assert golden_hash["base_full_res"].tolist() == [float(w) for w in want]
    from oracle import torch_port as tp
    assert tp.level_resolutions(16, 2048, 16) == [float(w) for w in want]
    assert tp.level_resolutions(2048, 2048, 64) == [2048.0] * 64


def test_c_hash_oracle_backward_matches_reference_autograd(golden_hash, oracle_cpu):
    g = golden_hash
    grads = oracle_cpu.hash_encode_backward(g["base_small_x"], g["base_small_gout"], g["base_small_res"], 16,
                                            int(g["base_small_log2T"]))
    ref = g["base_small_gtab"]
    got = np.stack(grads)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())
    assert np.array_equal(got == 0, ref == 0)


@pytest.mark.parametrize("tag", ["small", "md32", "md48"])
def test_c_msg_oracle_matches_reference_module(golden_hash, oracle_cpu, tag):
    g = golden_hash
    log2_T, md = int(g[f"msg_{tag}_log2T"]), int(g[f"msg_{tag}_md"])
    tabs = make_tables(int(g[f"msg_{tag}_seed"]), 2 * md, log2_T)
    feat = oracle_cpu.msg_encode_forward(g[f"msg_{tag}_x"], tabs, g[f"msg_{tag}_message"], 2048.0, log2_T)
    want = g[f"msg_{tag}_feat"]
    # the reference sums the per-bit results with torch.sum (pairwise order); the oracle sums in bit order
    np.testing.assert_allclose(feat, want, rtol=1e-5, atol=1e-6 * np.abs(want).max())


def test_c_msg_oracle_backward_matches_reference_autograd(golden_hash, oracle_cpu):
    g = golden_hash
    md, log2_T = int(g["msg_small_md"]), int(g["msg_small_log2T"])
    G = oracle_cpu.msg_encode_backward(g["msg_small_x"], g["msg_small_gout"], 2048.0, log2_T)
    ref = g["msg_small_gtab"]  # [2*md, T, 2]: selected tables carry the gradient, unselected none (zeros)
    msg = g["msg_small_message"].astype(int)
    for i in range(md):
        sel, unsel = ref[2 * i + msg[i]], ref[2 * i + 1 - msg[i]]
        np.testing.assert_allclose(G, sel, rtol=1e-5, atol=1e-6 * np.abs(sel).max())  # every selected == dS
        assert not unsel.any()


def test_torch_port_matches_reference_modules(golden_hash):
    from oracle import torch_port as tp
    g = golden_hash
    tabs = [torch.from_numpy(t) for t in make_tables(int(g["base_small_seed"]), 16, int(g["base_small_log2T"]))]
    feat = tp.hash_embed(torch.from_numpy(g["base_small_x"]), tabs, g["base_small_res"].tolist(), int(g["base_small_log2T"]))
    assert np.array_equal(bits(feat.numpy()), bits(g["base_small_feat"]))
    md, log2_T = int(g["msg_small_md"]), int(g["msg_small_log2T"])
    mt = [torch.from_numpy(t) for t in make_tables(int(g["msg_small_seed"]), 2 * md, log2_T)]
    mf = tp.msg_embed(torch.from_numpy(g["msg_small_x"]), mt, torch.from_numpy(g["msg_small_message"]), 2048.0, log2_T)
    assert np.array_equal(bits(mf.numpy()), bits(g["msg_small_feat"]))
    sh = tp.sh_degree4(torch.from_numpy(g["sh_dirs"]))
    np.testing.assert_allclose(sh.numpy(), g["sh_out"], rtol=1e-6, atol=1e-7)


def test_sh_restatement_matches_reference_shencoder(golden_hash):
    from oracle import field_oracle as fo
    np.testing.assert_allclose(fo.sh4(torch.from_numpy(golden_hash["sh_dirs"])).numpy(), golden_hash["sh_out"],
                               rtol=1e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------------
# raymarching
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="session")
def golden_rm():
    path = os.path.join(ROOT, "tests", "golden", "raymarch_golden.npz")
    if not os.path.exists(path):
        pytest.fail("tests/golden/raymarch_golden.npz is missing: regenerate it on a GPU box with "
                    "tests/golden/make_golden_raymarch.py")
    return np.load(path)


@pytest.mark.parametrize("case", mgr.CASES, ids=[c[0] for c in mgr.CASES])
def test_c_raymarch_oracle_matches_reference_cuda(golden_rm, oracle_cpu, case):
    g, name = golden_rm, case[0]
    assert json.loads(str(g[f"{name}_params"])) == list(case)
    _, cam, N, seed, bound, C, kind, dt_gamma, perturb = case
    rays_o, rays_d, bitfield, aabb, noises = mgr.case_inputs(case)
    nears, fars = oracle_cpu.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    assert np.array_equal(bits(nears), bits(g[f"{name}_nears"])) and np.array_equal(bits(fars), bits(g[f"{name}_fars"]))
    xyzs, dirs, deltas, rays, counter = oracle_cpu.march_rays_train(rays_o, rays_d, bound, bitfield, C, 128, nears, fars,
                                                                    noises=noises, dt_gamma=dt_gamma)
    assert counter.tolist() == [int(g[f"{name}_total"]), N]
    assert np.array_equal(rays[:, 0], np.arange(N))                            # the oracle's own order is ray order
    assert np.array_equal(rays[:, 2], g[f"{name}_counts"])                     # per-ray sample counts: bit-exact
    m = int(counter[0])
    for got, key in ((xyzs, "xyzs"), (dirs, "dirs"), (deltas, "deltas")):
        assert np.array_equal(bits(got[:m]), bits(g[f"{name}_{key}"])), key    # every emitted float: bit-exact

    sig, rgb = mgr.seeded_field(m, seed + 11)
    ws, depth, image = oracle_cpu.composite_rays_train_forward(sig, rgb, deltas[:m], rays, 1e-4)
    np.testing.assert_allclose(ws, g[f"{name}_ws"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(depth, g[f"{name}_depth"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(image, g[f"{name}_image"], rtol=2e-5, atol=2e-6)
    rs = np.random.RandomState(seed + 12)
    gws = rs.normal(size=N).astype(np.float32); gim = rs.normal(size=(N, 3)).astype(np.float32)
    gs, gc = oracle_cpu.composite_rays_train_backward(gws, gim, sig, rgb, deltas[:m], rays, ws, image, 1e-4)
    scale = np.abs(g[f"{name}_gsig"]).max()
    np.testing.assert_allclose(gs, g[f"{name}_gsig"], rtol=1e-3, atol=1e-5 * scale)
    np.testing.assert_allclose(gc, g[f"{name}_grgb"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("case", mgr.CASES, ids=[c[0] for c in mgr.CASES])
def test_c_inference_oracle_matches_reference_cuda_loop(golden_rm, oracle_cpu, case):
    g, name = golden_rm, case[0]
    _, cam, N, seed, bound, C, kind, dt_gamma, perturb = case
    rays_o, rays_d, bitfield, aabb, _ = mgr.case_inputs(case)
    nears, fars = oracle_cpu.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    ws = np.zeros(N, np.float32); dp = np.zeros(N, np.float32); im = np.zeros((N, 3), np.float32)
    alive = np.arange(N, dtype=np.int32); rt = nears.copy()
    step = iters = 0
    while step < 1024:
        n_alive = alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        x, d, l = oracle_cpu.march_rays(n_alive, n_step, alive, rt, rays_o, rays_d, bound, bitfield, C, 128, nears, fars,
                                        align=128, dt_gamma=dt_gamma)
        s, c = mgr.closed_form_field(x)
        oracle_cpu.composite_rays(n_alive, n_step, alive, rt, s, c, l, ws, dp, im, 1e-4)
        alive = alive[alive >= 0]
        step += n_step; iters += 1
    assert iters == int(g[f"{name}_inf_iters"])   # same alive-ray schedule as the reference loop
    np.testing.assert_allclose(ws, g[f"{name}_inf_ws"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(im, g[f"{name}_inf_image"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dp, g[f"{name}_inf_depth"], rtol=1e-4, atol=1e-4)


def test_c_morton_packbits_match_reference_cuda(golden_rm, oracle_cpu):
    g = golden_rm
    assert np.array_equal(oracle_cpu.morton3D(g["morton_coords"]), g["morton_idx"])
    assert np.array_equal(oracle_cpu.morton3D_invert(g["morton_idx"]), g["morton_back"])
    assert np.array_equal(g["morton_back"], g["morton_coords"])
    assert np.array_equal(oracle_cpu.packbits(g["packbits_grid"], 0.25), g["packbits_bits"])


# ---------------------------------------------------------------------------------------------------
# non-cuda_ray renderer: oracle/torch_port.render_run against the reference's own NeRFRenderer.run
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def golden_run():
    return np.load(os.path.join(ROOT, "tests", "golden", "run_golden.npz"))


@pytest.mark.parametrize("case", ["clean", "wtmk", "wtmk_bound2", "wtmk_upsample"])
def test_torch_port_render_run_matches_reference_run(golden_run, case):
    """tests/golden/run_golden.npz holds what the reference's unmodified NeRFRenderer.run (+ sample_pdf) returned on CPU
    for these rays (tests/golden/make_golden_run.py).  The port is bench.py's CPU baseline and the oracle of the
    product's `run`: image, weights_sum and depth within fp32 rounding of the reference's."""
    import make_golden_run as mg
    from oracle import torch_port as tp
    field, o, d, msg, steps, up = mg.case_inputs(case)
    with torch.no_grad():
        image, ws, depth = tp.render_run(field, o, d, msg, num_steps=steps, upsample_steps=up, return_depth=True)
    g = golden_run
    assert float(g[f"{case}_weights_sum"].max()) > 0.5 and np.ptp(g[f"{case}_image"]) > 0.1      # a visible scene
    np.testing.assert_allclose(ws.numpy(), g[f"{case}_weights_sum"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(image.numpy(), g[f"{case}_image"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(depth.numpy(), g[f"{case}_depth"], rtol=0, atol=2e-6)


def test_product_sample_pdf_matches_reference_sample_pdf(golden_run):
    """nerf_signature_b200.nerf.renderer_wtmk.sample_pdf is plain torch (no kernel) and therefore checked here, on CPU,
    against the reference's own function (renderer_wtmk.py:12-46): deterministic and seeded-random draws, incl. an all-zero
    ray, a ray with all its weight in one bin and one with weight in the outermost bins only."""
    from nerf_signature_b200.nerf.renderer_wtmk import sample_pdf
    from oracle import torch_port as tp
    g = golden_run
    bins, weights = torch.from_numpy(g["pdf_bins"]), torch.from_numpy(g["pdf_weights"])
    det = sample_pdf(bins, weights, 24, det=True)
    assert np.array_equal(det.numpy().view(np.uint32), g["pdf_det"].view(np.uint32))
    torch.manual_seed(1234)
    rnd = sample_pdf(bins, weights, 24, det=False)
    assert np.array_equal(rnd.numpy().view(np.uint32), g["pdf_rand_seed1234"].view(np.uint32))
    assert np.array_equal(tp.resample_depths(bins, weights, 24).numpy().view(np.uint32), g["pdf_det"].view(np.uint32))
    # samples stay inside their ray's bin range and are sorted for the deterministic draw
    assert (det >= bins[:, :1]).all() and (det <= bins[:, -1:]).all() and (det[:, 1:] >= det[:, :-1]).all()


# ---------------------------------------------------------------------------------------------------
# one watermark training step: oracle/torch_port.train_step against the reference's own Trainer.train_step
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["md4", "md8_lw1"])
def test_torch_port_train_step_matches_reference_train_step(case):
    """tests/golden/trainstep_golden.npz holds what the reference's unmodified Trainer.train_step body
    (utils_wtmk_disen.py:579-646; tests/golden/make_golden_trainstep.py) computed on CPU: the three losses, the clamped
    block pixels handed to the decoder, and the gradients autograd delivered to the message tables and the decoder.  The
    port (bench.py's CPU baseline step; the semantics harness.Scene and nerf.loss_ops follow) reproduces them, and only the
    tables the message selects receive a gradient."""
    import make_golden_trainstep as mg
    from nerf_signature_b200.nerf.hidden_models import get_hidden_decoder_multi_views
    from oracle import torch_port as tp
    g = np.load(os.path.join(ROOT, "tests", "golden", "trainstep_golden.npz"))
    field, batch, message, lw, li, seed = mg.case_inputs(case)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        torch.manual_seed(seed)
        decoder = get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=8, input_ch=3, channels=64)
        loss, lossi, lossw, pred = tp.train_step(field, decoder, batch, message, lambda_w=lw, lambda_i=li,
                                                 num_steps=mg.NUM_STEPS, return_terms=True)
    finally:
        torch.set_num_threads(threads)
    got = mg.summarize(field, decoder, (loss, lossi, lossw))
    np.testing.assert_allclose(got["losses"], g[f"{case}_losses"], rtol=1e-5)
    assert abs(loss - (lw * lossw + li * lossi)) < 1e-6 * max(1.0, abs(loss))
    np.testing.assert_allclose(pred.numpy(), g[f"{case}_pred_rgb"], rtol=0, atol=2e-6)
    want_sums = g[f"{case}_table_grad_sums"]
    selected = np.zeros(len(field.msg_tables), bool)
    selected[[2 * i + int(b) for i, b in enumerate(message.tolist())]] = True
    assert np.array_equal(want_sums > 0, selected) and np.array_equal(got["table_grad_sums"] > 0, selected)
    np.testing.assert_allclose(got["table_grad_sums"], want_sums, rtol=1e-3)
    ref0 = g[f"{case}_table_grad_0"]
    assert np.abs(got["table_grad_0"] - ref0).max() <= 1e-3 * np.abs(ref0).max()
    np.testing.assert_allclose(got["decoder_grad_norms"], g[f"{case}_decoder_grad_norms"], rtol=1e-3)
