"""-m gpu: the fused frame renderer (nsig_render_rays, one persistent kernel) against the reference-shaped
alive-ray loop (march_rays -> forward -> composite_rays with compaction, renderer_wtmk.py:323-372) that
tests/test_raymarching_gpu.py pins to the oracle kernel by kernel.  Tolerance 1e-3 (north_star), observed ~1e-5:
the loop re-derives the march parameter through composite_rays' `t += deltas[1]`, the fused kernel keeps the
exact lattice; both use fp16 MMA operands with fp32 accumulation."""
import numpy as np
import pytest
import torch

from nerf_signature_b200 import synthetic as syn
from conftest import record_parity

pytestmark = pytest.mark.gpu


def _net(bound, md, table_scale=300.0, seed=0, sigma_gain=0.0):
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork
    torch.manual_seed(seed)
    net = NeRFNetwork(bound=bound, cuda_ray=True, message_dim=md)
    with torch.no_grad():
        for e in list(net.encoder.embeddings) + list(net.msg_encoder.embeddings):
            e.weight.mul_(table_scale)
        if sigma_gain:
            # bias the density logit upwards (through the weights of the sigma output row) so that rays terminate early
            net.sigma_net.params[2048:2048 + 64] += sigma_gain
    net = net.cuda().eval()
    grid = syn.sphere_grid(net.cascade)
    net.density_grid.copy_(torch.from_numpy(grid))
    net.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, 0.5)))
    return net


@pytest.mark.parametrize("bound,rays_fn,dt_gamma,sigma_gain", [(1.0, syn.blender_rays, 0.0, 0.0),
                                                               (1.0, syn.blender_rays, 0.0, 0.6),
                                                               (2.0, syn.rays_360, 1.0 / 128, 0.3)])
def test_fused_renderer_matches_reference_loop(bound, rays_fn, dt_gamma, sigma_gain):
    md = 8
    net = _net(bound, md, sigma_gain=sigma_gain)
    rays_o, rays_d = rays_fn(1500, seed=11)
    ro, rd = torch.from_numpy(rays_o)[None].cuda(), torch.from_numpy(rays_d)[None].cuda()
    msg = torch.from_numpy(np.random.RandomState(2).randint(0, 2, size=md).astype(np.float32)).cuda()
    kw = dict(staged=False, bg_color=1, perturb=False, dt_gamma=dt_gamma, max_steps=1024, T_thresh=1e-4)
    with torch.no_grad():
        net.fused_inference = True
        a = net.render(ro, rd, msg, **kw)
        n_samples = int(net.last_render_samples)
        net.fused_inference = False
        b = net.render(ro, rd, msg, **kw)
    assert n_samples > 1500
    img_a, img_b = a["image"].cpu().numpy(), b["image"].cpu().numpy()
    assert np.abs(img_b - 1.0).max() > 0.05        # the scene is actually visible (not all background)
    np.testing.assert_allclose(img_a, img_b, rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(a["depth"].cpu().numpy(), b["depth"].cpu().numpy(), rtol=1e-3, atol=2e-4)
    # staged rendering in chunks gives the same frame
    with torch.no_grad():
        net.fused_inference = True
        c = net.render(ro, rd, msg, staged=True, max_ray_batch=400, bg_color=1, perturb=False, dt_gamma=dt_gamma,
                       max_steps=1024, T_thresh=1e-4)
    np.testing.assert_allclose(c["image"].cpu().numpy(), img_a, rtol=1e-6, atol=1e-7)


def _oracle_frame(oracle_cpu, net, rays_o, rays_d, msg, bound, T_thresh):
    """The frame the oracle gives for these rays: C oracle march + hash encoders, torch-fp32 MLP oracle, and the INFERENCE
    composite rule restated in fp64 - composite_rays accumulates a sample and THEN stops if the transmittance BEFORE it was
    below T_thresh (raymarching.cu:868-883), one sample later than composite_rays_train (raymarching.cu:548-551).  With
    perturb off the inference march (raymarching.cu:701-800) visits the lattice points of the training march
    (raymarching.cu:312-480), so the training march supplies the samples.  Returns (image on white, weights_sum, samples
    marched, samples consumed, image under the TRAINING rule by the C oracle)."""
    from oracle import field_oracle as fo
    N = rays_o.shape[0]
    bitfield = net.density_bitfield.cpu().numpy()
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    on, of = oracle_cpu.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    xyz, dirs, deltas, rays, cnt = oracle_cpu.march_rays_train(rays_o, rays_d, bound, bitfield, 1, 128, on, of)
    m = int(cnt[0])
    xn = ((xyz[:m] + np.float32(bound)) * np.float32(0.5 / bound)).astype(np.float32)
    tabs = [e.weight.detach().cpu().numpy() for e in net.encoder.embeddings]
    feat = oracle_cpu.hash_encode_forward(xn, tabs, net.encoder.resolutions, 19)
    mt = [e.weight.detach().cpu().numpy() for e in net.msg_encoder.embeddings]
    feat[:, 30:32] += oracle_cpu.msg_encode_forward(xn, mt, msg, net.msg_encoder.resolution, 19)
    sig, rgb, _, _ = fo.mlp_forward(torch.from_numpy(feat), torch.from_numpy(dirs[:m]),
                                    net.sigma_net.params.detach().cpu(), net.color_net.params.detach().cpu())
    sig, rgb = sig.numpy(), rgb.numpy()
    ws_t, _, img_t = oracle_cpu.composite_rays_train_forward(sig, rgb, deltas[:m], rays, T_thresh)
    alpha = 1.0 - np.exp(-sig.astype(np.float64) * deltas[:m, 0].astype(np.float64))
    img, ws, used = np.zeros((N, 3)), np.zeros(N), 0
    for n in range(N):
        lo, c = int(rays[n, 1]), int(rays[n, 2])
        for k in range(lo, lo + c):
            T = 1.0 - ws[n]
            w = alpha[k] * T
            ws[n] += w
            img[n] += w * rgb[k]
            used += 1
            if T < T_thresh:
                break
    return (img + (1 - ws)[:, None]).astype(np.float32), ws.astype(np.float32), m, used, img_t + (1 - ws_t)[:, None]


@pytest.mark.parametrize("sigma_gain,T_thresh", [(0.0, 1e-4), (0.6, 1e-4), (12.0, 1e-2)])
def test_fused_renderer_matches_cpu_oracle(oracle_cpu, sigma_gain, T_thresh):
    """nsig_render_rays DIRECTLY against the oracle (not through the repo's own alive-ray loop).  Cases 1-2: no ray
    reaches T_thresh, every marched sample is consumed, 1e-3 (north_star; observed 2e-6).  Case 3: a dense scene in which
    rays are killed after ~17 % of their samples.  Measured on a B200 against the TRAINING-rule composite the kernel's
    frame was at most 3.3613e-3 away; the oracle's own inference-rule and training-rule frames differ by at most
    3.3610e-3 (the one extra sample the inference rule accumulates), i.e. the kernel follows the inference rule.  The
    round's GPU budget ended before the direct comparison with the inference-rule frame could be re-run, so the bound
    asserted for case 3 is the one that FOLLOWS from that measurement (3.4e-3 + one sample's weight < T_thresh), not
    the tight one; conftest.record_parity logs the observed error."""
    md, N, bound = 8, 600, 1.0
    net = _net(bound, md, sigma_gain=sigma_gain)
    rays_o, rays_d = syn.blender_rays(N, seed=23)
    msg = np.random.RandomState(4).randint(0, 2, size=md).astype(np.float32)
    with torch.no_grad():
        net.fused_inference = True
        out = net.render(torch.from_numpy(rays_o)[None].cuda(), torch.from_numpy(rays_d)[None].cuda(),
                         torch.from_numpy(msg).cuda(), staged=False, bg_color=1, perturb=False, dt_gamma=0.0,
                         max_steps=1024, T_thresh=T_thresh)
    n_samples = int(net.last_render_samples)
    image = out["image"].cpu().numpy().reshape(-1, 3)
    img, ws, m, used, img_train_rule = _oracle_frame(oracle_cpu, net, rays_o, rays_d, msg, bound, T_thresh)
    assert np.abs(img - 1.0).max() > 0.05           # the scene is visible
    dense = sigma_gain > 10
    if dense:
        assert used < 0.25 * m and used <= n_samples < 0.8 * m     # killed rays; the kernel evaluates 32 samples at a time
        assert np.abs(img - img_train_rule).max() < T_thresh        # the two rules differ by one sample's weight
    else:
        assert used == m == n_samples                               # nothing terminates: every marched sample is consumed
        assert np.abs(img - img_train_rule).max() < 1e-5
    per_ray = np.abs(image - img).max(axis=1)
    err = float(per_ray.max())
    record_parity("render_rays_vs_cpu_oracle", {"sigma_gain": sigma_gain, "T_thresh": T_thresh, "image_max_abs": err,
                                                "samples": n_samples, "oracle_samples": m, "oracle_samples_consumed": used})
    assert err < (1e-3 + 1.5 * T_thresh if dense else 1e-3), err
    # ... and against the frame the REFERENCE'S OWN run_cuda loop returned for this network and these rays
    # (tests/golden/runcuda_golden.npz; tests/test_runcuda_oracle_cpu.py shows it equals the oracle frame above to 5e-6)
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runcuda_golden.npz"))
    key = {0.0: "gain0", 0.6: "gain0.6", 12.0: "gain12_dense"}[sigma_gain]
    err_ref = float(np.abs(image - gold[f"{key}_eval_image"]).max())
    assert err_ref < (1e-3 + 1.5 * T_thresh if dense else 1e-3), err_ref
    if "weights_sum" in out and not dense:
        assert np.abs(out["weights_sum"].cpu().numpy().reshape(-1) - ws).max() < 1e-3


def test_fused_renderer_early_termination_saves_samples():
    md = 4
    rays_o, rays_d = syn.blender_rays(2000, seed=5)
    ro, rd = torch.from_numpy(rays_o)[None].cuda(), torch.from_numpy(rays_d)[None].cuda()
    msg = torch.zeros(md).cuda()
    counts = []
    for gain in (0.0, 12.0):
        net = _net(1.0, md, sigma_gain=gain)
        with torch.no_grad():
            net.render(ro, rd, msg, staged=False, bg_color=1, perturb=False, T_thresh=1e-2)
        counts.append(int(net.last_render_samples))
    assert counts[1] < 0.7 * counts[0], counts


def test_fused_renderer_clean_model_and_edge_cases():
    from nerf_signature_b200.nerf.network_hash import NeRFNetwork
    torch.manual_seed(0)
    net = NeRFNetwork(bound=1, cuda_ray=True).cuda().eval()
    grid = syn.sphere_grid(1)
    net.density_grid.copy_(torch.from_numpy(grid))
    net.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, 0.5)))
    rays_o, rays_d = syn.blender_rays(300, seed=1)
    rays_o[:5] = np.array([5, 5, 5], np.float32)      # rays that miss the box
    rays_d[:5] = np.array([0, 0, 1], np.float32)
    ro, rd = torch.from_numpy(rays_o)[None].cuda(), torch.from_numpy(rays_d)[None].cuda()
    with torch.no_grad():
        a = net.render(ro, rd, staged=False, bg_color=1, perturb=False)
        net.fused_inference = False
        b = net.render(ro, rd, staged=False, bg_color=1, perturb=False)
    np.testing.assert_allclose(a["image"].cpu().numpy(), b["image"].cpu().numpy(), rtol=1e-3, atol=1e-4)
    assert torch.all(a["image"][0, :5] == 1.0)        # pure background
    # empty grid: every ray is background, zero samples
    net.density_bitfield.zero_()
    net.fused_inference = True
    with torch.no_grad():
        e = net.render(ro, rd, staged=False, bg_color=1, perturb=False)
    assert torch.all(e["image"] == 1.0) and int(net.last_render_samples) == 0


# ---------------------------------------------------------------------------------------------------------
# non-cuda_ray renderer (NeRFRenderer.run): BASELINE configs[0] shape against the CPU port of the reference's run()
# ---------------------------------------------------------------------------------------------------------
def _load_port_weights(net, field):
    with torch.no_grad():
        for e, t in zip(net.encoder.embeddings, field.base_tables):
            e.weight.copy_(t)
        net.sigma_net.params.copy_(torch.cat([w.reshape(-1) for w in field.Ws]))
        net.color_net.params.copy_(torch.cat([w.reshape(-1) for w in field.Wc]))


def test_run_non_cuda_ray_matches_reference_port():
    """configs[0]: random-init clean HashNeRF, non-cuda_ray render with 512 uniform samples per ray, against
    oracle/torch_port.render_run (the reference's NeRFRenderer.run, upsample_steps=0) on the CPU: 1e-3."""
    from nerf_signature_b200.nerf.network_hash import NeRFNetwork
    from oracle import torch_port as tp
    field = tp.PortField(bound=1.0, message_dim=0, seed=3, train_msg=False)
    for t in field.base_tables:
        t.mul_(300.0)     # features O(0.03): the MLPs matter
    net = NeRFNetwork(bound=1.0, cuda_ray=False)
    _load_port_weights(net, field)
    net = net.cuda().eval()
    o, d = syn.blender_rays(192, seed=5)
    with torch.no_grad():
        want, want_ws = tp.render_run(field, torch.from_numpy(o), torch.from_numpy(d), None, num_steps=512)
        got = net.render(torch.from_numpy(o)[None].cuda(), torch.from_numpy(d)[None].cuda(), staged=False, num_steps=512,
                         upsample_steps=0, bg_color=1, perturb=False)
    np.testing.assert_allclose(got["image"][0].cpu().numpy(), want.numpy(), rtol=0, atol=1e-3)
    np.testing.assert_allclose(got["weights_sum"].cpu().numpy(), want_ws.numpy(), rtol=0, atol=1e-3)
    # staged == unstaged, importance resampling runs and keeps the image close (same field, more samples)
    with torch.no_grad():
        st = net.render(torch.from_numpy(o)[None].cuda(), torch.from_numpy(d)[None].cuda(), staged=True, max_ray_batch=50,
                        num_steps=512, upsample_steps=0, bg_color=1, perturb=False)
        up = net.render(torch.from_numpy(o)[None].cuda(), torch.from_numpy(d)[None].cuda(), staged=False, num_steps=256,
                        upsample_steps=128, bg_color=1, perturb=False)
    assert torch.allclose(st["image"], got["image"], atol=1e-6)
    assert up["image"].shape == got["image"].shape and float((up["image"] - got["image"]).abs().max()) < 0.05


def test_run_non_cuda_ray_is_differentiable_in_watermark_training():
    """with cuda_ray=False the reference trains through run(); the message tables must receive a gradient."""
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork
    torch.manual_seed(0)
    net = NeRFNetwork(bound=1.0, cuda_ray=False, message_dim=4).cuda().train()
    with torch.no_grad():
        for e in list(net.encoder.embeddings) + list(net.msg_encoder.embeddings):
            e.weight.mul_(300.0)
    o, d = syn.blender_rays(64, seed=6)
    msg = torch.tensor([1.0, 0.0, 1.0, 1.0]).cuda()
    out = net.render(torch.from_numpy(o)[None].cuda(), torch.from_numpy(d)[None].cuda(), msg, staged=False, num_steps=64,
                     upsample_steps=32, bg_color=1, perturb=True)
    out["image"].square().mean().backward()
    bits = [1, 0, 1, 1]
    for i, b in enumerate(bits):
        g = net.msg_encoder.embeddings[2 * i + b].weight.grad
        assert g is not None and float(g.abs().sum()) > 0
        assert net.msg_encoder.embeddings[2 * i + 1 - b].weight.grad is None
