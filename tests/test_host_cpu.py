"""CPU tests (-m "not gpu") of the host-side mirror of the reference interface: module structure,
state-dict keys, freezing policy, level resolutions, synthetic batch shapes.  No kernels run."""
import inspect
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_raymarching_module_has_the_reference_callables_and_signatures():
    from nerf_signature_b200 import raymarching as rm
    # reference raymarching/raymarching.py:22,55,85,108,132,164,241,300,354
    want = {
        "_near_far_from_aabb": ["ctx", "rays_o", "rays_d", "aabb", "min_near"],
        "_sph_from_ray": ["ctx", "rays_o", "rays_d", "radius"],
        "_morton3D": ["ctx", "coords"],
        "_morton3D_invert": ["ctx", "indices"],
        "_packbits": ["ctx", "grid", "thresh", "bitfield"],
        "_march_rays_train": ["ctx", "rays_o", "rays_d", "bound", "density_bitfield", "C", "H", "nears", "fars", "step_counter",
                              "mean_count", "perturb", "align", "force_all_rays", "dt_gamma", "max_steps"],
        "_composite_rays_train": ["ctx", "sigmas", "rgbs", "deltas", "rays", "T_thresh"],
        "_march_rays": ["ctx", "n_alive", "n_step", "rays_alive", "rays_t", "rays_o", "rays_d", "bound", "density_bitfield", "C",
                        "H", "near", "far", "align", "perturb", "dt_gamma", "max_steps"],
        "_composite_rays": ["ctx", "n_alive", "n_step", "rays_alive", "rays_t", "sigmas", "rgbs", "deltas", "weights_sum", "depth",
                            "image", "T_thresh"],
    }
    mod = rm.raymarching
    for cls, params in want.items():
        fwd = getattr(mod, cls).forward
        fwd = inspect.unwrap(fwd)
        assert list(inspect.signature(fwd).parameters) == params, cls
        assert callable(getattr(rm, cls[1:]))
    d = inspect.signature(inspect.unwrap(mod._march_rays_train.forward)).parameters
    assert (d["mean_count"].default, d["align"].default, d["max_steps"].default, d["dt_gamma"].default) == (-1, -1, 1024, 0)
    assert inspect.signature(inspect.unwrap(mod._composite_rays.forward)).parameters["T_thresh"].default == 1e-2
    assert inspect.signature(inspect.unwrap(mod._composite_rays_train.forward)).parameters["T_thresh"].default == 1e-4


def test_watermark_network_state_dict_and_freezing_policy():
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork
    net = NeRFNetwork(bound=2, cuda_ray=True, message_dim=4)
    keys = set(net.state_dict().keys())
    for i in range(16):
        assert f"encoder.embeddings.{i}.weight" in keys
    for i in range(8):
        assert f"msg_encoder.embeddings.{i}.weight" in keys
    for k in ("sigma_net.params", "color_net.params", "density_grid", "density_bitfield", "step_counter", "aabb_train",
              "aabb_infer"):
        assert k in keys, k
    assert any(k.startswith("msg_decoder.") for k in keys)
    assert net.cascade == 2 and net.density_bitfield.numel() == 2 * 128 ** 3 // 8 and net.density_grid.shape == (2, 128 ** 3)
    assert net.sigma_net.params.numel() == 3072 and net.color_net.params.numel() == 7168
    # network_wtmk_tcnn.py:90-95: base encoder + both MLPs frozen, message tables + decoder trainable
    assert not any(p.requires_grad for p in net.encoder.parameters())
    assert not net.sigma_net.params.requires_grad and not net.color_net.params.requires_grad
    assert all(p.requires_grad for p in net.msg_encoder.parameters())
    assert all(p.requires_grad for p in net.msg_decoder.parameters())
    groups = net.get_params(1e-2)
    assert len(groups) == 2 and all(g["lr"] == 1e-2 for g in groups)
    assert NeRFNetwork(bound=1, cuda_ray=True, message_dim=4, finetune_decoder=True).get_params(1e-3).__len__() == 1
    # tables initialised U(-1e-4, 1e-4) (hash_encoding.py:65-66)
    w = net.encoder.embeddings[3].weight
    assert w.shape == (2 ** 19, 2) and float(w.abs().max()) <= 1e-4


def test_clean_network_trains_everything():
    from nerf_signature_b200.nerf.network_hash import NeRFNetwork
    net = NeRFNetwork(bound=1, cuda_ray=True)
    assert all(p.requires_grad for p in net.parameters())
    assert len(net.get_params(1e-2)) >= 3


def test_encoder_resolutions_follow_the_reference_expression():
    from nerf_signature_b200.hash_encoding import HashEmbedder
    from nerf_signature_b200.hash_encoding_wtmk_bit import HashEmbedder as MsgEmbedder, message_bits
    enc = HashEmbedder(bounding_box=(0, 1), n_levels=16, log2_hashmap_size=8, base_resolution=16, finest_resolution=2048)
    assert enc.resolutions == [16, 22, 30, 42, 58, 80, 111, 153, 212, 294, 406, 561, 776, 1072, 1482, 2047]
    m = MsgEmbedder(bounding_box=(0, 1), n_levels=8, log2_hashmap_size=8, base_resolution=2048, finest_resolution=2048,
                    message_dim=4)
    assert m.resolution == 2048.0 and len(m.tables()) == 8
    assert message_bits(torch.tensor([1.0, 0.0, 1.0, 1.0])) == (1, 0, 1, 1)


def test_synthetic_batch_shapes_match_the_reference_provider():
    from nerf_signature_b200 import harness
    cfg = harness.CONFIGS["blender_wtmk"]
    b = harness.make_batch(cfg, seed=0)
    # provider_wtmk.py:481-496: [message_dim, pH, pW, 3] blocks of a 400x400 frame cut 32x32; 4096 content rays
    assert b["rays_o_block"].shape == (32, 12, 12, 3) and b["rays_d_block"].shape == (32, 12, 12, 3)
    assert b["rays_o"].shape == (1, 4096, 3) and b["gt"].shape == (1, 4096, 3)
    np.testing.assert_allclose(np.linalg.norm(b["rays_d"], axis=-1), 1.0, atol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(b["rays_o"], axis=-1), 4.0311 * 0.8, atol=1e-4)
    b2 = harness.make_batch(cfg, seed=0)
    assert all(np.array_equal(b[k], b2[k]) for k in b)  # seeded


def test_workloads_carry_the_parameters_baseline_json_names():
    """harness.CONFIGS vs the prose of BASELINE.json configs[1], [2], [4] (the numbers the bench lines are quoted on)."""
    import json
    from nerf_signature_b200 import harness
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    c1, c2, c4 = harness.CONFIGS["blender_wtmk"], harness.CONFIGS["360_wtmk"], harness.CONFIGS["shard262144_wtmk"]
    assert "bound 1.0, scale 0.8, dt_gamma 0" in cfgs[1] and "message_dim 32, 32x32 codebook, num_rays 4096" in cfgs[1]
    assert (c1["bound"], c1["scale"], c1["dt_gamma"], c1["message_dim"], c1["num_rows"], c1["num_cols"], c1["num_rays"]) == \
        (1.0, 0.8, 0.0, 32, 32, 32, 4096)
    assert "scale 0.33, dt_gamma 0" in cfgs[2] and "num_rays 4096, message_dim 32" in cfgs[2] and "every 16 iters" in cfgs[2]
    assert (c2["scale"], c2["dt_gamma"], c2["message_dim"], c2["num_rays"], c2["grid_update_every"]) == (0.33, 0.0, 32, 4096, 16)
    assert c2["bound"] > 1.0                                   # unbounded scene: more than one cascade
    assert "262144 rays/step" in cfgs[4] and "message_dim 48" in cfgs[4]
    assert (c4["num_rays"], c4["message_dim"]) == (262144, 48)
    # configs[1]'s step: 4096 content rays + message_dim blocks of (400/32)^2 pixels = 8704 rays
    assert c1["num_rays"] + c1["message_dim"] * (c1["H"] // c1["num_rows"]) * (c1["W"] // c1["num_cols"]) == 8704


def test_shard_batch_partitions_one_global_batch_exactly():
    """the ranks' shards of configs[4]-shaped batches are contiguous, disjoint, complete and balanced,
    for the content rays (with their ground truth) and for the flattened watermark-block rays."""
    from nerf_signature_b200 import harness
    cfg = dict(harness.CONFIGS["shard262144_wtmk"], num_rays=1000 + 3)        # ragged on purpose
    g = harness.make_batch(cfg, seed=3)
    md, pH, pW = g["rays_o_block"].shape[:3]
    assert (md, pH, pW) == (48, 12, 12)
    for world in (1, 2, 3, 4, 8):
        shards = [harness.shard_batch(g, r, world) for r in range(world)]
        assert all(s[1] == (md, pH, pW) for s in shards)
        counts = shards[0][2]
        assert all(s[2] == counts for s in shards) and sum(counts) == md * pH * pW
        assert [s[0]["rays_o_block"].shape[0] for s in shards] == counts and max(counts) - min(counts) <= 1
        for key, full in (("rays_o_block", g["rays_o_block"].reshape(-1, 3)), ("rays_d_block", g["rays_d_block"].reshape(-1, 3))):
            assert np.array_equal(np.concatenate([s[0][key] for s in shards], axis=0), full)
        for key in ("rays_o", "rays_d", "gt"):
            assert np.array_equal(np.concatenate([s[0][key] for s in shards], axis=1), g[key])
        sizes = [s[0]["rays_o"].shape[1] for s in shards]
        assert max(sizes) - min(sizes) <= 1 and all(s[0]["gt"].shape == s[0]["rays_o"].shape for s in shards)
        assert all(a.flags["C_CONTIGUOUS"] for s in shards for a in s[0].values())   # handed to pinned copies as they are


def test_fused_index_identity_double_multiply_equals_fp32_division():
    """csrc/hash_common.cuh locate_axis_fused obtains fl32(x / g) as fl32(fl64(x) * fl64(1/g)).  The identity is
    checked here in numpy (IEEE fp32 division vs the double-multiply form) at every cell boundary of all 17
    resolutions +-40 ulp and on 2 M random points; the device function itself is checked in tests/test_hash_gpu.py."""
    from nerf_signature_b200.hash_encoding import level_resolutions
    base, finest = torch.tensor(16), torch.tensor(2048)
    b = torch.exp((torch.log(finest) - torch.log(base)) / 15)
    resolutions = sorted(set(level_resolutions(base, b, 16) + [2048.0]))
    assert resolutions[0] == 16.0 and resolutions[-1] == 2048.0
    rs = np.random.RandomState(0)
    rnd = rs.uniform(0, 1, size=2_000_000).astype(np.float32)
    n = 0
    for res in resolutions:
        gs = np.float32(1.0) / np.float32(res)
        k = np.arange(0, int(res) + 2, dtype=np.float32)
        pts = [np.clip(k * gs, 0, 1).astype(np.float32), np.clip(k / np.float32(res), 0, 1).astype(np.float32)]
        for seed in list(pts):
            up, dn = seed.copy(), seed.copy()
            for _ in range(40):
                up = np.nextafter(up, np.float32(2.0)); dn = np.nextafter(dn, np.float32(-1.0))
                pts += [up, dn]
        x = np.clip(np.concatenate(pts + [rnd]), 0.0, 1.0).astype(np.float32)
        q_ref = x / gs                                                    # IEEE fp32 division (hash_encoding.py:39)
        q_fused = (x.astype(np.float64) * (1.0 / np.float64(gs))).astype(np.float32)
        assert np.array_equal(q_ref.view(np.uint32), q_fused.view(np.uint32)), res
        n += x.size
    assert n > 3 * 10 ** 7


def test_trunc_exp_forward_is_exp_and_gradient_is_clamped():
    """activation.py:5-18 of the reference: fp32 exp forward, gradient g * exp(clamp(x, -15, 15)) - a pre-activation
    beyond +-15 keeps a bounded gradient; half inputs are computed in fp32."""
    from nerf_signature_b200.activation import trunc_exp
    x = torch.tensor([-20.0, -15.0, -1.0, 0.0, 2.5, 15.0, 20.0], requires_grad=True)
    y = trunc_exp(x)
    assert torch.equal(y, torch.exp(x.detach()))
    g = torch.tensor([1.0, 2.0, -1.0, 0.5, 1.0, 1.0, 3.0])
    y.backward(g)
    assert torch.equal(x.grad, g * torch.exp(x.detach().clamp(-15, 15)))
    assert torch.equal(x.grad[-1], torch.tensor(3.0) * torch.exp(torch.tensor(15.0)))   # exp(15), not exp(20)
    h = torch.tensor([0.5, 3.0], dtype=torch.float16)
    assert trunc_exp(h).dtype == torch.float32 and torch.equal(trunc_exp(h), torch.exp(h.float()))


REF = os.environ.get("NSIG_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "nerf", "renderer_wtmk.py")),
                    reason="needs the reference sources (build container only)")
@pytest.mark.parametrize("cuda_ray", [True, False])
def test_render_chunks_rays_like_the_reference_render(cuda_ray):
    """NeRFRenderer.render (renderer_wtmk.py:541-574) against the reference's own method body run on the same stub runner:
    the same sequence of (batch entry, ray range, message, kwargs) calls and the same assembled image / depth, staged and
    unstaged, ragged last chunk and B > 1 included."""
    import types
    import make_golden_field as mgf
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork
    env = {"torch": torch}
    exec(compile(mgf.cut_methods(os.path.join(REF, "nerf", "renderer_wtmk.py"), {"render"})["render"], "ref:render", "exec"), env)

    def make_runner(log):
        def runner(rays_o, rays_d, message, **kw):
            log.append((tuple(rays_o.shape), float(rays_o[0, 0, 0]), None if message is None else message.tolist(), sorted(kw.items())))
            return {"image": rays_o * 2 + rays_d, "depth": rays_o[..., 0] - rays_d[..., 1], "weights_sum": rays_o[..., 2]}
        return runner

    net = NeRFNetwork(bound=1, cuda_ray=cuda_ray, message_dim=2)
    torch.manual_seed(0)
    rays_o, rays_d, msg = torch.randn(2, 1000, 3), torch.randn(2, 1000, 3), torch.tensor([1.0, 0.0])
    for staged, chunk in ((False, 4096), (True, 4096), (True, 300), (True, 1000), (True, 1)):
        if chunk == 1:
            rays_o, rays_d = rays_o[:, :7], rays_d[:, :7]
        log_a, log_b = [], []
        name = "run_cuda" if cuda_ray else "run"
        setattr(net, name, make_runner(log_a))
        stub = types.SimpleNamespace(cuda_ray=cuda_ray, run_cuda=make_runner(log_b), run=make_runner(log_b))
        kw = dict(bg_color=1, perturb=False, dt_gamma=0.0)
        got = net.render(rays_o, rays_d, msg, staged=staged, max_ray_batch=chunk, **kw)
        want = env["render"](stub, rays_o, rays_d, msg, staged=staged, max_ray_batch=chunk, **kw)
        assert log_a == log_b and len(log_a) == (2 * -(-rays_o.shape[1] // chunk) if staged else 1)
        assert set(got) == set(want)
        for k in want:
            assert torch.equal(got[k], want[k]), k
