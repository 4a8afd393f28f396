"""CPU tests (-m "not gpu"): the oracle frames the GPU render tests compare `nsig_render_rays` and `run_cuda` with, against
the REFERENCE'S OWN NeRFRenderer.run_cuda body (renderer_wtmk.py:256-377) run unmodified on CPU over the C oracle
(tests/golden/make_golden_runcuda.py -> tests/golden/runcuda_golden.npz).

tests/test_render_gpu.py::_oracle_frame takes a shortcut: instead of replaying the host-driven alive-ray loop it composites
the samples of the TRAINING march with the inference kill rule.  Here that shortcut is checked against what the reference's
loop really returns - march_rays / composite_rays with the n_step schedule and compaction - including the dense case in which
rays are killed early, so a frame the GPU test accepts is a frame the reference's loop would have produced."""
import os

import numpy as np
import pytest
import torch

import make_golden_runcuda as mg
from test_render_gpu import _oracle_frame

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "runcuda_golden.npz"))


@pytest.mark.parametrize("case", list(mg.CASES))
def test_oracle_frame_is_the_reference_inference_loop_frame(golden, oracle_cpu, case):
    net, rays_o, rays_d, msg, T_thresh = mg.case_inputs(case)
    img, ws, m, used, img_train_rule = _oracle_frame(oracle_cpu, net, rays_o, rays_d, msg, mg.BOUND, T_thresh)
    g_eval, g_train = golden[f"{case}_eval_image"], golden[f"{case}_train_image"]
    # inference rule restated in fp64 on the training march's samples == the reference's alive-ray loop (fp32, `t` re-derived
    # through composite_rays' t += deltas[1])
    assert np.abs(img - g_eval).max() < 5e-6, np.abs(img - g_eval).max()
    # training branch: the oracle chain of __graft_entry__.smoke() == run_cuda(training) of the reference
    assert np.abs(img_train_rule - g_train).max() < 2e-6
    assert int(golden[f"{case}_train_counter"][0]) == m and int(golden[f"{case}_train_counter"][1]) == rays_o.shape[0]
    # the loop evaluates n_alive x n_step slots per iteration (padding included): never fewer than the samples consumed
    assert int(golden[f"{case}_eval_samples_evaluated"]) >= used
    if "dense" in case:
        assert used < 0.25 * m and np.abs(g_eval - g_train).max() > 1e-3      # the kill rule matters in this case
        assert np.abs(g_eval - g_train).max() < T_thresh
    else:
        assert used == m and np.abs(g_eval - g_train).max() < 2e-6


def test_training_branch_depth_and_weights(golden, oracle_cpu):
    """depth = clamp(sum_i w_i t_i - near, 0) / (far - near), weights_sum and the white-background blend of the training
    branch (renderer_wtmk.py:298-303) from the oracle's own march + composite."""
    from oracle import field_oracle as fo
    case = "gain0.6"
    net, rays_o, rays_d, msg, T_thresh = mg.case_inputs(case)
    aabb = np.array([-mg.BOUND] * 3 + [mg.BOUND] * 3, np.float32)
    on, of = oracle_cpu.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    xyz, dirs, deltas, rays, cnt = oracle_cpu.march_rays_train(rays_o, rays_d, mg.BOUND, net.density_bitfield.numpy(), 1, 128, on, of)
    m = int(cnt[0])
    field = mg.FieldOracle(net)
    sig, rgb = field(torch.from_numpy(xyz[:m]), torch.from_numpy(dirs[:m]), torch.from_numpy(msg))
    ws, depth, img = oracle_cpu.composite_rays_train_forward(sig.numpy(), rgb.numpy(), deltas[:m], rays, T_thresh)
    np.testing.assert_allclose(ws, golden[f"{case}_train_weights_sum"], rtol=0, atol=1e-6)
    with np.errstate(invalid="ignore"):                # rays that miss the box: near == far, 0 / 0 on both sides
        want_depth = np.clip(depth - on, 0, None) / (of - on)
    np.testing.assert_allclose(want_depth, golden[f"{case}_train_depth"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(img + (1 - ws)[:, None], golden[f"{case}_train_image"], rtol=0, atol=1e-6)
