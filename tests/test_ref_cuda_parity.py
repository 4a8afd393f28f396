"""-m gpu: pin the C oracle (oracle/raymarch_oracle.c) directly against the UNMODIFIED reference CUDA
extension (oracle/_ref/_raymarching.so, built by oracle/build_ref.py) at sizes larger than the committed
golden fixtures.  Skipped (not failed) when the extension did not travel with the snapshot — the
committed fixtures in tests/golden/raymarch_golden.npz then remain the pin (tests/test_oracle_cpu.py)."""
import os

import numpy as np
import pytest
import torch

import make_golden_raymarch as mgr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "_raymarching.so")):
        pytest.skip("oracle/_ref/_raymarching.so not present")
    return mgr.load_ref()


BIG = [
    ("blender_sphere", "blender", 2048, 13, 1.0, 1, "sphere", 0.0, False),
    ("blender_bernoulli_gamma", "blender", 2048, 14, 1.0, 1, "bernoulli", 1.0 / 128, True),
    ("r360_bernoulli_gamma", "360", 2048, 16, 2.0, 2, "bernoulli", 1.0 / 128, True),
    ("bound1p5_bernoulli", "360", 1024, 17, 1.5, 2, "bernoulli", 0.0, True),
    ("bound4_gamma", "360", 1024, 18, 4.0, 3, "bernoulli", 1.0 / 256, False),
]


@pytest.mark.parametrize("case", BIG, ids=[c[0] for c in BIG])
def test_oracle_march_equals_reference_cuda(ref, oracle_cpu, case):
    name, cam, N, seed, bound, C, kind, dt_gamma, perturb = case
    rays_o, rays_d, bitfield, aabb, noises = mgr.case_inputs(case)
    cu = mgr.cu
    t_o, t_d, t_b = cu(rays_o), cu(rays_d), cu(bitfield)
    nears = torch.empty(N, device="cuda"); fars = torch.empty(N, device="cuda")
    ref.near_far_from_aabb(t_o, t_d, cu(aabb), N, 0.2, nears, fars)
    on, of = oracle_cpu.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    assert np.array_equal(nears.cpu().numpy().view(np.uint32), on.view(np.uint32))
    assert np.array_equal(fars.cpu().numpy().view(np.uint32), of.view(np.uint32))
    M = N * 1024
    xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
    rays = torch.empty(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    ref.march_rays_train(t_o, t_d, t_b, bound, dt_gamma, 1024, N, C, 128, M, nears, fars, xyzs, dirs, deltas, rays, counter,
                         cu(noises))
    torch.cuda.synchronize()
    ox, od, ol, orays, ocnt = oracle_cpu.march_rays_train(rays_o, rays_d, bound, bitfield, C, 128, on, of, noises=noises,
                                                          dt_gamma=dt_gamma)
    assert counter.cpu().numpy().tolist() == ocnt.tolist()
    counts, (cx, cd, cl) = mgr.canonical(rays.cpu().numpy(), xyzs.cpu().numpy(), dirs.cpu().numpy(), deltas.cpu().numpy())
    assert np.array_equal(counts, orays[:, 2])
    m = int(ocnt[0])
    assert np.array_equal(cx.view(np.uint32), ox[:m].view(np.uint32))
    assert np.array_equal(cd.view(np.uint32), od[:m].view(np.uint32))
    assert np.array_equal(cl.view(np.uint32), ol[:m].view(np.uint32))
