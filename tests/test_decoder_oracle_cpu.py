"""CPU tests (-m "not gpu"): the repo's plain-PyTorch HiDDeN decoder (nerf_signature_b200/nerf/hidden_models.py) - the
oracle the fused decoder kernels are tested against in tests/test_decoder_gpu.py - reproduces the REFERENCE module
(nerf/hidden_models.py:16-35, 104-137, imported unmodified by tests/golden/make_golden_decoder.py): parameter names,
shapes, seeded initial values, logits, the watermark loss and its gradients."""
import os

import numpy as np
import pytest
import torch

import make_golden_decoder as mg
from nerf_signature_b200.nerf import hidden_models as hm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "decoder_golden.npz"))


@pytest.mark.parametrize("case", list(mg.CASES))
def test_decoder_module_matches_reference_module(golden, case):
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        dec, got = mg.run_case(hm, case)
    finally:
        torch.set_num_threads(threads)
    g = {k[len(case) + 1:]: golden[k] for k in golden.files if k.startswith(case + "_")}
    # checkpoints of the reference load: same names, same shapes, in the same order
    assert list(dec.state_dict().keys()) == list(g["keys"])
    assert [str(tuple(v.shape)) for v in dec.state_dict().values()] == list(g["shapes"])
    # same constructor order => the same seed gives the same initial parameters
    sums = np.array([float(p.detach().double().sum()) for p in dec.parameters()])
    np.testing.assert_allclose(sums, g["param_sums"], rtol=0, atol=1e-9)
    # tolerances leave room for another host CPU's convolution kernels (fp32 summation order); on the generating machine
    # every quantity is bit-identical
    np.testing.assert_allclose(got["logits"], g["logits"], rtol=0, atol=2e-5)
    assert abs(float(got["loss"]) - float(g["loss"])) < 1e-4

    def close(a, b, rel):
        assert np.abs(a - b).max() <= rel * np.abs(b).max(), (np.abs(a - b).max(), np.abs(b).max())
    close(got["dpred"], g["dpred"], 1e-3)
    close(got["grad_norms"], g["grad_norms"], 1e-3)
    close(got["grad_linear_w"], g["grad_linear_w"], 1e-3)
    close(got["grad_conv0_w"], g["grad_conv0_w"], 1e-3)


def test_normalisation_is_torchvision_normalize():
    """normalize_img / unnormalize_img: (x - mean) / std per channel with the ImageNet constants (hidden_models.py:13-14)."""
    x = torch.rand(5, 3, 7, 9)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    assert torch.equal(hm.normalize_img(x), (x - mean) / std)
    assert torch.allclose(hm.unnormalize_img(hm.normalize_img(x)), x, atol=1e-6)
    # BatchNorm uses batch statistics in eval mode too (track_running_stats=False): no running buffers in the state dict
    dec = hm.get_hidden_decoder_multi_views(num_bits=1, num_blocks=8)
    assert not any("running" in k for k in dec.state_dict())
    assert all(abs(m.eps - 1e-3) < 1e-12 for m in dec.modules() if isinstance(m, torch.nn.BatchNorm2d))
