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


REF = os.environ.get("NSIG_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "nerf", "utils_wtmk_disen.py")),
                    reason="needs the reference sources (build container only)")
@pytest.mark.parametrize("kind", ["none", "noise", "rotation", "scaling", "blurring", "brightness"])
def test_distortion_layer_matches_reference_body(kind):
    """nerf.distortion.distortion_layer against Trainer.distortion_layer (utils_wtmk_disen.py:551-577), cut out of the
    reference source and run unmodified: identical output bits from the same generator state, same gradient to the pixels."""
    import types
    import torchvision.transforms as T
    import make_golden_field as mgf
    from nerf_signature_b200.nerf.distortion import distortion_layer
    env = {"torch": torch, "F": torch.nn.functional, "T": T}
    exec(compile(mgf.cut_methods(os.path.join(REF, "nerf", "utils_wtmk_disen.py"), {"distortion_layer"})["distortion_layer"],
                 "ref:distortion_layer", "exec"), env)
    g = torch.Generator().manual_seed(7)
    pixels = torch.rand(6, 12, 12, 3, generator=g)
    outs = []
    for fn in (lambda p: env["distortion_layer"](types.SimpleNamespace(distortion=kind), p), lambda p: distortion_layer(p, kind)):
        p = pixels.clone().requires_grad_(True)
        torch.manual_seed(123)
        out = fn(p)
        (out * torch.linspace(0, 1, out.numel()).view(out.shape)).sum().backward()
        outs.append((out.detach(), p.grad))
    (want, gwant), (got, ggot) = outs
    assert got.shape == want.shape and torch.equal(got, want)
    assert torch.equal(ggot, gwant)
    if kind == "scaling":
        assert got.shape[:2] == (6, 12) and got.shape[3] == 3 and 9 <= got.shape[2] <= 15      # only the width is resampled
    elif kind != "none":
        assert got.shape == pixels.shape and not torch.equal(got, pixels)


def test_distortion_layer_rejects_unknown_kinds():
    from nerf_signature_b200.nerf.distortion import distortion_layer, KINDS, CAPTURABLE
    assert set(CAPTURABLE) < set(KINDS)
    with pytest.raises(ValueError):
        distortion_layer(torch.zeros(1, 2, 2, 3), "jpeg")
