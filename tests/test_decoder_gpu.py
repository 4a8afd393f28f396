"""-m gpu parity of the fused HiDDeN decoder kernels (csrc/decoder.cu) against the plain PyTorch module
(nerf/hidden_models.py: HiddenDecoder_multi_views under float16 autocast, exactly as the training step calls it) and
against the same module in fp32.  Tolerances: both the autocast module and the kernels round activations and
activation gradients to fp16 at the same points but sum in different orders, so each is compared with the fp32
result and they must be equally close: logits 3e-3 absolute, gradients 3e-2 relative L2."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _decoder(seed, num_bits=1, redundancy=1):
    from nerf_signature_b200.nerf.hidden_models import get_hidden_decoder_multi_views
    torch.manual_seed(seed)
    dec = get_hidden_decoder_multi_views(num_bits=num_bits, redundancy=redundancy, num_blocks=8, input_ch=3, channels=64)
    with torch.no_grad():   # non-trivial BatchNorm affine parameters and biases
        for m in dec.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    return dec.cuda()


def _run_module(dec, image, gout, autocast):
    from nerf_signature_b200.nerf.hidden_models import normalize_img
    dec.zero_grad(set_to_none=True)
    image = image.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
        out = dec(normalize_img(image.permute(0, 3, 1, 2)))
    (out.float() * gout).sum().backward()
    return out.float().detach(), image.grad.detach(), [p.grad.detach().clone() for p in dec.parameters()]


def _run_fused(dec, image, gout):
    from nerf_signature_b200.nerf import decoder_ops
    dec.zero_grad(set_to_none=True)
    image = image.clone().requires_grad_(True)
    out = decoder_ops.decode(dec, image)
    (out * gout).sum().backward()
    return out.detach(), image.grad.detach(), [p.grad.detach().clone() for p in dec.parameters()]


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("B,H,W,bits,red", [(32, 12, 12, 1, 1), (8, 23, 31, 1, 1), (3, 5, 7, 1, 1), (5, 9, 16, 2, 3)])
def test_fused_decoder_matches_module(B, H, W, bits, red):
    from nerf_signature_b200.nerf import decoder_ops
    dec = _decoder(0, bits, red)
    assert decoder_ops.decoder_params(dec) is not None
    g = torch.Generator(device="cuda").manual_seed(1)
    image = torch.rand(B, H, W, 3, device="cuda", generator=g)
    gout = torch.randn(B, bits, device="cuda", generator=g) * 64.0     # GradScaler-sized upstream gradient
    o32, di32, gp32 = _run_module(dec, image, gout, autocast=False)
    o16, di16, gp16 = _run_module(dec, image, gout, autocast=True)
    of, dif, gpf = _run_fused(dec, image, gout)
    assert of.shape == o32.shape == (B, bits)
    scale = float(o32.abs().max()) + 1e-6
    err_f, err_16 = float((of - o32).abs().max()), float((o16 - o32).abs().max())
    assert err_f <= max(3e-3 * max(scale, 1.0), 2.0 * err_16), (err_f, err_16)
    assert _rel(dif, di32) <= max(3e-2, 2.0 * _rel(di16, di32)), (_rel(dif, di32), _rel(di16, di32))
    names = [n for n, _ in dec.named_parameters()]
    for n, a, b16, b32 in zip(names, gpf, gp16, gp32):
        if n.endswith("layers.0.bias"):
            # conv bias under BatchNorm: the true gradient is 0, what remains is rounding noise - compare magnitudes only
            assert float(a.abs().max()) <= 10 * float(b16.abs().max()) + 1e-2 * float(gp32[0].abs().max()), n
            continue
        assert a.shape == b32.shape
        assert _rel(a, b32) <= max(3e-2, 2.0 * _rel(b16, b32)), (n, _rel(a, b32), _rel(b16, b32))


def test_fused_decoder_accumulates_into_existing_grads_and_skips_frozen():
    from nerf_signature_b200.nerf import decoder_ops
    dec = _decoder(2)
    image = torch.rand(4, 12, 12, 3, device="cuda")
    gout = torch.ones(4, 1, device="cuda")
    _, _, g1 = _run_fused(dec, image, gout)
    out = decoder_ops.decode(dec, image)            # second backward without zeroing: gradients add up
    (out * gout).sum().backward()
    for (n, p), a in zip(dec.named_parameters(), g1):
        if n.endswith("layers.0.bias"):
            continue                                  # conv bias under BatchNorm: pure rounding noise (true gradient 0)
        torch.testing.assert_close(p.grad, 2 * a, rtol=1e-3, atol=1e-6 + 1e-2 * float(a.abs().max()))  # fp32 atomics: order noise
    dec.zero_grad(set_to_none=True)
    for p in dec.parameters():
        p.requires_grad_(False)
    img = image.clone().requires_grad_(True)
    decoder_ops.decode(dec, img).sum().backward()    # frozen decoder: only the image gradient
    assert img.grad is not None and all(p.grad is None for p in dec.parameters())


def test_weight_gradients_are_deterministic_and_prepared_weights_change_nothing():
    """(1) The weight-gradient reduction sums the per-CTA partials in a fixed order: two backward passes over the same inputs
    give bit-identical parameter gradients.  (2) Handing the kernels fp16 weights converted ahead of time
    (decoder_ops.PreparedWeights) is the same arithmetic as converting them inside the forward: bit-identical outputs and
    gradients; after a weight update a refresh() is needed and sufficient."""
    from nerf_signature_b200.nerf import decoder_ops
    dec = _decoder(3)
    g = torch.Generator(device="cuda").manual_seed(5)
    image = torch.rand(32, 12, 12, 3, device="cuda", generator=g)
    gout = torch.randn(32, 1, device="cuda", generator=g) * 64.0
    o1, di1, gp1 = _run_fused(dec, image, gout)
    o2, di2, gp2 = _run_fused(dec, image, gout)
    assert torch.equal(o1, o2) and torch.equal(di1, di2)
    for a, b in zip(gp1, gp2):
        assert torch.equal(a, b)

    def run_prepared(prep):
        dec.zero_grad(set_to_none=True)
        img = image.clone().requires_grad_(True)
        out = decoder_ops.decode(dec, img, prep)
        (out * gout).sum().backward()
        return out.detach(), img.grad.detach(), [p.grad.detach().clone() for p in dec.parameters()]

    prep = decoder_ops.PreparedWeights(dec)
    o3, di3, gp3 = run_prepared(prep)
    assert torch.equal(o1, o3) and torch.equal(di1, di3)
    for a, b in zip(gp1, gp3):
        assert torch.equal(a, b)
    with torch.no_grad():
        for p in dec.parameters():
            p.add_(0.01 * torch.randn(p.shape, device="cuda", generator=g))
    o4, di4, gp4 = _run_fused(dec, image, gout)          # converts the new weights itself
    o_stale, _, _ = run_prepared(prep)                    # still the old conv weights
    assert not torch.equal(o4, o_stale)
    prep.refresh()
    o5, di5, gp5 = run_prepared(prep)
    assert torch.equal(o4, o5) and torch.equal(di4, di5)
    for a, b in zip(gp4, gp5):
        assert torch.equal(a, b)


def test_gelu_arithmetic_over_all_fp16_inputs():
    """The conv kernels apply GELU / GELU' (exact erf form, nn.GELU(): hidden_models.py:26) while staging their tiles.  Sweep
    EVERY finite fp16 input through the kernels' own device functions (nsig_decoder_gelu_probe): fp16(GELU(y)) must be torch's
    float16 GELU bit for bit (same fp32 erff formula), and GELU' within one fp16 step of the float64 value, exact on > 99 %."""
    from nerf_signature_b200 import _lib
    P = _lib.ptr
    bits = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16)
    y = bits.view(torch.float16)
    y = y[torch.isfinite(y)].cuda().contiguous()
    n = y.numel()
    g = torch.empty_like(y)
    dg = torch.empty_like(y)
    _lib.call("nsig_decoder_gelu_probe", P(y), n, P(g), P(dg))
    ref16 = torch.nn.functional.gelu(y)                       # torch's fp16 kernel (fp32 erff inside)
    y64 = y.double()
    cdf = 0.5 * (1.0 + torch.erf(y64 / 2.0 ** 0.5))
    exact_grad = (cdf + y64 * torch.exp(-0.5 * y64 * y64) / (2.0 * torch.pi) ** 0.5).half()

    def steps(a, b):   # distance in fp16 steps (monotone integer map of the bit patterns)
        def key(t):
            i = t.view(torch.int16).int()
            return torch.where(i < 0, -(i & 0x7FFF), i)
        return (key(a) - key(b)).abs()

    assert int(steps(g, ref16).max()) == 0, int((g != ref16).sum())
    assert int(steps(dg, exact_grad).max()) <= 1
    assert int((dg != exact_grad).sum()) < 0.01 * n, int((dg != exact_grad).sum())
    with pytest.raises(_lib.NsigError):
        _lib.call("nsig_decoder_gelu_probe", P(y), n, None, P(dg))


def test_deferred_weight_gradients_equal_the_joined_backward():
    """decode(..., defer_weight_grads=True): the backward hands back the image gradient with the conv weight gradients still on
    the library's side streams; after decoder_ops.finish_backward() every gradient is bit-identical to the backward that joins
    them itself (same kernels, same deterministic reduction).  A second backward without finish_backward() completes the first
    one's tail before it starts (gradients accumulate into .grad)."""
    from nerf_signature_b200.nerf import decoder_ops
    dec = _decoder(3)
    g = torch.Generator(device="cuda").manual_seed(9)
    image = torch.rand(32, 12, 12, 3, device="cuda", generator=g)
    gout = torch.randn(32, 1, device="cuda", generator=g) * 64.0
    o_ref, di_ref, gp_ref = _run_fused(dec, image, gout)

    def run(defer):
        img = image.clone().requires_grad_(True)
        out = decoder_ops.decode(dec, img, defer_weight_grads=defer)
        (out * gout).sum().backward()
        return out.detach(), img.grad.detach()

    dec.zero_grad(set_to_none=True)
    o, di = run(True)
    junk = torch.randn(1 << 22, device="cuda").sin_()          # work on the caller's stream between backward and finish
    decoder_ops.finish_backward()
    torch.cuda.synchronize()
    assert torch.equal(o, o_ref) and torch.equal(di, di_ref)
    for p, ref in zip(dec.parameters(), gp_ref):
        assert torch.equal(p.grad, ref)
    # two deferred backwards in a row, one finish at the end: .grad holds exactly the sum of both
    dec.zero_grad(set_to_none=True)
    run(True)
    run(True)
    decoder_ops.finish_backward()
    decoder_ops.finish_backward()      # idempotent
    torch.cuda.synchronize()
    for p, ref in zip(dec.parameters(), gp_ref):
        torch.testing.assert_close(p.grad, 2 * ref, rtol=1e-6, atol=1e-6 * float(ref.abs().max()))
    del junk
