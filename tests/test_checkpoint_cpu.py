"""CPU tests (-m "not gpu") of checkpoint compatibility (reference nerf/utils_wtmk_disen.py:1385-1517):
a CLEAN torch-ngp/tiny-cuda-nn-layout checkpoint loads into the watermark network with strict=False semantics, the
tiny-cuda-nn parameter conversion is the documented one, mean_count/mean_density are restored, and save/load round-trips."""
import numpy as np
import torch

from nerf_signature_b200 import checkpoint as ck


def _tcnn_clean_checkpoint(bound=2, seed=0):
    """A synthetic clean checkpoint in the reference's file layout with tiny-cuda-nn-convention MLP parameters
    (flat [out,in] row-major matrices, weight column 31 of the colour net's first matrix acting as a bias)."""
    g = torch.Generator().manual_seed(seed)
    C = 1 + int(np.ceil(np.log2(bound)))
    sd = {f"encoder.embeddings.{l}.weight": (torch.rand(2 ** 19, 2, generator=g) * 2 - 1) * 1e-2 for l in range(16)}
    sd["sigma_net.params"] = torch.randn(3072, generator=g).half() * 0.2           # tcnn checkpoints may hold fp16
    sd["color_net.params"] = torch.randn(7168, generator=g) * 0.2
    sd["density_grid"] = torch.rand(C, 128 ** 3, generator=g)
    sd["density_bitfield"] = torch.randint(0, 256, (C * 128 ** 3 // 8,), generator=g, dtype=torch.uint8)
    sd["step_counter"] = torch.zeros(16, 2, dtype=torch.int32)
    sd["aabb_train"] = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32)
    sd["aabb_infer"] = sd["aabb_train"].clone()
    return {"epoch": 7, "global_step": 1234, "stats": {"loss": [0.1]}, "mean_count": 4321, "mean_density": 0.37, "model": sd}


def test_tcnn_conversion_is_identity_plus_bias_fold():
    flat = torch.arange(7168, dtype=torch.float32) * 1e-3
    out = ck.convert_tcnn_params(flat, "color")
    w_in, w_out = flat[:2048].view(64, 32), out[:2048].view(64, 32)
    assert torch.equal(out[2048:], flat[2048:])                       # hidden and output matrices: untouched
    assert torch.equal(w_out[:, 1:31], w_in[:, 1:31])
    assert float(w_out[:, 31].abs().max()) == 0.0
    np.testing.assert_allclose(w_out[:, 0].numpy(), (w_in[:, 0] + w_in[:, 31] / ck.SH_C0_FP16).numpy(), rtol=1e-6)
    s = torch.randn(3072).half()
    assert torch.equal(ck.convert_tcnn_params(s, "sigma"), s.float())  # sigma net: no padded input, identity
    # the fold reproduces the pad-with-one network: W x(pad=1) == W' x(pad=0) for inputs whose SH band 0 is the constant
    x = torch.randn(5, 32); x[:, 0] = ck.SH_C0_FP16
    x1, x0 = x.clone(), x.clone()
    x1[:, 31], x0[:, 31] = 1.0, 0.0
    np.testing.assert_allclose((x0 @ w_out.t()).numpy(), (x1 @ w_in.t()).numpy(), rtol=1e-5, atol=1e-5)
    try:
        ck.convert_tcnn_params(torch.zeros(100), "sigma")
        assert False
    except ValueError:
        pass


def test_clean_checkpoint_loads_into_watermark_network_with_strict_false_semantics(tmp_path):
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork
    ckpt = _tcnn_clean_checkpoint(bound=2)
    path = tmp_path / "ngp_ep0007.pth"
    torch.save(ckpt, path)
    net = NeRFNetwork(bound=2, cuda_ray=True, message_dim=4)
    msg0 = net.msg_encoder.embeddings[0].weight.detach().clone()
    info = ck.load_checkpoint(net, str(path), model_only=True, map_location="cpu")
    assert info["unexpected_keys"] == []
    assert info["missing_keys"] and all(k.startswith(("msg_encoder.", "msg_decoder.")) for k in info["missing_keys"])
    assert torch.equal(net.msg_encoder.embeddings[0].weight, msg0)          # untouched: randomly initialised, trainable
    sd = ckpt["model"]
    assert torch.equal(net.encoder.embeddings[5].weight, sd["encoder.embeddings.5.weight"])
    assert torch.equal(net.density_grid, sd["density_grid"]) and torch.equal(net.density_bitfield, sd["density_bitfield"])
    assert torch.equal(net.sigma_net.params, sd["sigma_net.params"].float())
    assert torch.equal(net.color_net.params, ck.convert_tcnn_params(sd["color_net.params"], "color"))
    assert net.mean_count == 4321 and abs(net.mean_density - 0.37) < 1e-12
    # frozen parts stay frozen after loading (network_wtmk_tcnn.py:90-95)
    assert not net.sigma_net.params.requires_grad and not net.encoder.embeddings[0].weight.requires_grad
    # a bare state dict is loaded strictly, like the reference (L1469-1472)
    from nerf_signature_b200.nerf.network_hash import NeRFNetwork as Clean
    clean = Clean(bound=2, cuda_ray=True)
    ck.load_checkpoint(clean, dict(ckpt["model"]))
    assert torch.equal(clean.color_net.params, net.color_net.params)


def test_save_load_round_trip_is_not_converted_twice(tmp_path):
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork
    a = NeRFNetwork(bound=1, cuda_ray=True, message_dim=2)
    with torch.no_grad():
        a.color_net.params[:2048].view(64, 32)[:, 31] = 0.5      # would be folded if the file were taken for a tcnn one
    a.mean_count, a.mean_density = 99, 1.5
    p = ck.save_checkpoint(a, str(tmp_path / "w.pth"), epoch=3, global_step=30)
    b = NeRFNetwork(bound=1, cuda_ray=True, message_dim=2)
    info = ck.load_checkpoint(b, p, map_location="cpu")
    assert info["missing_keys"] == [] and info["unexpected_keys"] == [] and info["epoch"] == 3 and info["global_step"] == 30
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(va, vb), k
    assert b.mean_count == 99 and b.mean_density == 1.5
