"""-m gpu: a synthetic tiny-cuda-nn-convention checkpoint (colour net with a constant-1 padding input, i.e. weight column
31 acting as a bias) loaded through nerf_signature_b200.checkpoint reproduces the pad-with-one network: the fused kernels'
outputs are compared with oracle/field_oracle.mlp_forward(pad_value=1) on the ORIGINAL, unconverted parameters."""
import numpy as np
import pytest
import torch

from test_checkpoint_cpu import _tcnn_clean_checkpoint

pytestmark = pytest.mark.gpu


def test_tcnn_layout_checkpoint_reproduces_pad_one_network(oracle_cpu, tmp_path):
    from nerf_signature_b200 import checkpoint as ck
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork
    from oracle import field_oracle as fo
    ckpt = _tcnn_clean_checkpoint(bound=1, seed=3)
    sd = ckpt["model"]
    # a bias large enough to matter: column 31 of the first colour matrix
    sd["color_net.params"][:2048].view(64, 32)[:, 31] = torch.linspace(-1.0, 1.0, 64)
    path = tmp_path / "clean.pth"
    torch.save(ckpt, path)
    net = NeRFNetwork(bound=1, cuda_ray=True, message_dim=4)
    info = ck.load_checkpoint(net, str(path), model_only=True, map_location="cpu")
    assert not info["unexpected_keys"]
    net = net.cuda().eval()
    rs = np.random.RandomState(0)
    M = 3000
    x = rs.uniform(-1, 1, size=(M, 3)).astype(np.float32)
    d = rs.normal(size=(M, 3)); d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    with torch.no_grad():
        sigma, rgb = net(torch.from_numpy(x).cuda(), torch.from_numpy(d).cuda(), None)
    xn = ((x + np.float32(1.0)) * np.float32(0.5)).astype(np.float32)
    tabs = [sd[f"encoder.embeddings.{l}.weight"].numpy() for l in range(16)]
    feat = oracle_cpu.hash_encode_forward(xn, tabs, net.encoder.resolutions, 19)
    osig, orgb, _, _ = fo.mlp_forward(torch.from_numpy(feat), torch.from_numpy(d), sd["sigma_net.params"].float(),
                                      sd["color_net.params"].float(), pad_value=1.0)
    o0sig, o0rgb, _, _ = fo.mlp_forward(torch.from_numpy(feat), torch.from_numpy(d), sd["sigma_net.params"].float(),
                                        sd["color_net.params"].float(), pad_value=0.0)
    assert float((orgb - o0rgb).abs().max()) > 0.05          # the bias does change the colours ...
    np.testing.assert_allclose(rgb.cpu().numpy(), orgb.numpy(), rtol=0, atol=3e-3)   # ... and the loaded network has it
    np.testing.assert_allclose(sigma.cpu().numpy(), osig.numpy(), rtol=3e-3, atol=1e-6)
