"""torch-fp32 restatement of NeRFNetwork.forward for the parts the reference delegates to
tiny-cuda-nn (SH degree 4 + two bias-free ReLU MLPs) — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED for the MLP arithmetic: tiny-cuda-nn is an unvendored, unpinned, uninstalled
dependency of the reference (nerf/network_wtmk_tcnn.py:7,52-88) and the reference has
no test or golden vector for it.  What is pinned: the SH basis (against the reference's own
hash_encoding.SHEncoder, tests/golden/hash_golden.npz `sh_*`), the network wiring
(network_wtmk_tcnn.py:97-176: against the reference's own forward / density / color bodies run unmodified on CPU around
stand-ins for the tcnn modules, tests/golden/field_golden.npz, tests/test_field_oracle_cpu.py) and the hash features (oracle/hash_oracle.c, bit-exact against the
reference modules).  The MLP follows nerf/"network copy.py":33-68 (bias-free Linear + ReLU) with the
storage precision of this implementation made explicit: fp16 weights, fp16 inputs and hidden
activations, fp32 accumulation.
"""
import numpy as np
import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def sh4(v):
    """hash_encoding.py:162-193 (degree 4) at unit vectors v [M,3]."""
    x, y, z = v.unbind(-1)
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    out = [torch.full_like(x, C0), -C1 * y, C1 * z, -C1 * x,
           C2[0] * xy, C2[1] * yz, C2[2] * (2.0 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy),
           C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
           C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
           C3[6] * x * (xx - 3 * yy)]
    return torch.stack(out, -1)


def _q(t):
    """round to fp16 storage, keep computing in fp32 (straight-through for autograd)."""
    return t + (t.half().float() - t).detach()


def split_params(sigma_params, color_params):
    sp, cp = sigma_params.float(), color_params.float()
    Ws = [sp[:2048].view(64, 32), sp[2048:3072].view(16, 64)]
    Wc = [cp[:2048].view(64, 32), cp[2048:6144].view(64, 64), cp[6144:7168].view(16, 64)]
    return [_q(w) for w in Ws], [_q(w) for w in Wc]


def mlp_forward(feat, dirs, sigma_params, color_params, density_scale=1.0, pad_value=0.0):
    """feat [M,32] fp32 encoder output (incl. message feature), dirs [M,3].  Returns sigma [M], rgb [M,3],
    and the pre-activation outputs (logit [M], geo [M,15]).  network_wtmk_tcnn.py:107-124.
    pad_value: what the colour net's 32nd (padding) input holds.  0 = this repo's native parameterisation;
    1 = tiny-cuda-nn's convention as recalled in SURVEY 8c (inputs padded to a multiple of 16 with the constant 1, which
    turns weight column 31 into a bias) - used by the checkpoint-conversion test."""
    Ws, Wc = split_params(sigma_params, color_params)
    x = _q(feat)
    h = _q(torch.relu(x @ Ws[0].t()))
    out = h @ Ws[1].t()
    logit, geo = out[:, 0], out[:, 1:16]
    sigma = density_scale * torch.exp(logit)
    d = (dirs + 1) / 2            # network_wtmk_tcnn.py:114
    v = d * 2 - 1                 # tcnn's SphericalHarmonics maps [0,1] back to [-1,1]
    cin = torch.cat([_q(sh4(v)), _q(geo), torch.full_like(geo[:, :1], float(pad_value))], -1)
    h1 = _q(torch.relu(cin @ Wc[0].t()))
    h2 = _q(torch.relu(h1 @ Wc[1].t()))
    o = h2 @ Wc[2].t()
    rgb = torch.sigmoid(o[:, :3])
    return sigma, rgb, logit, geo


def trunc_exp_grad(logit, g):
    """activation.py:14-16"""
    return g * torch.exp(logit.clamp(-15, 15))
