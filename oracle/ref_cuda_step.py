"""Reference-COMPOSED CUDA training step — TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product).

What the reference executes for one watermark training step on a GPU (nerf/utils_wtmk_disen.py:1164-1181 +
579-646 -> nerf/renderer_wtmk.py:256-321 -> nerf/network_wtmk_tcnn.py:97-124), assembled from:

  * the UNMODIFIED reference raymarching extension (oracle/_ref/_raymarching.so, built from
    /root/reference/raymarching/src by oracle/build_ref.py), driven the way raymarching/raymarching.py:161-291
    drives it (worst-case zero-filled sample buffers, `.item()` on the counter, `empty_cache()`);
  * the reference's pure-PyTorch hash encoders on the GPU (oracle/torch_port.py restates hash_encoding.py and
    hash_encoding_wtmk_bit.py op for op and is pinned bit-exactly to outputs of the reference modules,
    tests/test_oracle_cpu.py; the reference .py files cannot travel to the GPU box);
  * SH degree 4 + bias-free MLPs standing in for tiny-cuda-nn, which is not installed anywhere here
    (parity unpinned for this part).  Two arithmetic modes:
      mlp="fp32q": oracle/field_oracle.py - fp32 math on fp16-rounded weights and activations (parity checks);
      mlp="fp16" : torch half matmuls (timing; closest to tcnn's FullyFusedMLP cost);
  * the same HiDDeN decoder module under float16 autocast, torch losses.

Used by tests/test_e2e_ref_parity_gpu.py (same batch, message and weights as the repo's step: losses, rendered pixels,
decoder logits, decoded bits and dL/dS compared) and by tools/bench_ref_cuda.py (the "reference torch-ngp/tcnn CUDA
path" leg of bench.py).
"""
import importlib.util
import os

import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import field_oracle as fo
from . import torch_port as tp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_raymarching.so")


def load_ref():
    """Import the reference extension built by oracle/build_ref.py (pybind module `_raymarching`)."""
    spec = importlib.util.spec_from_file_location("_raymarching", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class RefComposedStep:
    def __init__(self, ref, dev, bound, cascade, bitfield, base_tables, msg_tables, sigma_params, color_params, decoder,
                 dt_gamma=0.0, max_steps=1024, mlp="fp32q", lambda_w=0.005, lambda_i=1.0, density_scale=1.0,
                 min_near=0.2):
        self.ref, self.dev = ref, dev
        self.bound, self.C, self.H = float(bound), int(cascade), 128
        self.bitfield = bitfield
        self.base_tables = [t.detach().to(dev) for t in base_tables]
        self.msg_tables = [t.detach().clone().to(dev).requires_grad_(True) for t in msg_tables]
        self.base_res = [r.to(dev) for r in tp.level_resolutions(16, 2048, 16)]
        self.sigma_params = sigma_params.detach().to(dev).float()
        self.color_params = color_params.detach().to(dev).float()
        self.decoder = decoder
        self.dt_gamma, self.max_steps, self.mlp = float(dt_gamma), int(max_steps), mlp
        self.lambda_w, self.lambda_i = lambda_w, lambda_i
        self.density_scale, self.min_near = float(density_scale), float(min_near)
        self.aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
        if mlp == "fp16":
            sp, cp = self.sigma_params, self.color_params
            self.Ws = [sp[:2048].view(64, 32).half(), sp[2048:3072].view(16, 64).half()]
            self.Wc = [cp[:2048].view(64, 32).half(), cp[2048:6144].view(64, 64).half(), cp[6144:7168].view(16, 64).half()]
        self.mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1)
        self.std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1)

    # ---- nerf/network_wtmk_tcnn.py:97-124 ------------------------------------------------------------------
    def network(self, x, d, message):
        xn = (x + self.bound) / (2 * self.bound)
        feat = tp.hash_embed(xn, self.base_tables, self.base_res, 19)
        if message is not None:
            m = tp.msg_embed(xn, self.msg_tables, message, 2048.0, 19)
            feat = torch.cat([feat[:, :-2], feat[:, -2:] + m], dim=-1)
        if self.mlp == "fp32q":
            sigma, rgb, _, _ = fo.mlp_forward(feat, d, self.sigma_params, self.color_params, self.density_scale)
            return sigma, rgb
        h = torch.relu(feat.half() @ self.Ws[0].t()) @ self.Ws[1].t()
        sigma = self.density_scale * torch.exp(h[..., 0].float())
        geo = h[..., 1:]
        dd = fo.sh4(((d + 1) / 2) * 2 - 1).half()
        hc = torch.cat([dd, geo, torch.zeros_like(geo[..., :1])], dim=-1)
        hc = torch.relu(hc @ self.Wc[0].t())
        hc = torch.relu(hc @ self.Wc[1].t())
        return sigma, torch.sigmoid((hc @ self.Wc[2].t())[..., :3].float())

    # ---- raymarching/raymarching.py:238-291 around the reference kernels --------------------------------------
    def _composite(self, sigmas, rgbs, deltas, rays, T_thresh):
        ref, dev = self.ref, self.dev

        class Composite(Function):
            @staticmethod
            def forward(ctx, sigmas, rgbs):
                sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
                M, N = sigmas.shape[0], rays.shape[0]
                ws = torch.empty(N, device=dev); depth = torch.empty(N, device=dev); image = torch.empty(N, 3, device=dev)
                ref.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, ws, depth, image)
                ctx.save_for_backward(sigmas, rgbs, ws, depth, image)
                ctx.dims = (M, N)
                return ws, depth, image

            @staticmethod
            def backward(ctx, g_ws, g_depth, g_image):
                sigmas, rgbs, ws, depth, image = ctx.saved_tensors
                M, N = ctx.dims
                gs, gc = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
                ref.composite_rays_train_backward(g_ws.contiguous(), g_image.contiguous(), sigmas, rgbs, deltas, rays, ws,
                                                  image, M, N, T_thresh, gs, gc)
                return gs, gc

        return Composite.apply(sigmas, rgbs)

    # ---- nerf/renderer_wtmk.py:256-321 (training branch, force_all_rays=True, perturb=False, bg_color=1) -----------
    def render(self, rays_o, rays_d, message, bg_color=1.0, T_thresh=1e-4):
        ref, dev = self.ref, self.dev
        rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        ref.near_far_from_aabb(rays_o, rays_d, self.aabb, N, self.min_near, nears, fars)
        M = N * self.max_steps
        xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.zeros(N, device=dev)
        ref.march_rays_train(rays_o, rays_d, self.bitfield, self.bound, self.dt_gamma, self.max_steps, N, self.C, self.H, M,
                             nears, fars, xyzs, dirs, deltas, rays, counter, noises)
        m = counter[0].item()
        m += 128 - m % 128
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        torch.cuda.empty_cache()
        sigmas, rgbs = self.network(xyzs, dirs, message)
        ws, depth, image = self._composite(sigmas, rgbs, deltas, rays, T_thresh)
        image = image + (1 - ws).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return {"image": image, "depth": depth, "weights_sum": ws, "samples": m, "rays": rays}

    # ---- nerf/utils_wtmk_disen.py:579-646 -----------------------------------------------------------------------
    def forward_losses(self, batch, message, autocast=True):
        """autocast=True is what the reference runs (`--fp16`, utils_wtmk_disen.py:1170); autocast=False evaluates the
        decoder in fp32 (the ground truth both fp16 paths are measured against in the parity test)."""
        ob = batch["rays_o_block"]
        out_w = self.render(ob, batch["rays_d_block"], message)
        pred = out_w["image"].view(*ob.shape).clamp(0, 1)
        with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            decoded = self.decoder((pred.permute(0, 3, 1, 2) - self.mean) / self.std)
        out_c = self.render(batch["rays_o"], batch["rays_d"], message)
        image_c = out_c["image"].view(batch["gt"].shape)
        lossi = F.mse_loss(image_c, batch["gt"], reduction="none").mean()
        lossw = F.binary_cross_entropy_with_logits(decoded.float() * 10.0, message.unsqueeze(-1), reduction="mean")
        loss = self.lambda_w * lossw + self.lambda_i * lossi
        return {"loss": loss, "lossi": lossi, "lossw": lossw, "pred": pred, "image_c": image_c, "decoded": decoded,
                "block": out_w, "content": out_c, "samples": out_w["samples"] + out_c["samples"]}
