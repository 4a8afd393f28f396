#!/usr/bin/env python
"""Time the REFERENCE-COMPOSED CUDA training step on the GPU box ("Reference CUDA path timing"):

  * the UNMODIFIED reference raymarching extension (oracle/_ref/_raymarching.so, built from
    /root/reference/raymarching/src by oracle/build_ref.py) driven by a restatement of its Python wrappers
    (raymarching.py:161-291: worst-case zero-filled sample buffers, `.item()` on the counter, `empty_cache()`);
  * the reference's pure-PyTorch hash encoders running on the GPU (oracle/torch_port.py restates hash_encoding.py /
    hash_encoding_wtmk_bit.py op for op; the reference .py files themselves cannot travel to the box);
  * a bias-free fp16 torch MLP + torch SH standing in for tiny-cuda-nn (NOT installed - this substitution makes the
    reference arm SLOWER than real tcnn would be for the MLP part; the encoders, which dominate, are the reference's own);
  * the same HiDDeN decoder, losses, GradScaler and torch.optim.Adam over get_params-shaped parameter groups.

Same fixture as bench.py (blender_wtmk: 4608 block + 4096 content rays, sphere occupancy, random-init weights).
This is a measurement aid for DESIGN.md ("x the reference CUDA path"); bench.py's --impl reference remains the CPU
port the contract asks for.

    python tools/bench_ref_cuda.py [--steps 10] [--warmup 3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.nn.functional as F
    from torch.autograd import Function

    import make_golden_raymarch as mgr
    from nerf_signature_b200 import harness, synthetic as syn
    from nerf_signature_b200.nerf.hidden_models import get_hidden_decoder_multi_views
    from oracle import torch_port as tp

    dev = torch.device("cuda:0")
    ref = mgr.load_ref()
    cfg = dict(harness.CONFIGS["blender_wtmk"])
    md, bound = cfg["message_dim"], cfg["bound"]
    tp._OFFSETS = tp._OFFSETS.to(dev)

    # ---- field parameters (shapes of nerf/network_wtmk_tcnn.py) on the GPU -------------------------------------
    field = tp.PortField(bound=bound, message_dim=md, seed=0)
    field.base_tables = [t.to(dev) for t in field.base_tables]
    field.base_res = [r.to(dev) for r in field.base_res]
    field.msg_tables = [t.detach().to(dev).requires_grad_(True) for t in field.msg_tables]
    Ws = [w.to(dev).half() for w in field.Ws]
    Wc = [w.to(dev).half() for w in field.Wc]
    torch.manual_seed(0)
    decoder = get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=8, input_ch=3, channels=64).to(dev)
    grid = syn.sphere_grid(1)
    bitfield = torch.from_numpy(syn.packbits_np(grid, 0.5)).to(dev)
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
    C, H, max_steps = 1, 128, 1024

    def sh4(d):
        x, y, z = d.unbind(-1)
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        c = [torch.full_like(x, 0.28209479177387814), -tp._C1 * y, tp._C1 * z, -tp._C1 * x, tp._C2[0] * xy, tp._C2[1] * yz,
             tp._C2[2] * (2.0 * zz - xx - yy), tp._C2[3] * xz, tp._C2[4] * (xx - yy), tp._C3[0] * y * (3 * xx - yy),
             tp._C3[1] * xy * z, tp._C3[2] * y * (4 * zz - xx - yy), tp._C3[3] * z * (2 * zz - 3 * xx - 3 * yy),
             tp._C3[4] * x * (4 * zz - xx - yy), tp._C3[5] * z * (xx - yy), tp._C3[6] * x * (xx - 3 * yy)]
        return torch.stack(c, dim=-1)

    def network(x, d, message):  # nerf/network_wtmk_tcnn.py:97-124
        xn = (x + bound) / (2 * bound)
        feat = tp.hash_embed(xn, field.base_tables, field.base_res, 19)
        m = tp.msg_embed(xn, field.msg_tables, message, 2048.0, 19)
        feat = torch.cat([feat[:, :-2], feat[:, -2:] + m], dim=-1)
        h = torch.relu(feat.half() @ Ws[0].t()) @ Ws[1].t()
        sigma = torch.exp(h[..., 0].float())
        geo = h[..., 1:]
        dd = sh4(((d + 1) / 2) * 2 - 1).half()
        hc = torch.cat([dd, geo, torch.zeros_like(geo[..., :1])], dim=-1)
        hc = torch.relu(hc @ Wc[0].t())
        hc = torch.relu(hc @ Wc[1].t())
        return sigma, torch.sigmoid((hc @ Wc[2].t())[..., :3].float())

    class Composite(Function):  # raymarching.py:238-291
        @staticmethod
        def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh):
            sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
            M, N = sigmas.shape[0], rays.shape[0]
            ws = torch.empty(N, device=dev); depth = torch.empty(N, device=dev); image = torch.empty(N, 3, device=dev)
            ref.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, ws, depth, image)
            ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, depth, image)
            ctx.dims = (M, N, T_thresh)
            return ws, depth, image

        @staticmethod
        def backward(ctx, g_ws, g_depth, g_image):
            sigmas, rgbs, deltas, rays, ws, depth, image = ctx.saved_tensors
            M, N, T_thresh = ctx.dims
            gs, gc = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
            ref.composite_rays_train_backward(g_ws.contiguous(), g_image.contiguous(), sigmas, rgbs, deltas, rays, ws, image,
                                              M, N, T_thresh, gs, gc)
            return gs, gc, None, None, None

    def render(rays_o, rays_d, message):  # renderer_wtmk.py:256-321 with force_all_rays=True, perturb=False
        rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        ref.near_far_from_aabb(rays_o, rays_d, aabb, N, 0.2, nears, fars)
        M = N * max_steps
        xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.zeros(N, device=dev)
        ref.march_rays_train(rays_o, rays_d, bitfield, bound, cfg["dt_gamma"], max_steps, N, C, H, M, nears, fars, xyzs, dirs,
                             deltas, rays, counter, noises)
        m = counter[0].item()
        m += 128 - m % 128
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        torch.cuda.empty_cache()
        sigmas, rgbs = network(xyzs, dirs, message)
        ws, depth, image = Composite.apply(sigmas, rgbs, deltas, rays, 1e-4)
        return image + (1 - ws).unsqueeze(-1) * 1.0, m

    params = [{"params": field.msg_tables}, {"params": list(decoder.parameters())}]
    opt = torch.optim.Adam(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1)
    batches = [{k: torch.from_numpy(v).to(dev) for k, v in harness.make_batch(cfg, seed=1000 + i).items()} for i in range(2)]
    gen = torch.Generator().manual_seed(7)
    n_rays = batches[0]["rays_o"].shape[1] + batches[0]["rays_o_block"].numel() // 3

    def step(b):
        message = torch.randint(0, 2, (md,), generator=gen).float().to(dev)
        opt.zero_grad(set_to_none=True)
        img_w, m1 = render(b["rays_o_block"], b["rays_d_block"], message)
        pred = img_w.view(*b["rays_o_block"].shape).clamp(0, 1)
        with torch.autocast("cuda", dtype=torch.float16):
            decoded = decoder((pred.permute(0, 3, 1, 2) - mean) / std)
        img_c, m2 = render(b["rays_o"], b["rays_d"], message)
        lossi = F.mse_loss(img_c.view(1, -1, 3), b["gt"], reduction="none").mean()
        lossw = F.binary_cross_entropy_with_logits(decoded.float() * 10.0, message.unsqueeze(-1), reduction="mean")
        loss = 0.005 * lossw + lossi
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return float(loss), m1 + m2

    for i in range(args.warmup):
        step(batches[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    samples = 0
    for i in range(args.steps):
        _, m = step(batches[i % 2])
        samples += m
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"what": "reference-composed CUDA step (reference raymarching.cu + reference torch hash encoders on GPU + "
                              "torch fp16 MLP standing in for tcnn)", "ms_per_step": ms, "rays_per_step": n_rays,
                      "rays_per_s": n_rays / ms * 1e3, "samples_per_step": samples / args.steps, "steps": args.steps}))


if __name__ == "__main__":
    main()
