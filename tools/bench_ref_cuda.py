#!/usr/bin/env python
"""Time the REFERENCE-COMPOSED CUDA training step on the GPU box ("Reference CUDA path timing"; the
denominator of north_star's ">= 10x the reference torch-ngp/tcnn CUDA path" target).

The step is oracle/ref_cuda_step.py: the UNMODIFIED reference raymarching extension (oracle/_ref/_raymarching.so) + the
reference's torch hash encoders on the GPU + a torch fp16 bias-free MLP / SH standing in for tiny-cuda-nn (NOT installed -
the stand-in is slower than tcnn for the MLP part; the encoders, which dominate, are the reference's own) + the same
HiDDeN decoder, losses, GradScaler and torch.optim.Adam over get_params-shaped groups.  Same fixture as bench.py
(`--config`: 4608 block + 4096 content rays, sphere occupancy, random-init weights).

bench.py runs this file in a SUBPROCESS for its `ref_cuda` leg, so the bench process itself never loads the reference
extension.  Prints one JSON line.

    python tools/bench_ref_cuda.py [--steps 20] [--warmup 3] [--config blender_wtmk]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="blender_wtmk")
    args = ap.parse_args()
    import torch
    from nerf_signature_b200 import harness, synthetic as syn
    from nerf_signature_b200.nerf.hidden_models import get_hidden_decoder_multi_views
    from oracle import ref_cuda_step as rcs
    from oracle import torch_port as tp

    if not os.path.exists(rcs.REF_SO):
        print(json.dumps({"what": "reference-composed CUDA step", "unavailable": "oracle/_ref/_raymarching.so not built"}))
        return
    dev = torch.device("cuda:0")
    ref = rcs.load_ref()
    cfg = dict(harness.CONFIGS[args.config])
    md, bound = cfg["message_dim"], cfg["bound"]
    import math
    cascade = 1 + math.ceil(math.log2(bound))
    field = tp.PortField(bound=bound, message_dim=md, seed=0)
    torch.manual_seed(0)
    decoder = get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=8, input_ch=3, channels=64).to(dev)
    grid = syn.sphere_grid(cascade)
    bitfield = torch.from_numpy(syn.packbits_np(grid, 0.5)).to(dev)
    sp = torch.cat([w.reshape(-1) for w in field.Ws])
    cp = torch.cat([w.reshape(-1) for w in field.Wc])
    step = rcs.RefComposedStep(ref, dev, bound, cascade, bitfield, field.base_tables, field.msg_tables, sp, cp, decoder,
                               dt_gamma=cfg["dt_gamma"], mlp="fp16")
    params = [{"params": step.msg_tables}, {"params": list(decoder.parameters())}]
    opt = torch.optim.Adam(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)   # main_nerf_wtmk.py:107
    scaler = torch.amp.GradScaler("cuda")
    batches = [{k: torch.from_numpy(v).to(dev) for k, v in harness.make_batch(cfg, seed=1000 + i).items()} for i in range(4)]
    gen = torch.Generator().manual_seed(7)
    n_rays = batches[0]["rays_o"].shape[1] + batches[0]["rays_o_block"].numel() // 3

    def one(b):
        message = torch.randint(0, 2, (md,), generator=gen).float().to(dev)
        opt.zero_grad(set_to_none=True)
        out = step.forward_losses(b, message)
        scaler.scale(out["loss"]).backward()
        scaler.step(opt)
        scaler.update()
        return float(out["loss"]), out["samples"]

    for i in range(args.warmup):
        one(batches[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    samples = 0
    for i in range(args.steps):
        _, m = one(batches[i % 4])
        samples += m
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"what": "reference-composed CUDA step (reference raymarching.cu + reference torch hash encoders on GPU "
                              "+ torch fp16 MLP standing in for tcnn, torch.optim.Adam + GradScaler)",
                      "config": args.config, "ms_per_step": ms, "rays_per_step": n_rays, "value": n_rays / ms * 1e3,
                      "unit": "rays/s", "samples_per_step": samples / args.steps, "steps": args.steps,
                      "warmup": args.warmup}))


if __name__ == "__main__":
    main()
