#!/usr/bin/env python
"""Time the fused field kernels in isolation on ray-coherent samples of the Blender fixture (what one render pass
of the training step feeds them): nsig_field_forward, nsig_field_backward, and the march/composite kernels.

    python tools/bench_field.py [--rays 4608] [--iters 50]
Between timed launches a 256 MB buffer is written so the tables' L2 residency is what a real step would see
(the step's other kernels also flush L2)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4608)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    from nerf_signature_b200 import _lib, harness, synthetic as syn
    from nerf_signature_b200 import raymarching as rm
    from nerf_signature_b200.nerf import field_ops as fo
    dev = torch.device("cuda:0")
    cfg = dict(harness.CONFIGS["blender_wtmk"])
    scene = harness.Scene(cfg, dev, seed=0, optimizer="torch")
    net = scene.model
    o, d = syn.blender_rays(args.rays, seed=3, H=400, W=400)
    ro, rd = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    nears, fars = rm.near_far_from_aabb(ro, rd, net.aabb_train, net.min_near)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = rm.march_rays_train(ro, rd, net.bound, net.density_bitfield, net.cascade, net.grid_size, nears,
                                                   fars, counter, -1, False, 128, True, 0.0, 1024)
    M = xyzs.shape[0]
    msg = torch.randint(0, 2, (cfg["message_dim"],)).float()
    S = net._summed_table(msg).detach()
    cfgf = net._cfg(1.0)
    tabs = net.encoder.tables()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def timeit(fn, iters):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(iters):
            if not args.no_flush:
                flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    out = {"samples": M, "rays": args.rays}
    Sg = S.clone().requires_grad_(True)
    holder = {}

    def fwd():
        holder["o"] = fo.field_forward(xyzs, dirs, Sg, None, cfgf, net.sigma_net, net.color_net, tabs)

    def fwd_nograd():
        with torch.no_grad():
            fo.field_forward(xyzs, dirs, S, None, cfgf, net.sigma_net, net.color_net, tabs)

    def timed(name, fn, iters):
        for _ in range(3):
            fn()
        _lib.timing_enable([name])
        for _ in range(iters):
            if not args.no_flush:
                flush.fill_(1.0)
            fn()
        t = _lib.timing_collect()[name]
        return t["ms"] / t["n"]

    out["field_fwd_ms"] = timed("nsig_field_forward", fwd, args.iters)
    out["field_fwd_nosave_ms"] = timed("nsig_field_forward", fwd_nograd, args.iters)
    fwd()
    sig, rgb = holder["o"]
    gs, gc = torch.randn_like(sig) * 1e-3, torch.randn_like(rgb) * 1e-3
    # NSIG_BWD = masks (default) | recompute | tc selects the kernel
    bwd_names = ["nsig_field_backward", "nsig_field_backward_tc", "nsig_field_backward_masks", "nsig_field_backward_tc_masks"]
    _lib.timing_enable(bwd_names)

    def bwd():
        torch.autograd.grad([sig, rgb], [Sg], [gs, gc], retain_graph=True)

    for _ in range(3):
        bwd()
    _lib.timing_enable(bwd_names)
    for _ in range(args.iters):
        if not args.no_flush:
            flush.fill_(1.0)
        bwd()
    tt = _lib.timing_collect()
    name = [n for n in bwd_names if n in tt][0]
    t = tt[name]
    out["field_bwd_kernel"] = name
    out["field_bwd_ms"] = t["ms"] / t["n"]
    out["fwd_Gsamples_per_s"] = M / out["field_fwd_ms"] / 1e6
    out["fwd_alg_GBps"] = 1128 * M / out["field_fwd_ms"] / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
