#!/usr/bin/env python
"""Full-frame inference timing (BASELINE configs[3]): 800x800 Blender and 1008x756 LLFF-shaped frames through
NeRFRenderer.render (eval mode), fused persistent renderer vs the reference-shaped host loop."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=10)
    ap.add_argument("--loop-views", type=int, default=0, help="also time the host-driven loop on this many views")
    ap.add_argument("--sigma-gain", type=float, default=0.0)
    args = ap.parse_args()
    import numpy as np
    import torch
    from nerf_signature_b200 import synthetic as syn
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork
    dev = torch.device("cuda:0")
    out = {}
    for name, H, W, bound, fov, radius in (("blender_800x800", 800, 800, 1.0, 0.6911112, 4.0311 * 0.8),
                                           ("llff_1008x756", 756, 1008, 2.0, 0.9, 4.0 * 0.33)):
        torch.manual_seed(0)
        net = NeRFNetwork(bound=bound, cuda_ray=True, message_dim=32)
        if args.sigma_gain:
            with torch.no_grad():
                net.sigma_net.params[2048:2048 + 64] += args.sigma_gain
        net = net.to(dev).eval()
        grid = syn.sphere_grid(net.cascade)
        net.density_grid.copy_(torch.from_numpy(grid))
        net.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, 0.5)))
        msg = torch.randint(0, 2, (32,)).float().to(dev)
        focal = 0.5 * W / math.tan(0.5 * fov)
        rs = np.random.RandomState(0)
        frames = []
        for v in range(args.views):
            pose = syn.orbit_pose(rs.uniform(math.pi / 3, 2 * math.pi / 3), rs.uniform(0, 2 * math.pi), radius)
            o, d = syn.camera_rays(pose, H, W, focal, np.arange(H * W))
            frames.append((torch.from_numpy(o)[None].to(dev), torch.from_numpy(d)[None].to(dev)))
        kw = dict(bg_color=1, perturb=False, dt_gamma=0.0, max_steps=1024)
        res = {}
        for mode, fused, staged, nviews in (("fused_whole_frame", True, False, args.views),
                                            ("fused_staged_4096", True, True, args.views),
                                            ("host_loop_staged_4096", False, True, args.loop_views)):
            if nviews == 0:
                continue
            net.fused_inference = fused
            with torch.no_grad():
                net.render(*frames[0], msg, staged=staged, **kw)  # warm-up
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                samples = 0
                e0.record()
                for v in range(nviews):
                    r = net.render(*frames[v], msg, staged=staged, **kw)
                    if fused and not staged:
                        samples += int(net.last_render_samples)   # D2H read of the frame's sample count
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / nviews
            res[mode] = {"ms_per_frame": ms, "views": nviews}
            if samples:
                res[mode]["samples_per_frame"] = samples / nviews
                res[mode]["Gsamples_per_s"] = samples / nviews / ms / 1e6
        out[name] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
