#!/usr/bin/env python
"""Break one watermark training step down by kernel (torch.profiler) and by host time.

    python tools/profile_step.py [--config blender_wtmk] [--steps 5] [--out gpurun_out/profile_step.txt]

Measurement aid only (numbers taken under a profiler are never bench values)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="blender_wtmk")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/profile_step.txt")
    args = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    from nerf_signature_b200 import harness

    dev = torch.device("cuda:0")
    cfg = dict(harness.CONFIGS[args.config])
    mode = os.environ.get("NSIG_RENDER_MODE", "merged")
    scene = harness.Scene(cfg, dev, seed=0, merged_render=mode == "merged", overlap_decoder=mode == "overlap", fused_decoder=os.environ.get("NSIG_TORCH_DECODER") != "1",
                          fused_losses=os.environ.get("NSIG_TORCH_LOSSES") != "1")
    batches = [scene.to_device(harness.make_batch(cfg, seed=i)) for i in range(2)]
    gen = torch.Generator().manual_seed(7)
    for i in range(5):
        scene.train_step(batches[i % 2], scene.new_message(gen))
    torch.cuda.synchronize()

    # host-only cost: wall time of issuing a step vs. device time of executing it
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        scene.train_step(batches[i % 2], scene.new_message(gen))
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    lines = [f"config {args.config}: {args.steps} steps, host issue {1e3 * t_issue / args.steps:.3f} ms/step, "
             f"device {e0.elapsed_time(e1) / args.steps:.3f} ms/step"]

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            scene.train_step(batches[i % 2], scene.new_message(gen))
        torch.cuda.synchronize()
    ka = prof.key_averages()
    rows = [(e.key, e.device_time_total / args.steps, e.count / args.steps) for e in ka if e.device_time_total > 0 and
            e.device_type.name == "CUDA"]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    lines.append(f"sum of kernel device time {tot / 1e3:.3f} ms/step over {sum(r[2] for r in rows):.0f} launches/step")
    for k, us, n in rows[:40]:
        lines.append(f"{us:9.1f} us/step {n:6.1f} x  {k[:110]}")
    out = "\n".join(lines)
    print(out)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        f.write(out + "\n")


if __name__ == "__main__":
    main()
