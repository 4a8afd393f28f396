#!/usr/bin/env python
"""Kernel timeline of ONE replay of the captured training step (the graph bench.py times): start offset, duration and
stream of every kernel, from torch.profiler's CUPTI activity records.  Shows which branches of the step graph actually
overlap and where the device idles.  Measurement aid only (numbers under a profiler are never bench values).

    python tools/timeline_step.py [--config blender_wtmk] [--out gpurun_out/timeline_step.txt] [--no-defer]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="blender_wtmk")
    ap.add_argument("--out", default="gpurun_out/timeline_step.txt")
    ap.add_argument("--no-defer", action="store_true")
    ap.add_argument("--replays", type=int, default=3)
    args = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    from nerf_signature_b200 import harness

    dev = torch.device("cuda:0")
    cfg = dict(harness.CONFIGS[args.config])
    scene = harness.Scene(cfg, dev, seed=0, optimizer="fused", graph=True, merged_render=True, fused_decoder=True,
                          fused_losses=True, defer_optimizer=not args.no_defer)
    batches = [scene.to_device(harness.make_batch(cfg, seed=i)) for i in range(2)]
    gen = torch.Generator().manual_seed(7)
    for i in range(6):
        scene.train_step(batches[i % 2], scene.new_message(gen))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(args.replays):
            scene.train_step(batches[i % 2], scene.new_message(gen))
            torch.cuda.synchronize()
    import json
    trace = os.path.join(os.path.dirname(args.out) or ".", "timeline_trace.json")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    prof.export_chrome_trace(trace)
    with open(trace) as f:
        tr = json.load(f)
    os.remove(trace)

    class Ev:
        def __init__(self, d):
            self.name, self.start, self.end = d["name"], float(d["ts"]), float(d["ts"]) + float(d["dur"])
            self.stream = d.get("args", {}).get("stream", 0)

    evs = [Ev(d) for d in tr["traceEvents"] if d.get("ph") == "X" and d.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    evs.sort(key=lambda e: e.start)
    # split into replays at gaps > 200 us and keep the last one
    groups, cur = [], []
    for e in evs:
        if cur and e.start - max(x.end for x in cur) > 200:
            groups.append(cur)
            cur = []
        cur.append(e)
    if cur:
        groups.append(cur)
    groups = [g for g in groups if len(g) > 20]
    g = groups[-1]
    t0 = g[0].start
    streams = sorted({e.stream for e in g})
    lines = [f"config {args.config}, defer_optimizer={not args.no_defer}: {len(g)} device activities in the last replay, "
             f"span {max(e.end for e in g) - t0:.1f} us, streams {streams}",
             f"{'start us':>9} {'dur us':>8} {'end us':>8}  stream  kernel"]
    for e in g:
        lines.append(f"{e.start - t0:9.1f} {e.end - e.start:8.1f} {e.end - t0:8.1f}  {streams.index(e.stream):>6}  {e.name[:100]}")
    # idle time of the device: gaps in the union of all kernel intervals
    iv = sorted((e.start, e.end) for e in g)
    idle, end = 0.0, iv[0][1]
    for s, e_ in iv[1:]:
        if s > end:
            idle += s - end
        end = max(end, e_)
    lines.append(f"device idle (no kernel of any stream running): {idle:.1f} us of {end - t0:.1f} us")
    out = "\n".join(lines)
    print(out)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        f.write(out + "\n")


if __name__ == "__main__":
    main()
