#!/usr/bin/env python
"""Where the branches of the captured training step actually start, WITHOUT a profiler: every C-ABI call of the step is
bracketed by external CUDA events (event-record nodes of the graph, _lib.timing_enable) and the offsets of their start /
end from the first kernel of the replay are printed, averaged over a few replays.  (Event nodes add a little latency per
call; compare offsets, not the total.)

    python tools/graph_offsets.py [--config blender_wtmk] [--replays 10] [--out gpurun_out/graph_offsets.txt]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="blender_wtmk")
    ap.add_argument("--replays", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/graph_offsets.txt")
    ap.add_argument("--names", default="")
    ap.add_argument("--march-ahead", type=int, default=0, help="1 = harness.Scene(march_ahead=True); offsets of graph 0")
    args = ap.parse_args()
    import torch
    from nerf_signature_b200 import _lib, harness

    dev = torch.device("cuda:0")
    cfg = dict(harness.CONFIGS[args.config])
    scene = harness.Scene(cfg, dev, seed=0, optimizer="fused", graph=True, merged_render=True, fused_decoder=True,
                          fused_losses=True, defer_optimizer=True, march_ahead=bool(args.march_ahead))
    batches = [scene.to_device(harness.make_batch(cfg, seed=i)) for i in range(2)]
    gen = torch.Generator().manual_seed(7)
    names = args.names.split(",") if args.names else list(_lib._SIGNATURES)
    _lib.timing_enable(names)
    msgs = [scene.new_message(gen) for _ in range(2 * args.replays + 8)]

    def step(i):
        if args.march_ahead:
            scene.train_step(batches[i % 2], msgs[i], next_batch=batches[(i + 1) % 2], next_message=msgs[i + 1])
        else:
            scene.train_step(batches[i % 2], msgs[i])

    for i in range(6):
        step(i)
    torch.cuda.synchronize()
    evs = [e[:3] for e in _lib._timing_events if e[3] in (None, 0)]
    acc = [[0.0, 0.0] for _ in evs]
    for r in range(args.replays):
        step(6 + 2 * r)            # march_ahead: graph 0 (even steps) ...
        torch.cuda.synchronize()
        first = min(range(len(evs)), key=lambda k: -evs[k][1].elapsed_time(evs[0][1]))   # earliest start
        ref = evs[first][1]
        for k, (name, e0, e1) in enumerate(evs):
            acc[k][0] += ref.elapsed_time(e0) * 1e3
            acc[k][1] += ref.elapsed_time(e1) * 1e3
        step(7 + 2 * r)            # ... graph 1 in between
    rows = sorted(((a[0] / args.replays, a[1] / args.replays, evs[k][0]) for k, a in enumerate(acc)))
    lines = [f"config {args.config}: {len(evs)} bracketed calls, mean of {args.replays} replays (us from the first call's start)",
             f"{'start':>9} {'end':>9} {'dur':>8}  call"]
    for s, e, n in rows:
        lines.append(f"{s:9.1f} {e:9.1f} {e - s:8.1f}  {n}")
    out = "\n".join(lines)
    print(out)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        f.write(out + "\n")


if __name__ == "__main__":
    main()
