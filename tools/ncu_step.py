#!/usr/bin/env python
"""One eager watermark training step between cudaProfilerStart/Stop, for
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^k_ -o OUT python tools/ncu_step.py
(every library kernel of the step once; numbers taken under the profiler are never bench values)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from nerf_signature_b200 import harness
    dev = torch.device("cuda:0")
    cfg = dict(harness.CONFIGS[os.environ.get("NSIG_CONFIG", "blender_wtmk")])
    scene = harness.Scene(cfg, dev, seed=0, merged_render=True, fused_decoder=True, fused_losses=True)
    batches = [scene.to_device(harness.make_batch(cfg, seed=i)) for i in range(2)]
    gen = torch.Generator().manual_seed(7)
    for i in range(4):
        scene.train_step(batches[i % 2], scene.new_message(gen))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    scene.train_step(batches[0], scene.new_message(gen))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
