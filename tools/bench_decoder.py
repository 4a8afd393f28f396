#!/usr/bin/env python
"""Time the fused HiDDeN decoder (csrc/decoder.cu) forward + backward on the bench shape, as a CUDA-graph replay.

    python tools/bench_decoder.py [--B 32] [--H 12] [--W 12] [--iters 200]

Environment switches read by the library (A/B): NSIG_DEC_NO_SIDE=1 (weight gradients on the caller's stream),
NSIG_DEC_WGRAD_G=<image groups of the weight-gradient grid>."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--H", type=int, default=12)
    ap.add_argument("--W", type=int, default=12)
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    import torch
    from nerf_signature_b200.nerf.hidden_models import get_hidden_decoder_multi_views
    from nerf_signature_b200.nerf.decoder_ops import decode

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    dec = get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=8, input_ch=3, channels=64).to(dev).train()
    for p in dec.parameters():
        p.grad = torch.zeros_like(p)
    img = torch.rand(args.B, args.H, args.W, 3, device=dev, requires_grad=True)
    target = torch.randint(0, 2, (args.B, 1), device=dev).float()

    def step():
        logits = decode(dec, img)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.float() * 10.0, target)
        loss.backward()
        return loss

    def fwd_only():
        with torch.no_grad():
            return decode(dec, img)

    for name, fn in (("forward+backward", step), ("forward only", fwd_only)):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(10):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"decoder {name}: {1e3 * e0.elapsed_time(e1) / args.iters:.1f} us per replay "
              f"(B={args.B}, {args.H}x{args.W}; NO_SIDE={os.environ.get('NSIG_DEC_NO_SIDE', '0')}, "
              f"WGRAD_G={os.environ.get('NSIG_DEC_WGRAD_G', 'default')})", flush=True)


if __name__ == "__main__":
    main()
