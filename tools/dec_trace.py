#!/usr/bin/env python
"""Phase timeline of the decoder's conv kernels, from a -DNSIG_DEC_TRACE build (CTA 0 of every conv launch stamps %globaltimer):

    python tools/build_variant.py trace -DNSIG_DEC_TRACE
    NSIG_LIB=tools/scratch/libs/libnsig_trace.so python tools/dec_trace.py

Prints, per conv launch of one forward + backward (graph replay, like the training step), the ns between the phase stamps:
entry | weights cp.async issued | tile loads issued | BN coefficients + barrier | tile transformed | weights arrived |
barrier | act_out written | k-loop done (warp 0) | epilogue done | barrier; and the gap from the previous launch's last stamp."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from nerf_signature_b200 import _lib
    from nerf_signature_b200.nerf.hidden_models import get_hidden_decoder_multi_views
    from nerf_signature_b200.nerf.decoder_ops import decode
    lib = _lib.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    dec = get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=8, input_ch=3, channels=64).to(dev).train()
    for p in dec.parameters():
        p.grad = torch.zeros_like(p)
    img = torch.rand(32, 12, 12, 3, device=dev, requires_grad=True)
    target = torch.randint(0, 2, (32, 1), device=dev).float()

    def step():
        logits = decode(dec, img)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.float() * 10.0, target)
        loss.backward()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    buf = (ctypes.c_ulonglong * (64 * 12))()
    n = ctypes.c_uint(0)
    lib.nsig_debug_dec_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.nsig_debug_dec_trace(buf, ctypes.byref(n), 1)
    g.replay()
    lib.nsig_debug_dec_trace(buf, ctypes.byref(n), 0)
    rows = [[buf[i * 12 + k] for k in range(12)] for i in range(min(n.value, 64))]
    rows.sort(key=lambda r: r[0])
    names = ["w-issue", "ld-issue", "bn+bar", "transform", "w-wait", "barrier", "act_out", "k-loop", "epilogue", "barrier"]
    print(f"{n.value} conv launches; ns per phase (CTA 0, thread 0)")
    print(f"{'#':>3} {'gap':>7} " + " ".join(f"{x:>9}" for x in names) + f" {'total':>8}")
    prev_end = None
    for i, r in enumerate(rows):
        st = [r[k] for k in range(11)]
        d = []
        last = st[0]
        for k in range(1, 11):
            if st[k] == 0:
                d.append(0)
            else:
                d.append(st[k] - last)
                last = st[k]
        gap = (st[0] - prev_end) if prev_end is not None else 0
        prev_end = last
        extra = f"  [stamp 11 at +{r[11] - st[0]}]" if r[11] else ""
        print(f"{i:3d} {gap:7d} " + " ".join(f"{x:9d}" for x in d) + f" {last - st[0]:8d}" + extra)
    wgrad_rows(lib)


def wgrad_rows(lib):
    import ctypes
    buf = (ctypes.c_ulonglong * (64 * 12))()
    n = ctypes.c_uint(0)
    lib.nsig_debug_dec_trace(buf, ctypes.byref(n), 2)
    rows = [[buf[i * 12 + k] for k in range(12)] for i in range(min(n.value, 64))]
    rows.sort(key=lambda r: r[0])
    print(f"{n.value} weight-gradient launches; ns from entry (CTA (tap 0, group 0), thread 0; 7-9: the CTA that reduces tap 0)")
    print("  #  staged0   item0   item1  stored  fenced  ticket | last: start  summed  written")
    for i, r in enumerate(rows):
        rel = [(r[k] - r[0]) if r[k] else 0 for k in range(1, 10)]
        print(f"{i:3d} " + " ".join(f"{x:7d}" for x in rel[:6]) + "  |     " + " ".join(f"{x:7d}" for x in rel[6:]))


if __name__ == "__main__":
    main()
