"""2+ GPU check of the one-kernel gradient exchange against NCCL: correctness, eager and graph-replay timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
rank, lr, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from nerf_signature_b200 import parallel
n = (1 << 20) + 262144 + 37
def timeit(fn, iters=50):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for mode in (["mc", "p2p"]):
    os.environ["NSIG_AR_NO_MULTICAST"] = "1" if mode == "p2p" else "0"
    try:
        b = parallel.SymmetricBucket(n, dev)
    except Exception as e:
        print(rank, mode, "bucket failed:", repr(e)[:300], flush=True); continue
    if rank == 0: print(mode, "multicast ptr", hex(b.multicast), "n", b.n, flush=True)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    ok = True
    for trial in range(3):
        x = torch.randn(b.n, device=dev, generator=g)
        ref = x.clone(); dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        b.buf.copy_(x); b.all_reduce_mean(); torch.cuda.synchronize()
        err = float((b.buf - ref).abs().max())
        ok &= err < 1e-6
    t_k = timeit(b.all_reduce_mean)
    y = torch.randn(b.n, device=dev)
    t_n = timeit(lambda: dist.all_reduce(y, op=dist.ReduceOp.AVG))
    # graph replay
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        b.all_reduce_mean()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        b.all_reduce_mean()
    t_g = timeit(gr.replay)
    x = torch.randn(b.n, device=dev, generator=g); ref = x.clone(); dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    b.buf.copy_(x); gr.replay(); torch.cuda.synchronize()
    okg = float((b.buf - ref).abs().max()) < 1e-6
    print(f"rank {rank} mode {mode}: correct {ok} graph-correct {okg} kernel {t_k:.1f} us, graph replay {t_g:.1f} us, nccl {t_n:.1f} us", flush=True)
dist.barrier()
os._exit(0)
