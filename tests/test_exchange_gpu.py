"""-m gpu, needs >= 2 GPUs on one NVLink node (skipped otherwise): the one-kernel gradient exchange
(csrc/collective.cu, nsig_allreduce_mean_inplace) against NCCL's all-reduce, with NVSwitch multicast and with plain
P2P, eager and replayed from a CUDA graph (the training step captures it)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, ws, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    from nerf_signature_b200 import parallel
    res = {}
    for mode in ("multicast", "p2p"):
        os.environ["NSIG_AR_NO_MULTICAST"] = "1" if mode == "p2p" else "0"
        b = parallel.SymmetricBucket(n, dev)
        gen = torch.Generator(device=dev).manual_seed(10 + rank)
        errs = []
        for _ in range(3):
            x = torch.randn(b.n, device=dev, generator=gen)
            ref = x.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.AVG)
            b.buf.copy_(x)
            b.all_reduce_mean()
            torch.cuda.synchronize()
            errs.append(float((b.buf - ref).abs().max()))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            b.all_reduce_mean()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            b.all_reduce_mean()
        for _ in range(4):  # the flag exchange resets itself: replays keep working
            x = torch.randn(b.n, device=dev, generator=gen)
            ref = x.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.AVG)
            b.buf.copy_(x)
            g.replay()
            torch.cuda.synchronize()
            errs.append(float((b.buf - ref).abs().max()))
        res[mode] = (max(errs), bool(b.multicast))
    out[rank] = res
    dist.barrier()
    os._exit(0)


def test_one_kernel_allreduce_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    n = (1 << 20) + 262144 + 37        # the watermark bucket: 4 MiB of dL/dS + the decoder, not a multiple of 8
    ctx = mp.spawn(_worker, args=(ws, 29641, n, out), nprocs=ws, join=False)
    ctx.join(timeout=240)
    assert len(out) == ws, "a rank did not finish"
    for rank in range(ws):
        for mode, (err, mc) in out[rank].items():
            assert err <= 1e-6, (rank, mode, err)   # fp32 sum of 2 terms: bit-equal up to the order of one addition
