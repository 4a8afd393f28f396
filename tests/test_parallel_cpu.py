"""CPU tests (-m "not gpu") of the ray-sharding host logic with a real 2-process gloo group:
shard ranges, the flat-bucket gradient all-reduce, the message-table gradient hook and the differentiable
pixel all-gather that keeps the HiDDeN decoder's BatchNorm statistics global."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_signature_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 262144, 4608 + 5):
        for ws in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        sync = parallel.GradSync()
        assert sync.enabled
        # ---- flat bucket over ragged parameter shapes, some without grads ----
        params = [torch.nn.Parameter(torch.zeros(s)) for s in ((3, 5), (7,), (2, 2, 2), (4,))]
        local = []
        for i, p in enumerate(params):
            if i == 3:
                continue  # no grad on any rank: skipped
            p.grad = torch.randn_like(p)
            local.append(p.grad.clone())
        sync.reduce_params(params)
        gathered = [None] * world
        dist.all_gather_object(gathered, [g.tolist() for g in local])
        for i, p in enumerate(params[:3]):
            want = sum(torch.tensor(gathered[r][i]) for r in range(world)) / world
            assert torch.allclose(p.grad, want, atol=1e-6)
        assert params[3].grad is None
        # ---- message-table gradient hook: mean over ranks of dL/dS ----
        g = torch.full((16, 2), float(rank + 1))
        out = sync.reduce_table_grad(g)
        assert torch.allclose(out, torch.full((16, 2), sum(range(1, world + 1)) / world))
        # ---- pixel all-gather with ragged counts, gradient returns the local slice ----
        counts = [5, 3][:world]
        x = (torch.arange(counts[rank] * 3, dtype=torch.float32).view(-1, 3) + 100 * rank).requires_grad_(True)
        full = parallel.all_gather_pixels(x, counts)
        assert full.shape == (sum(counts), 3)
        lo = sum(counts[:rank])
        assert torch.equal(full[lo:lo + counts[rank]], x.detach())
        other = 1 - rank
        olo = sum(counts[:other])
        assert float(full[olo, 0]) == 100.0 * other
        w = torch.arange(sum(counts) * 3, dtype=torch.float32).view(-1, 3)
        (full * w).sum().backward()
        assert torch.equal(x.grad, w[lo:lo + counts[rank]])
        # grad_scale = world: every rank differentiates the FULL-batch loss, the exchange then averages over ranks
        x2 = x.detach().clone().requires_grad_(True)
        (parallel.all_gather_pixels(x2, counts, grad_scale=world) * w).sum().backward()
        assert torch.equal(x2.grad, world * w[lo:lo + counts[rank]])
        # ---- watermark-block sharding: rank-mean of the sharded gradients == single-process gradient ----
        torch.manual_seed(11)
        table = torch.nn.Parameter(torch.randn(6, 3))          # stands in for the message table: pixel = f(table)
        Wd = torch.nn.Parameter(torch.randn(3, 1))             # stands in for the decoder (sees the FULL batch)
        sel = torch.tensor([0, 1, 2, 3, 4, 5, 0, 2])           # 8 "block pixels"
        plo, phi = parallel.shard_range(8, rank, world)
        pcounts = [parallel.shard_range(8, r, world)[1] - parallel.shard_range(8, r, world)[0] for r in range(world)]
        local_px = torch.tanh(table[sel[plo:phi]])
        full_px = parallel.all_gather_pixels(local_px, pcounts, grad_scale=world)
        bn = (full_px - full_px.mean(0)) / full_px.std(0)       # batch statistics over ALL pixels, like BatchNorm
        (bn @ Wd).pow(2).mean().backward()
        sync.reduce_params([table, Wd])
        t1, w1 = torch.nn.Parameter(table.detach().clone()), torch.nn.Parameter(Wd.detach().clone())
        px = torch.tanh(t1[sel])
        bn1 = (px - px.mean(0)) / px.std(0)
        (bn1 @ w1).pow(2).mean().backward()
        assert torch.allclose(table.grad, t1.grad, atol=1e-5) and torch.allclose(Wd.grad, w1.grad, atol=1e-5)
        # ---- equivalence: sharded mean-loss gradients == single-process gradients ----
        torch.manual_seed(7)
        W = torch.nn.Parameter(torch.randn(3, 4))
        rays = torch.randn(10, 4)
        lo, hi = parallel.shard_range(10, rank, world)
        loss = ((rays[lo:hi] @ W.t()) ** 2).sum() / 10 * world   # local mean scaled so that the rank-mean is global
        loss.backward()
        sync.reduce_params([W])
        W2 = torch.nn.Parameter(W.detach().clone())
        ((rays @ W2.t()) ** 2).sum().div(10).backward()
        assert torch.allclose(W.grad, W2.grad, atol=1e-5)
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gloo_world_size_2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}


def _flat_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)
        sync = parallel.GradSync()
        params = [torch.nn.Parameter(torch.randn(s)) for s in ((4, 3), (5,), (2, 2))]
        G = sync.make_flat_buffer(8, params, torch.device("cpu"))
        assert G.numel() == 8 and sync.flat.numel() == 8 + 12 + 5 + 4
        # gradients accumulate straight into the bucket (.grad are views of it)
        G.copy_(torch.full((8,), float(rank + 1)))
        loss = sum((p * (rank + 1)).sum() for p in params)
        loss.backward()
        assert all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in params)
        # gloo has no AVG: emulate it the way reduce_flat does on NCCL (SUM then divide) to check the layout
        dist.all_reduce(sync.flat, op=dist.ReduceOp.SUM)
        sync.flat.div_(world)
        mean = sum(range(1, world + 1)) / world
        assert torch.allclose(G, torch.full((8,), mean))
        for p in params:
            assert torch.allclose(p.grad, torch.full_like(p, mean))
        sync.zero_flat()
        # one fill clears the whole bucket: dL/dS is scatter-ADDED by the field backward, like the other gradients
        assert float(G.abs().sum()) == 0 and all(float(p.grad.abs().sum()) == 0 for p in params)
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_flat_gradient_bucket_gloo_world_size_2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_flat_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}


def _decoder_worker(rank, world, port, results):
    """SURVEY 8(e) with the REAL HiDDeN decoder (BatchNorm on batch statistics): a 'table' produces the block pixels; rank r
    renders the contiguous slice r of the flattened block rays (harness.shard_batch's partitioning), the pixels are
    all-gathered, every rank decodes the FULL batch, and after the rank-mean exchange the table and decoder gradients equal
    the single-process ones.  Without the all-gather (each rank decoding its own blocks) they would not."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch.nn.functional as F
        from nerf_signature_b200.nerf.hidden_models import get_hidden_decoder_multi_views, normalize_img
        torch.set_num_threads(1)
        md, ph, pw = 6, 5, 4
        nb = md * ph * pw                                            # 120 block rays; 2 ranks: 60 each = 3 whole blocks
        g = torch.Generator().manual_seed(5)
        table0 = torch.rand(nb, 3, generator=g)                      # stands in for the message tables: pixel = sigmoid(table)
        message = torch.randint(0, 2, (md,), generator=g).float()
        gt = torch.rand(50, 3, generator=g)
        content_w0 = torch.rand(50, 3, generator=g)                  # a content-ray path that also depends on "the tables"

        def make():
            torch.manual_seed(9)
            dec = get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=3, input_ch=3, channels=16)
            return torch.nn.Parameter(table0.clone()), torch.nn.Parameter(content_w0.clone()), dec

        def loss_fn(dec, pixels_full, content_pred, n_content_total, content_gt):
            pred = pixels_full.view(md, ph, pw, 3).clamp(0, 1)
            decoded = dec(normalize_img(pred.permute(0, 3, 1, 2)))
            lossw = F.binary_cross_entropy_with_logits(decoded * 10.0, message.unsqueeze(-1), reduction="mean")
            lossi_sum = F.mse_loss(content_pred, content_gt, reduction="none").sum() / (3 * n_content_total)
            return lossw, lossi_sum

        # ---- single process ----
        t1, c1, d1 = make()
        lw, li = loss_fn(d1, torch.sigmoid(t1), torch.sigmoid(c1), 50, gt)
        (0.005 * lw + 1.0 * li).backward()

        # ---- sharded: contiguous ranges of the block rays and of the content rays ----
        t2, c2, d2 = make()
        blo, bhi = parallel.shard_range(nb, rank, world)
        clo, chi = parallel.shard_range(50, rank, world)
        counts = [parallel.shard_range(nb, r, world)[1] - parallel.shard_range(nb, r, world)[0] for r in range(world)]
        local_px = torch.sigmoid(t2[blo:bhi])
        full_px = parallel.all_gather_pixels(local_px, counts, grad_scale=world)
        lw2, li2 = loss_fn(d2, full_px, torch.sigmoid(c2[clo:chi]), 50, gt[clo:chi])
        # every rank holds the FULL watermark loss (its table gradient slice x world) and its share of the content loss x world
        (0.005 * lw2 + 1.0 * li2 * world).backward()
        sync = parallel.GradSync()
        sync.reduce_params([t2, c2] + list(d2.parameters()))
        assert torch.allclose(lw2.detach(), lw.detach(), atol=1e-6)
        assert torch.allclose(t2.grad, t1.grad, atol=1e-7, rtol=1e-4), float((t2.grad - t1.grad).abs().max())
        assert torch.allclose(c2.grad, c1.grad, atol=1e-8, rtol=1e-5)
        for pa, pb in zip(d2.parameters(), d1.parameters()):
            assert torch.allclose(pa.grad, pb.grad, atol=1e-6, rtol=1e-3), float((pa.grad - pb.grad).abs().max())

        # ---- what the all-gather is for: per-rank decoding (local BatchNorm statistics) gives a different gradient ----
        t3, _, d3 = make()
        n_loc = (bhi - blo) // (ph * pw)
        pred3 = torch.sigmoid(t3[blo:bhi]).view(n_loc, ph, pw, 3)
        dec3 = d3(normalize_img(pred3.permute(0, 3, 1, 2)))
        mlo = blo // (ph * pw)
        F.binary_cross_entropy_with_logits(dec3 * 10.0, message[mlo:mlo + n_loc].unsqueeze(-1), reduction="mean").backward()
        sync.reduce_params([t3])
        assert (t3.grad * 0.005 - t1.grad).abs().max() > 1e-2 * t1.grad.abs().max()
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_sharded_blocks_with_the_real_decoder_gloo_world_size_2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_decoder_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}
