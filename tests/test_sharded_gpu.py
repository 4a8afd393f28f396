"""-m gpu, needs >= 2 GPUs on one NVLink node (skipped otherwise): N-GPU == 1-GPU semantics of the ray-sharded training step
(SURVEY 8e).  ONE global batch is (a) sharded over 2 ranks - contiguous ray ranges of the content rays and of the
watermark-block rays, block pixels all-gathered before the decoder, one gradient exchange (the one-kernel NVLink all-reduce),
message-table optimizer sharded over the ranks with the summed table all-gathered - and (b) processed whole by a single
process.  Compared after every step: exchanged dL/dS, decoder gradients, and (after gather_tables) the updated tables."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CFG = dict(bound=1.0, scale=0.8, dt_gamma=0.0, message_dim=6, num_rows=32, num_cols=32, H=128, W=128, num_rays=512,
           camera="blender", occupancy="sphere")


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    from nerf_signature_b200 import harness
    kw = dict(seed=0, optimizer="fused", graph=False, merged_render=True, fused_decoder=True, fused_losses=True,
              table_scale=300.0)
    gen = torch.Generator().manual_seed(4)
    res = {"dG": [], "dD": [], "dT": []}
    gb0 = harness.make_batch(CFG, seed=70)
    local, bshape, counts = harness.shard_batch(gb0, rank, ws)
    lcfg = dict(CFG); lcfg["num_rays"] = local["rays_o"].shape[1]
    sn = harness.Scene(lcfg, dev, shard_blocks=(bshape, counts), **kw)
    s1 = harness.Scene(dict(CFG), dev, distributed=False, **kw)
    res["sharded"] = sn.optimizer.shard is not None
    res["exchange"] = sn.sync.exchange
    rel = lambda a, b: float((a - b).norm() / b.norm())
    for step in range(3):
        gb = harness.make_batch(CFG, seed=70 + step)
        local, _, _ = harness.shard_batch(gb, rank, ws)
        msg = torch.randint(0, 2, (CFG["message_dim"],), generator=gen).float()
        scale = sn.scaler.get_scale()
        sn.train_step(sn.to_device(local), msg)
        s1.train_step(s1.to_device(gb), msg)
        Gn, G1 = sn.optimizer.G / scale, s1.optimizer.G / scale
        Dn = torch.cat([p.grad.reshape(-1) for p in sn._decoder_params])
        D1 = torch.cat([p.grad.reshape(-1) for p in s1._decoder_params])
        res["dG"].append(rel(Gn, G1)); res["dD"].append(rel(Dn, D1))
        sn.optimizer.gather_tables()
        tn, t1 = sn.model.msg_encoder.tables(), s1.model.msg_encoder.tables()
        bad = max(((a - b).abs() > 1e-5 + 1e-3 * b.abs()).float().mean().item() for a, b in zip(tn, t1))
        res["dT"].append(bad)
    torch.cuda.synchronize()
    out[rank] = res
    dist.barrier()
    os._exit(0)


def test_two_gpu_sharded_step_equals_single_gpu_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    ctx = mp.spawn(_worker, args=(ws, 29653, out), nprocs=ws, join=False)
    ctx.join(timeout=420)
    assert len(out) == ws, "a rank did not finish"
    for rank in range(ws):
        r = out[rank]
        assert r["sharded"], r
        # step 0 starts from identical parameters: the sharded step IS the single-GPU step up to fp32 summation order
        assert r["dG"][0] < 1e-5 and r["dD"][0] < 1e-4, (rank, r)
        assert r["dT"][0] < 1e-3, (rank, r)      # fraction of table entries off by more than 1e-3 after the first update
        # later steps compare two TRAJECTORIES: Adam's update is sign-like for near-zero gradients (eps = 1e-15), so rounding
        # differences of step 0 move a few entries by ~lr and the runs drift apart like any two runs of the same schedule do
        # (measured on B200: 1.7e-5 after one update, 6e-3 after two)
        assert max(r["dG"]) < 5e-2 and max(r["dD"]) < 5e-2 and max(r["dT"]) < 2e-2, (rank, r)
