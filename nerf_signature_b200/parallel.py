"""Ray-sharded data parallelism (new functionality: the reference has no multi-GPU hot path,
SURVEY.md 2.4 / 8e).  One process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

Rays are independent through march -> field -> composite, parameters are replicated, so the only
exchange is the gradient reduction once per step:
  * the message-table gradient: every selected table's gradient equals dL/dS, so ONE
    [2^19, 2] fp32 tensor (4 MiB) is all-reduced before it fans out to the message_dim selected tables
    (hook called from hash_encoding_wtmk_bit._msg_table_sum.backward) instead of message_dim tensors;
  * all remaining trainable parameters (HiDDeN decoder, ~262 k values; base tables + MLPs in clean
    mode) as one flat bucket.
For a watermark batch sharded across ranks the rendered block pixels are all-gathered before the
decoder (its BatchNorm uses batch statistics over the blocks): `all_gather_pixels`.
"""
import os
import sys

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous slice [lo, hi) of n units owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class SymmetricBucket:
    """A flat fp32 gradient bucket in peer-mapped (symmetric) memory with its one-kernel mean all-reduce
    (csrc/collective.cu: nsig_allreduce_mean_inplace - NVSwitch multicast when the fabric offers it, plain P2P
    otherwise).  Raises when symmetric memory cannot be set up; the caller then stays on NCCL."""

    def __init__(self, n, device, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        rank, ws = world()
        group = group if group is not None else dist.group.WORLD
        self.n = (n + 4 * ws - 1) // (4 * ws) * (4 * ws)
        self.buf = symm_mem.empty(self.n, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group=group)
        grid = int(_lib.load().nsig_allreduce_grid())
        self.flags = symm_mem.empty(grid * ws, dtype=torch.int32, device=device)
        self.fhdl = symm_mem.rendezvous(self.flags, group=group)
        self.buf.zero_()
        self.flags.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # nobody raises a flag before every flag buffer is zero
        self._bufs = (ctypes.c_void_p * ws)(*[int(p) for p in self.hdl.buffer_ptrs])
        self._flags = (ctypes.c_void_p * ws)(*[int(p) for p in self.fhdl.buffer_ptrs])
        mc = int(self.hdl.multicast_ptr) if getattr(self.hdl, "has_multicast_support", False) else 0
        if os.environ.get("NSIG_AR_NO_MULTICAST") == "1":
            mc = 0
        self.multicast = mc
        self.rank, self.world = rank, ws

    def all_reduce_mean(self):
        from . import _lib
        _lib.call("nsig_allreduce_mean_inplace", self._bufs, self._flags, self.multicast or None, self.n, self.rank,
                  self.world)


class GradSync:
    """Averages gradients across ranks once per step."""

    def __init__(self, group=None):
        self.group = group
        self.enabled = world()[1] > 1
        self.bucket = None        # SymmetricBucket when the one-kernel exchange is in use
        self.exchange = "none" if not self.enabled else "nccl"

    # --- message-table path: called inside autograd with dL/dS -------------------------------------
    def reduce_table_grad(self, grad_S):
        if not self.enabled:
            return grad_S
        g = grad_S.contiguous()
        # stream-ordered (no host wait): autograd consumes g on the compute stream right after this call
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        g.div_(world()[1])
        return g

    # --- fused path: ONE all-reduce per step over a persistent flat gradient buffer -------------------
    def make_flat_buffer(self, table_numel, params, device):
        """Allocate [dL/dS | every other trainable gradient] as one fp32 buffer and point each parameter's .grad
        at its slice, so backward accumulates straight into the bucket and a single NCCL all-reduce (AVG)
        per step synchronises everything (no per-tensor copies, no separate division)."""
        n = table_numel + sum(p.numel() for p in params)
        self.flat = None
        if (self.enabled and dist.get_backend(self.group) == "nccl" and torch.device(device).type == "cuda"
                and os.environ.get("NSIG_AR_NCCL") != "1"):
            try:  # peer-memory bucket; NCCL remains the exchange when the platform refuses symmetric memory
                self.bucket = SymmetricBucket(n, device, self.group)
                self.flat = self.bucket.buf[:n]
                self.exchange = "nvls-multicast kernel" if self.bucket.multicast else "p2p kernel"
            except Exception as e:  # noqa: BLE001
                self.bucket = None
                if world()[0] == 0:
                    print(f"[nsig] symmetric-memory bucket unavailable ({type(e).__name__}: {str(e)[:120]}); "
                          "gradient exchange stays on NCCL", file=sys.stderr, flush=True)
        if self.flat is None:
            self.flat = torch.zeros(n, dtype=torch.float32, device=device)
        self.table_numel = table_numel
        o = table_numel
        for p in params:
            # same memory layout as the parameter (e.g. channels_last conv weights): torch's fused Adam insists on it
            p.grad = self.flat[o:o + p.numel()].as_strided(p.shape, p.stride())
            o += p.numel()
        return self.flat[:table_numel]

    def zero_flat(self):
        self.flat.zero_()  # one fill: dL/dS (scatter-added by the field backward) and the decoder gradients both accumulate

    def zero_decoder_grads(self):
        self.flat[self.table_numel:].zero_()

    def zero_table_grad(self):
        self.flat[:self.table_numel].zero_()

    def reduce_flat(self):
        if not self.enabled or os.environ.get("NSIG_DIAG_SKIP_REDUCE") == "1":  # diagnosis only: wrong gradients
            return
        if self.bucket is not None:
            self.bucket.all_reduce_mean()
        elif dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo (CPU tests) has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(world()[1])

    # --- everything else: one flat bucket after backward ---------------------------------------------
    def reduce_params(self, params):
        """All-reduce (mean) the .grad of `params` through a single flat buffer."""
        ws = world()[1]
        if not self.enabled:
            return
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(ws)
        o = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[o:o + n].view_as(g))
            o += n


def all_gather_pixels(local_pixels, counts, group=None, grad_scale=1.0):
    """Concatenate per-rank [n_r, 3] pixel blocks (n_r = counts[r]) into the full [sum, 3] tensor on every
    rank, differentiably: the backward pass hands each rank the gradient slice of its own pixels, multiplied by
    `grad_scale`.  Every rank evaluates the decoder loss on the FULL gathered batch, so the slice already is the complete
    d(loss)/d(own pixels); the step's gradient exchange averages over ranks, hence callers pass grad_scale = world size
    to make the rank-mean of the table gradient equal the single-GPU gradient."""
    return _AllGatherPixels.apply(local_pixels, tuple(int(c) for c in counts), group, float(grad_scale))


class _AllGatherPixels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, counts, group, grad_scale):
        rank, ws = world()
        ctx.counts, ctx.rank, ctx.grad_scale = counts, rank, grad_scale
        if ws == 1:
            return x.clone()
        mx = max(counts)
        pad = torch.zeros(mx, *x.shape[1:], dtype=x.dtype, device=x.device)
        pad[:x.shape[0]] = x
        full = torch.empty(ws * mx, *x.shape[1:], dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(full, pad, group=group)   # one collective, capturable in the step graph
        if all(c == mx for c in counts):
            return full
        return torch.cat([full[r * mx:r * mx + c] for r, c in enumerate(counts)], dim=0)

    @staticmethod
    def backward(ctx, g):
        lo = sum(ctx.counts[:ctx.rank])
        gl = g[lo:lo + ctx.counts[ctx.rank]]
        gl = gl * ctx.grad_scale if ctx.grad_scale != 1.0 else gl.contiguous()
        return gl, None, None, None
