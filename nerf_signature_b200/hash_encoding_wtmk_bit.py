"""Watermark-bit hash encoder with the reference's interface (hash_encoding_wtmk_bit.py:51-116).

The reference evaluates, for every message bit i, a full gather + trilinear interpolation from table
`2*i + bit_i` and sums the message_dim results.  All "levels" share one resolution
(base == finest == 2048 at the only call site, nerf/network_wtmk_tcnn.py:43-44), hence
identical hash slots and weights, so
    forward(x, message) == trilerp(S)[x],   S = sum_i embeddings[2*i + bit_i].weight
and d(loss)/d(embeddings[2*i + bit_i].weight) == d(loss)/dS for every selected table.
This module builds S with one streaming kernel (message read on the device: no .item() per bit) and
gathers 8 corners per sample instead of 8*message_dim.  Parameters stay 2*message_dim separate
`embeddings.{i}.weight` tensors, unselected tables keep `grad is None` exactly like the reference.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib
from .hash_encoding import _hash_encode, level_resolutions

_P = _lib.ptr

# Optional callable applied to dL/dS before it fans out to the selected tables; the data-parallel
# harness installs parallel.GradSync.reduce_table_grad here (one 4 MiB all-reduce instead of message_dim).
grad_reducer = None


def message_bits(message):
    """Host copy of the bits as a tuple of ints (one D2H transfer; the reference does one per bit)."""
    if isinstance(message, torch.Tensor):
        return tuple(int(v) for v in message.detach().to("cpu", torch.float32).tolist())
    return tuple(int(v) for v in message)


class _msg_table_sum(Function):
    """S = sum_i tables[2i + bit_i]; backward hands dS to every selected table."""

    @staticmethod
    def forward(ctx, message, bits, log2_T, *tables):
        md = len(bits)
        dev = tables[0].device
        msg = message.to(device=dev, dtype=torch.float32).contiguous()
        S = torch.empty_like(tables[0])
        tabs = [t.contiguous() for t in tables]
        _lib.call("nsig_msg_table_sum", _lib.pointer_array(tabs), md, _P(msg), log2_T, _P(S), 0, 0)
        ctx.bits = bits
        ctx.n = len(tables)
        return S

    @staticmethod
    def backward(ctx, grad_S):
        grads = [None] * ctx.n
        if grad_reducer is not None:
            grad_S = grad_reducer(grad_S)
        for i, b in enumerate(ctx.bits):
            if ctx.needs_input_grad[3 + 2 * i + b]:
                grads[2 * i + b] = grad_S
        return (None, None, None) + tuple(grads)


class _msg_table_sum_sink(Function):
    """S = sum_i tables[2i + bit_i] with the bits read on the device, for the fused optimizer path
    (optim.WatermarkAdam): backward deposits dL/dS into the persistent buffer `sink` instead of
    materialising message_dim identical table gradients, so nothing here depends on host-side knowledge of
    the message — the whole training step can be captured in a CUDA graph."""

    @staticmethod
    def forward(ctx, message, log2_T, sink, shard, pre, *tables):
        md = len(tables) // 2
        msg = message.to(device=tables[0].device, dtype=torch.float32).contiguous()
        # pre: S already computed for exactly this message by the optimizer's look-ahead (optim.WatermarkAdam.lookahead_sum;
        # with a sharded optimizer only this rank's slice of it): no kernel here
        S = pre.view_as(pre) if pre is not None else torch.empty_like(tables[0])
        tabs = [t.contiguous() for t in tables]
        if shard is None:
            if pre is None:
                _lib.call("nsig_msg_table_sum", _lib.pointer_array(tabs), md, _P(msg), log2_T, _P(S), 0, 0)
        else:
            # sharded optimizer state (optim.WatermarkAdam(shard=...)): this rank holds the up-to-date values of its own
            # slice of every table only, so it sums that slice and the slices of S are all-gathered (4 MiB in total)
            import torch.distributed as dist
            lo, n, group = shard
            if pre is None:
                _lib.call("nsig_msg_table_sum", _lib.pointer_array(tabs), md, _P(msg), log2_T, _P(S), lo, n)
            flat = S.view(-1)
            dist.all_gather_into_tensor(flat, flat[lo:lo + n].clone(), group=group)
        ctx.sink = sink
        ctx.n = len(tables)
        # the fused field backward scatter-adds dL/dS straight into `sink` (FieldConfig.S_sink) and returns no gradient
        # for S: this node then receives None, which must NOT be materialised as zeros and copied over the sink
        ctx.set_materialize_grads(False)
        return S

    @staticmethod
    def backward(ctx, grad_S):
        if grad_S is not None:  # a consumer that returned dL/dS the ordinary way (e.g. the stand-alone encoder)
            if grad_reducer is not None:
                grad_S = grad_reducer(grad_S)
            ctx.sink.add_(grad_S)
        return (None, None, None, None, None) + (None,) * ctx.n


class HashEmbedder(nn.Module):
    def __init__(self, bounding_box, n_levels=16, n_features_per_level=2,
                 log2_hashmap_size=19, base_resolution=16, finest_resolution=512, message_dim=16):
        super(HashEmbedder, self).__init__()
        if n_features_per_level != 2:
            raise NotImplementedError("the sm_100a kernels are specialised for 2 features per level")
        self.bounding_box = bounding_box
        self.n_levels = n_levels
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = torch.tensor(base_resolution)
        self.finest_resolution = torch.tensor(finest_resolution)
        self.out_dim = self.n_levels * self.n_features_per_level

        self.b = torch.exp((torch.log(self.finest_resolution) - torch.log(self.base_resolution)) / (n_levels - 1))
        self.message_dim = message_dim
        if n_levels < 2 * message_dim:
            raise ValueError("need 2*message_dim tables (embeddings[2*i + bit])")
        self.embeddings = nn.ModuleList([nn.Embedding(2 ** self.log2_hashmap_size,
                                                      self.n_features_per_level) for i in range(n_levels)])
        for i in range(n_levels):
            nn.init.uniform_(self.embeddings[i].weight, a=-0.0001, b=0.0001)
        res = level_resolutions(self.base_resolution, self.b, message_dim)
        if any(r != res[0] for r in res):
            raise NotImplementedError(
                "the pre-summed form needs one resolution for every bit (base_resolution == finest_resolution, "
                "as in nerf/network_wtmk_tcnn.py:43-44)")
        self.resolution = res[0]
        # set by optim.WatermarkAdam: persistent [T,2] buffer that receives dL/dS (see _msg_table_sum_sink)
        self.grad_sink = None
        # set by optim.WatermarkAdam(shard=...): (first float, float count, process group) of the table slice this rank owns
        self.shard = None
        # set by optim.WatermarkAdam.lookahead_sum: (message tensor, S) - consumed by the next summed_table(message) call
        self.presummed = None

    def tables(self):
        return [e.weight for e in self.embeddings[:2 * self.message_dim]]

    def summed_table(self, message, bits=None):
        """S [T,2] (differentiable w.r.t. the selected tables)."""
        if self.shard is not None or (self.grad_sink is not None and torch.is_grad_enabled()):
            # (sharded optimizer: only the owner's slice of a table is current, so S is ALWAYS built slice-wise + all-gather,
            # also for evaluation and the occupancy update - a collective every rank must enter)
            if message.shape[0] != self.message_dim:
                raise ValueError(f"message has {message.shape[0]} bits, encoder was built for {self.message_dim}")
            pre = None
            if self.presummed is not None:
                if self.presummed[0] is message:
                    pre = self.presummed[1]
                self.presummed = None      # one use; a different message means the look-ahead does not apply
            return _msg_table_sum_sink.apply(message, self.log2_hashmap_size, self.grad_sink, self.shard, pre, *self.tables())
        if bits is None:
            bits = message_bits(message)
        if len(bits) != self.message_dim:
            raise ValueError(f"message has {len(bits)} bits, encoder was built for {self.message_dim}")
        if not isinstance(message, torch.Tensor):
            message = torch.tensor(bits, dtype=torch.float32)
        return _msg_table_sum.apply(message, bits, self.log2_hashmap_size, *self.tables())

    def forward(self, x, message):
        S = self.summed_table(message)
        return _hash_encode.apply(x, [self.resolution], self.log2_hashmap_size, S)

    @torch.no_grad()
    def forward_perbit(self, x, message):
        """The reference's literal per-bit evaluation order (for parity tests)."""
        x = x.contiguous().float()
        msg = message.to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty(x.shape[0], 2, dtype=torch.float32, device=x.device)
        tabs = [t.contiguous() for t in self.tables()]
        _lib.call("nsig_msg_encode_forward_perbit", _P(x), x.shape[0], _lib.pointer_array(tabs), self.message_dim,
                  _P(msg), self.resolution, self.log2_hashmap_size, _P(out))
        return out
