"""Multiresolution hash encoder with the reference's interface (hash_encoding.py:48-111).

Same constructor, same `embeddings.{i}.weight` parameters ([2^log2_T, F] fp32, U(-1e-4, 1e-4)),
same level resolutions (floor(base * b**i) evaluated with the reference's torch fp32 expression,
SURVEY F2) and bit-identical outputs — but one CUDA kernel per call (nsig_hash_encode_forward)
instead of ~25 torch kernels and a host sync per level, and a scatter-add kernel for the table
gradients.  No CPU path.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib

_P = _lib.ptr


def level_resolutions(base_resolution, b, n):
    """floor(base_resolution * b**i) for i < n, exactly as hash_encoding.py:100 evaluates it."""
    return [float(torch.floor(base_resolution * b ** i)) for i in range(n)]


class _hash_encode(Function):
    """out[B, 2L] = encode(x; tables).  Differentiable w.r.t. the tables only (positions come from the
    ray marcher and carry no gradient in any of the reference's pipelines)."""

    @staticmethod
    def forward(ctx, x, resolutions, log2_T, *tables):
        x = x.contiguous().float()
        B, L = x.shape[0], len(tables)
        out = torch.empty(B, 2 * L, dtype=torch.float32, device=x.device)
        tabs = [t.contiguous() for t in tables]
        _lib.call("nsig_hash_encode_forward", _P(x), B, _lib.pointer_array(tabs), _lib.float_array(resolutions),
                  L, log2_T, _P(out), None)
        ctx.save_for_backward(x)
        ctx.meta = (resolutions, log2_T, [tuple(t.shape) for t in tables])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        resolutions, log2_T, shapes = ctx.meta
        L = len(shapes)
        need = ctx.needs_input_grad[3:]
        if not any(need):
            return (None,) * (3 + L)
        grad_out = grad_out.contiguous().float()
        grads = [torch.zeros(s, dtype=torch.float32, device=x.device) for s in shapes]
        _lib.call("nsig_hash_encode_backward", _P(x), _P(grad_out), x.shape[0], _lib.pointer_array(grads),
                  _lib.float_array(resolutions), L, log2_T)
        return (None, None, None) + tuple(g if n else None for g, n in zip(grads, need))


class HalfTables:
    """half2 shadow tables + their per-level de-scaling factors (device float[16])."""

    def __init__(self, tables, inv_scale, scratch):
        self.tables, self.inv_scale, self.scratch = tables, inv_scale, scratch


class HashEmbedder(nn.Module):
    def __init__(self, bounding_box, n_levels=16, n_features_per_level=2,
                 log2_hashmap_size=19, base_resolution=16, finest_resolution=512):
        super(HashEmbedder, self).__init__()
        if n_features_per_level != 2:
            raise NotImplementedError("the sm_100a kernels are specialised for 2 features per level")
        if n_levels > 16:
            raise NotImplementedError("at most 16 levels")
        box_min, box_max = bounding_box
        if float(torch.as_tensor(box_min).max()) != 0.0 or float(torch.as_tensor(box_max).min()) != 1.0:
            raise NotImplementedError("bounding_box must be (0, 1), as every reference call site passes it")
        self.bounding_box = bounding_box
        self.n_levels = n_levels
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = torch.tensor(base_resolution)
        self.finest_resolution = torch.tensor(finest_resolution)
        self.out_dim = self.n_levels * self.n_features_per_level

        self.b = torch.exp((torch.log(self.finest_resolution) - torch.log(self.base_resolution)) / (n_levels - 1))

        self.embeddings = nn.ModuleList([nn.Embedding(2 ** self.log2_hashmap_size,
                                                      self.n_features_per_level) for i in range(n_levels)])
        for i in range(n_levels):
            nn.init.uniform_(self.embeddings[i].weight, a=-0.0001, b=0.0001)
        self.resolutions = level_resolutions(self.base_resolution, self.b, n_levels)
        self._shadow = None      # HalfTables for the fused kernels (not part of the state dict)
        self._shadow_key = None

    def tables(self):
        return [e.weight for e in self.embeddings]

    @torch.no_grad()
    def half_tables(self):
        """half2 shadow copies of the level tables for the fused field kernels (nsig_tables_to_half2): one 32-bit
        gather per corner instead of 64.  Rebuilt whenever a table's storage or version changes (once in watermark
        training, where the base encoder is frozen; every optimizer step in clean training, ~96 MB of traffic)."""
        tabs = self.tables()
        key = tuple((t.data_ptr(), t._version) for t in tabs)
        if self._shadow is not None and self._shadow_key == key:
            return self._shadow
        dev = tabs[0].device
        if self._shadow is None or self._shadow.inv_scale.device != dev:
            self._shadow = HalfTables([torch.empty(t.shape[0], 2, dtype=torch.float16, device=dev) for t in tabs],
                                      torch.ones(16, dtype=torch.float32, device=dev),
                                      torch.zeros(16, dtype=torch.int32, device=dev))
        sh = self._shadow
        _lib.call("nsig_tables_to_half2", _lib.pointer_array([t.contiguous() for t in tabs]), len(tabs),
                  self.log2_hashmap_size, _lib.pointer_array(sh.tables), _P(sh.inv_scale), _P(sh.scratch))
        self._shadow_key = key
        return sh

    def forward(self, x):
        # x: B x 3 in the unit box
        return _hash_encode.apply(x, self.resolutions, self.log2_hashmap_size, *self.tables())

    @torch.no_grad()
    def hashed_indices(self, x):
        """int32 [B, n_levels, 8] table slots (hash_encoding.py:43-44), for parity tests."""
        x = x.contiguous().float()
        B = x.shape[0]
        out = torch.empty(B, 2 * self.n_levels, dtype=torch.float32, device=x.device)
        slots = torch.empty(B, self.n_levels, 8, dtype=torch.int32, device=x.device)
        tabs = [t.contiguous() for t in self.tables()]
        _lib.call("nsig_hash_encode_forward", _P(x), B, _lib.pointer_array(tabs), _lib.float_array(self.resolutions),
                  self.n_levels, self.log2_hashmap_size, _P(out), _P(slots))
        return slots

    @torch.no_grad()
    def fused_hashed_indices(self, x, resolutions=None, want_weights=False):
        """int32 [B, L, 8] slots as the FUSED field kernels derive them (csrc/hash_common.cuh locate_fused: one
        double multiply instead of the reference's fp32 division) — must equal `hashed_indices` bit for bit.
        `resolutions` overrides the level list (e.g. [2048.0] for the message encoder's geometry)."""
        x = x.contiguous().float()
        res = list(self.resolutions if resolutions is None else resolutions)
        B, L = x.shape[0], len(res)
        slots = torch.empty(B, L, 8, dtype=torch.int32, device=x.device)
        w = torch.empty(B, L, 3, dtype=torch.float32, device=x.device) if want_weights else None
        _lib.call("nsig_fused_hash_slots", _P(x), B, _lib.float_array(res), L, self.log2_hashmap_size, _P(slots), _P(w))
        return (slots, w) if want_weights else slots
