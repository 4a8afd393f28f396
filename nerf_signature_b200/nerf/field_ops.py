"""Host side of the fused field kernels (csrc/field.cu): parameter containers that stand in for the
tiny-cuda-nn modules the reference instantiates, and the autograd functions that launch
nsig_field_forward / nsig_field_density / nsig_field_backward.
"""
import math
import os

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib

_P = _lib.ptr

# Watermark-mode backward (only dL/dS is needed), three interchangeable kernels (A/B switch NSIG_BWD for tools and tests):
#   "masks" (default): the forward saves the ReLU sign masks (32 B/sample) and the backward runs the five dgrad GEMMs only;
#   "recompute":       the forward saves the fp16 encoder output (64 B/sample), the backward recomputes the MLPs (mma.sync);
#   "tc":              same data flow as "recompute" on tcgen05.mma + TMEM (csrc/field_tc.cu)
#   "tc_masks":        same data flow as "masks" on tcgen05.mma + TMEM (csrc/field_tc.cu)
BACKWARD_MODE = os.environ.get("NSIG_BWD", "tc" if os.environ.get("NSIG_BWD_TC", "0") == "1" else "masks")


class FusedMLP(nn.Module):
    """Parameter container replacing `tcnn.Network(otype=FullyFusedMLP, activation=ReLU,
    output_activation=None, n_neurons=64)` (nerf/network_wtmk_tcnn.py:52-62,78-88).

    `params` is one flat fp32 vector like tcnn's (state-dict key `<name>.params`): the bias-free
    weight matrices in layer order, each row-major [out, in], with the input width padded to a multiple
    of 16 and the output width padded to 16.  (tcnn's exact flat layout is not verifiable offline —
    SURVEY.md 8c; this is the documented layout of this implementation.)  Initialisation is Xavier
    uniform per matrix, tcnn's default.  The kernels consume an fp16 copy, refreshed when `params`
    changes; arithmetic is fp16 operands with fp32 accumulation.
    """

    def __init__(self, n_input_dims, n_output_dims, n_neurons=64, n_hidden_layers=1, seed=1337):
        super().__init__()
        if n_neurons != 64:
            raise NotImplementedError("the fused kernels are specialised for 64 neurons")
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.n_neurons = n_neurons
        self.n_hidden_layers = n_hidden_layers
        self.padded_input = (n_input_dims + 15) // 16 * 16
        self.padded_output = (n_output_dims + 15) // 16 * 16
        shapes = [(n_neurons, self.padded_input)]
        shapes += [(n_neurons, n_neurons)] * (n_hidden_layers - 1)
        shapes += [(self.padded_output, n_neurons)]
        self.shapes = shapes
        gen = torch.Generator().manual_seed(seed)
        chunks = []
        for (fo, fi) in shapes:
            s = math.sqrt(6.0 / (fi + fo))
            chunks.append(((torch.rand(fo * fi, generator=gen) * 2 - 1) * s))
        self.params = nn.Parameter(torch.cat(chunks).float())
        self._half = None
        self._half_key = None

    def matrices(self, params=None):
        """List of [out, in] views of the flat vector (fp32)."""
        p = self.params if params is None else params
        out, o = [], 0
        for (fo, fi) in self.shapes:
            out.append(p[o:o + fo * fi].view(fo, fi))
            o += fo * fi
        return out

    def half_weights(self):
        key = (self.params.data_ptr(), self.params._version, self.params.device)
        if self._half is None or self._half_key != key:
            self._half = self.params.detach().to(torch.float16).contiguous()
            self._half_key = key
        return self._half

    def forward(self, x):  # pragma: no cover - the networks call the fused kernels instead
        raise NotImplementedError("FusedMLP is evaluated inside the fused field kernels; call NeRFNetwork.forward/"
                                  "density/color")


class SHEncoding(nn.Module):
    """Stands in for `tcnn.Encoding(3, SphericalHarmonics degree 4)` (network_wtmk_tcnn.py:68-74); the
    16 basis values are computed inside the fused kernels (formulas of hash_encoding.py:162-193)."""

    def __init__(self, n_input_dims=3, degree=4):
        super().__init__()
        if n_input_dims != 3 or degree != 4:
            raise NotImplementedError("SH degree 4 on 3-D directions only")
        self.n_input_dims = n_input_dims
        self.degree = degree
        self.n_output_dims = degree ** 2


class FieldConfig:
    """Everything the field kernels need besides tensors."""

    def __init__(self, bound, resolutions, log2_T, msg_resolution, density_scale=1.0, shadow=None, S_sink=None):
        self.shadow = shadow  # hash_encoding.HalfTables or None (gather the fp32 tables)
        # persistent [T,2] fp32 buffer that accumulates dL/dS directly (optim.WatermarkAdam's G, zeroed by its
        # zero_grad): the backward kernel scatter-adds into it and no gradient is returned for S
        self.S_sink = S_sink
        self.bound = float(bound)
        self.resolutions = list(resolutions)
        self.log2_T = int(log2_T)
        self.msg_resolution = float(msg_resolution)
        self.density_scale = float(density_scale)


def _shadow_args(cfg):
    """(tables_h2, h2_inv_scale) arguments of the fused entry points."""
    if cfg.shadow is None:
        return None, None
    return _lib.pointer_array(cfg.shadow.tables), _P(cfg.shadow.inv_scale)


class _field_forward(Function):
    """(sigmas [M], rgbs [M,3]) = field(xyzs, dirs; S, base tables, MLP weights).

    Differentiable w.r.t. S (the pre-summed message table), the base tables and both MLPs' flat weight vectors
    (clean-model training; in watermark training everything but S is frozen, SURVEY F13, and the backward
    kernel then runs its dgrad-only variant).  `count` is an optional device int32 holding the live row count
    (the march counter).
    """
    N_FIXED = 9  # inputs before *tables

    @staticmethod
    def forward(ctx, xyzs, dirs, S, count, cfg, sigma_mlp, color_mlp, sigma_params, color_params, *tables):
        xyzs = xyzs.contiguous().float()
        dirs = dirs.contiguous().float()
        M = xyzs.shape[0]
        dev = xyzs.device
        sigmas = torch.empty(M, dtype=torch.float32, device=dev)
        rgbs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        need_S = S is not None and ctx.needs_input_grad[2]
        need_w = ctx.needs_input_grad[7] or ctx.needs_input_grad[8]
        need_tab = any(ctx.needs_input_grad[_field_forward.N_FIXED:])
        save = need_S or need_tab or need_w
        use_masks = need_S and not (need_w or need_tab) and BACKWARD_MODE in ("masks", "tc_masks")
        feat = torch.empty(M, 32, dtype=torch.float16, device=dev) if (save and not use_masks) else None
        masks = torch.empty(M, 4, 2, dtype=torch.int32, device=dev) if use_masks else None
        tabs = [t.contiguous() for t in tables]
        sw, cw = sigma_mlp.half_weights(), color_mlp.half_weights()
        Sc = S.contiguous() if S is not None else None
        _lib.call("nsig_field_forward", _P(xyzs), _P(dirs), M, cfg.bound, _lib.pointer_array(tabs),
                  _lib.float_array(cfg.resolutions), cfg.log2_T, _P(Sc), cfg.msg_resolution, _P(sw), _P(cw),
                  cfg.density_scale, _P(count), _P(sigmas), _P(rgbs), _P(feat), _P(masks), *_shadow_args(cfg))
        ctx.use_masks = use_masks
        if use_masks:
            ctx.save_for_backward(xyzs, masks, sigmas, rgbs, sw, cw)
            ctx.cfg, ctx.count, ctx.S_shape = cfg, count, tuple(S.shape)
        elif save:
            ctx.save_for_backward(xyzs, dirs, feat, sw, cw)
            ctx.cfg = cfg
            ctx.count = count
            ctx.S_shape = tuple(S.shape) if S is not None else None
            ctx.tab_shapes = [tuple(t.shape) for t in tables]
        ctx.n_tables = len(tables)
        ctx.has_graph = save
        return sigmas, rgbs

    @staticmethod
    def backward(ctx, grad_sigmas, grad_rgbs):
        n_in = _field_forward.N_FIXED + ctx.n_tables
        if not ctx.has_graph:
            return (None,) * n_in
        if ctx.use_masks:
            xyzs, masks, sigmas, rgbs, sw, cw = ctx.saved_tensors
            cfg = ctx.cfg
            direct = cfg.S_sink is not None
            G = cfg.S_sink if direct else torch.zeros(ctx.S_shape, dtype=torch.float32, device=xyzs.device)
            entry = "nsig_field_backward_tc_masks" if BACKWARD_MODE == "tc_masks" else "nsig_field_backward_masks"
            _lib.call(entry, _P(xyzs), xyzs.shape[0], cfg.bound, _P(masks), _P(sigmas), _P(rgbs),
                      _P(grad_sigmas.contiguous().float()), _P(grad_rgbs.contiguous().float()), _P(sw), _P(cw),
                      cfg.density_scale, _P(ctx.count), cfg.msg_resolution, cfg.log2_T, _P(G))
            return (None, None, None if direct else G) + (None,) * (n_in - 3)
        xyzs, dirs, feat, sw, cw = ctx.saved_tensors
        cfg = ctx.cfg
        M = xyzs.shape[0]
        dev = xyzs.device
        grad_sigmas = grad_sigmas.contiguous().float()
        grad_rgbs = grad_rgbs.contiguous().float()
        need_S = ctx.S_shape is not None and ctx.needs_input_grad[2]
        need_w = ctx.needs_input_grad[7] or ctx.needs_input_grad[8]
        need_tab = ctx.needs_input_grad[_field_forward.N_FIXED:]
        direct = need_S and cfg.S_sink is not None
        G = cfg.S_sink if direct else (torch.zeros(ctx.S_shape, dtype=torch.float32, device=dev) if need_S else None)
        # rows past the live count are not written by the kernel: start from zeros in that case
        alloc = torch.zeros if ctx.count is not None else torch.empty
        grad_feat = alloc(M, 32, dtype=torch.float32, device=dev) if any(need_tab) else None
        gsw = torch.zeros(sw.numel(), dtype=torch.float32, device=dev) if need_w else None
        gcw = torch.zeros(cw.numel(), dtype=torch.float32, device=dev) if need_w else None
        if G is not None and grad_feat is None and gsw is None and BACKWARD_MODE == "tc":
            # watermark training (the hot path): dL/dS only - tcgen05/TMEM kernel (csrc/field_tc.cu)
            _lib.call("nsig_field_backward_tc", _P(xyzs), _P(dirs), M, cfg.bound, _P(feat), _P(grad_sigmas), _P(grad_rgbs),
                      _P(sw), _P(cw), cfg.density_scale, _P(ctx.count), cfg.msg_resolution, cfg.log2_T, _P(G))
        else:
            _lib.call("nsig_field_backward", _P(xyzs), _P(dirs), M, cfg.bound, _P(feat), _P(grad_sigmas), _P(grad_rgbs),
                      _P(sw), _P(cw), cfg.density_scale, _P(ctx.count), cfg.msg_resolution, cfg.log2_T, _P(G),
                      _P(grad_feat), _P(gsw), _P(gcw))
        tab_grads = [None] * ctx.n_tables
        if any(need_tab):
            xn = (xyzs + cfg.bound) * (1.0 / (2.0 * cfg.bound))  # network_wtmk_tcnn.py:101
            gt = [torch.zeros(s, dtype=torch.float32, device=dev) for s in ctx.tab_shapes]
            _lib.call("nsig_hash_encode_backward", _P(xn.contiguous()), _P(grad_feat.contiguous()), M,
                      _lib.pointer_array(gt), _lib.float_array(cfg.resolutions), len(gt), cfg.log2_T)
            tab_grads = [g if n else None for g, n in zip(gt, need_tab)]
        return (None, None, None if direct else G, None, None, None, None,
                gsw if ctx.needs_input_grad[7] else None, gcw if ctx.needs_input_grad[8] else None) + tuple(tab_grads)


def field_forward(xyzs, dirs, S, count, cfg, sigma_mlp, color_mlp, tables):
    return _field_forward.apply(xyzs, dirs, S, count, cfg, sigma_mlp, color_mlp, sigma_mlp.params, color_mlp.params,
                                *tables)


@torch.no_grad()
def field_density(xyzs, S, cfg, sigma_mlp, tables, want_geo=True):
    """sigma [M] fp32 (and geo_feat [M,15] fp16) — NeRFNetwork.density, no gradient (it is only used by
    update_extra_state and the non-cuda_ray renderer's sampling passes)."""
    xyzs = xyzs.contiguous().float()
    M = xyzs.shape[0]
    dev = xyzs.device
    sigmas = torch.empty(M, dtype=torch.float32, device=dev)
    geo = torch.empty(M, 15, dtype=torch.float16, device=dev) if want_geo else None
    tabs = [t.contiguous() for t in tables]
    Sc = S.contiguous() if S is not None else None
    _lib.call("nsig_field_density", _P(xyzs), M, cfg.bound, _lib.pointer_array(tabs), _lib.float_array(cfg.resolutions),
              cfg.log2_T, _P(Sc), cfg.msg_resolution, _P(sigma_mlp.half_weights()), cfg.density_scale, _P(sigmas),
              _P(geo), *_shadow_args(cfg))
    return sigmas, geo


@torch.no_grad()
def color_forward(dirs, geo_feat, color_mlp):
    """rgb [M,3] fp32 from view directions and geo features (NeRFNetwork.color)."""
    dirs = dirs.contiguous().float()
    geo = geo_feat.contiguous().to(torch.float16)
    M = dirs.shape[0]
    rgbs = torch.empty(M, 3, dtype=torch.float32, device=dirs.device)
    _lib.call("nsig_color_forward", _P(dirs), _P(geo), M, _P(color_mlp.half_weights()), _P(rgbs))
    return rgbs


@torch.no_grad()
def render_rays(rays_o, rays_d, aabb, min_near, bitfield, cascade, grid_size, dt_gamma, max_steps, T_thresh, noises,
                S, cfg, sigma_mlp, color_mlp, tables):
    """Whole-frame inference in one persistent kernel (nsig_render_rays): returns the accumulators of the
    reference's alive-ray loop — weights_sum [N], depth [N] (sum w*t), image [N,3] without background —
    plus nears/fars [N] and a device counter holding the number of samples evaluated."""
    rays_o = rays_o.contiguous().float()
    rays_d = rays_d.contiguous().float()
    N = rays_o.shape[0]
    dev = rays_o.device
    weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
    depth = torch.empty(N, dtype=torch.float32, device=dev)
    image = torch.empty(N, 3, dtype=torch.float32, device=dev)
    nears = torch.empty(N, dtype=torch.float32, device=dev)
    fars = torch.empty(N, dtype=torch.float32, device=dev)
    counters = torch.zeros(2, dtype=torch.int32, device=dev)  # [work counter, samples evaluated]
    tabs = [t.contiguous() for t in tables]
    Sc = S.contiguous() if S is not None else None
    _lib.call("nsig_render_rays", _P(rays_o), _P(rays_d), N, _P(aabb.contiguous()), float(min_near), cfg.bound,
              _P(bitfield.contiguous()), int(cascade), int(grid_size), float(dt_gamma), int(max_steps), float(T_thresh),
              _P(noises), _lib.pointer_array(tabs), _lib.float_array(cfg.resolutions), cfg.log2_T, _P(Sc),
              cfg.msg_resolution, _P(sigma_mlp.half_weights()), _P(color_mlp.half_weights()), cfg.density_scale,
              _P(counters[0:1]), _P(weights_sum), _P(depth), _P(image), _P(nears), _P(fars), _P(counters[1:2]),
              *_shadow_args(cfg))
    return weights_sum, depth, image, nears, fars, counters[1]


@torch.no_grad()
def grid_update(density_grid, bitfield, full, decay, density_thresh, C, H, bound, seed, S, cfg, sigma_mlp, tables,
                cells=None, noise=None):
    """Host side of NeRFRenderer.update_extra_state (csrc/grid.cu): density sweep + EMA + packbits, in place on
    `density_grid` [C,H^3] and `bitfield`.  Returns a device float[2] = (mean_density, threshold used)."""
    dev = density_grid.device
    H3 = H ** 3
    sums = torch.zeros(1, dtype=torch.float64, device=dev)
    stats = torch.empty(2, dtype=torch.float32, device=dev)
    tabs = [t.contiguous() for t in tables]
    Sc = S.contiguous() if S is not None else None
    field = (_lib.pointer_array(tabs), _lib.float_array(cfg.resolutions), cfg.log2_T, _P(Sc), cfg.msg_resolution,
             _P(sigma_mlp.half_weights()), cfg.density_scale)
    if noise is not None:
        noise = noise.contiguous().float()
    if full:
        _lib.call("nsig_grid_sweep", _P(density_grid), None, None, H3, _P(noise), seed, C, H, bound, decay, *field,
                  _P(sums), *_shadow_args(cfg))
    else:
        if cells is None:
            n = H3 // 4
            cells = torch.empty(C, 2 * n, dtype=torch.int32, device=dev)
            scratch = torch.empty(_lib.load().nsig_grid_sample_cells_scratch_bytes(C, H), dtype=torch.uint8, device=dev)
            _lib.call("nsig_grid_sample_cells", _P(density_grid), C, H, n, n, seed ^ 0x5DEECE66D, _P(cells), _P(scratch))
        cells = cells.contiguous().to(torch.int32)
        tmp = torch.full_like(density_grid, -1.0)
        _lib.call("nsig_grid_sweep", _P(density_grid), _P(tmp), _P(cells), cells.shape[1], _P(noise), seed, C, H, bound,
                  decay, *field, None, *_shadow_args(cfg))
        _lib.call("nsig_grid_finalize", _P(density_grid), _P(tmp), C * H3, decay, _P(sums))
    _lib.call("nsig_grid_pack", _P(density_grid), C * H3 // 8, _P(sums), C * H3, density_thresh, _P(bitfield), _P(stats))
    return stats
