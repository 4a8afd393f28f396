"""Drop-in replacement for the reference's `raymarching` module
(/root/reference/raymarching/raymarching.py): the same nine callables with the same positional
signatures, return shapes, dtypes, in-place side effects and AMP behaviour, backed by the
sm_100a kernels of libnsig_b200.so through the C ABI (include/nsig.h).

Differences, all inside the reference's own nondeterminism or pure overhead:
  * `march_rays_train`: ray n owns row n of `rays` and offsets follow ray order (the reference
    assigns both with atomicAdd, raymarching.cu:405-406); no 128 MB zero-fill — only the padding
    rows are cleared; `torch.cuda.empty_cache()` (raymarching.py:231) is not called.
  * kernels run on torch's current stream instead of the legacy default stream.
There is no CPU path: CPU inputs are moved to CUDA exactly as the reference does.
"""
import torch
from torch.autograd import Function

from .. import _lib

try:  # torch >= 2.4
    from torch.amp import custom_bwd as _custom_bwd, custom_fwd as _custom_fwd

    def custom_fwd(**kw):
        return _custom_fwd(device_type="cuda", **kw)

    def custom_bwd(fn):
        return _custom_bwd(fn, device_type="cuda")
except ImportError:  # pragma: no cover
    from torch.cuda.amp import custom_bwd, custom_fwd

_P = _lib.ptr


def _scratch(N, device):
    nbytes = _lib.load().nsig_march_rays_train_scratch_bytes(N)
    return torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=device)


# ----------------------------------------
# utils
# ----------------------------------------

class _near_far_from_aabb(Function):
    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        ''' near_far_from_aabb (reference raymarching.py:19-49)
        rays_o, rays_d: float [N, 3]; aabb: float [6]; returns nears, fars: float [N]
        '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        aabb = aabb.to(device=rays_o.device, dtype=torch.float32).contiguous()
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        _lib.call("nsig_near_far_from_aabb", _P(rays_o), _P(rays_d), _P(aabb), N, float(min_near), _P(nears), _P(fars))
        return nears, fars

near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, radius):
        ''' sph_from_ray (reference raymarching.py:52-80): coords [N, 2] in [-1, 1] '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        _lib.call("nsig_sph_from_ray", _P(rays_o), _P(rays_d), float(radius), N, _P(coords))
        return coords

sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        ''' morton3D (reference raymarching.py:83-103): coords int32 [N,3] in [0,128) -> int32 [N] '''
        if not coords.is_cuda: coords = coords.cuda()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        _lib.call("nsig_morton3D", _P(coords.int().contiguous()), N, _P(indices))
        return indices

morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        ''' morton3D_invert (reference raymarching.py:105-126): int32 [N] -> int32 [N,3] '''
        if not indices.is_cuda: indices = indices.cuda()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        _lib.call("nsig_morton3D_invert", _P(indices.int().contiguous()), N, _P(coords))
        return coords

morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)
    def forward(ctx, grid, thresh, bitfield=None):
        ''' packbits (reference raymarching.py:129-155): grid float [C, H^3] -> uint8 [C*H^3/8] '''
        if not grid.is_cuda: grid = grid.cuda()
        grid = grid.contiguous()
        C = grid.shape[0]
        H3 = grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        _lib.call("nsig_packbits", _P(grid), N, float(thresh), _P(bitfield))
        return bitfield

packbits = _packbits.apply

# ----------------------------------------
# train functions
# ----------------------------------------

class _march_rays_train(Function):
    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        ''' march rays to generate points (reference raymarching.py:161-235)
        Returns xyzs [M,3], dirs [M,3], deltas [M,2] (float) and rays [N,3] int32 = (id, offset, count).
        '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        if not density_bitfield.is_cuda: density_bitfield = density_bitfield.cuda()

        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        density_bitfield = density_bitfield.contiguous()

        N = rays_o.shape[0]
        M = N * max_steps

        use_mean = (not force_all_rays) and mean_count > 0
        if use_mean:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count

        dev = rays_o.device
        xyzs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
        deltas = torch.empty(M, 2, dtype=rays_o.dtype, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)

        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)

        noises = torch.rand(N, dtype=rays_o.dtype, device=dev) if perturb else None

        _lib.call("nsig_march_rays_train", _P(rays_o), _P(rays_d), _P(density_bitfield), float(bound), float(dt_gamma),
                  int(max_steps), N, int(C), int(H), M, _P(nears.contiguous()), _P(fars.contiguous()), _P(xyzs), _P(dirs),
                  _P(deltas), _P(rays), _P(step_counter), _P(noises), _P(_scratch(N, dev)))
        # rows the reference obtains from torch.zeros: the alignment padding, or everything up to M
        _lib.call("nsig_zero_sample_padding", _P(xyzs), _P(dirs), _P(deltas), _P(step_counter),
                  0 if use_mean else (int(align) if align > 0 else 1), M)

        if not use_mean:
            m = step_counter[0].item()  # D2H copy (the API's output shape depends on it)
            if align > 0:
                m += align - m % align
            xyzs = xyzs[:m]
            dirs = dirs[:m]
            deltas = deltas[:m]

        return xyzs, dirs, deltas, rays

march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        ''' composite rays' rgbs (reference raymarching.py:238-271)
        Returns weights_sum [N], depth [N], image [N,3]; differentiable in sigmas and rgbs.
        '''
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        deltas = deltas.contiguous()
        rays = rays.contiguous()

        M = sigmas.shape[0]
        N = rays.shape[0]

        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)

        _lib.call("nsig_composite_rays_train_forward", _P(sigmas), _P(rgbs), _P(deltas), _P(rays), M, N, float(T_thresh),
                  _P(weights_sum), _P(depth), _P(image))

        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh]
        ctx.zero_fill = True

        return weights_sum, depth, image

    @staticmethod
    @custom_bwd
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is not propagated, as in the reference (raymarching.py:275)
        grad_weights_sum = grad_weights_sum.contiguous()
        grad_image = grad_image.contiguous()

        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims

        # rows owned by no ray (alignment padding, dropped rays) must read as zero (raymarching.py:283-284);
        # the kernel itself writes every row of every kept ray, zeros after early termination
        alloc = torch.zeros_like if ctx.zero_fill else torch.empty_like
        grad_sigmas = alloc(sigmas)
        grad_rgbs = alloc(rgbs)

        _lib.call("nsig_composite_rays_train_backward", _P(grad_weights_sum), _P(grad_image), _P(sigmas), _P(rgbs),
                  _P(deltas), _P(rays), _P(weights_sum), _P(image), M, N, float(T_thresh), _P(grad_sigmas), _P(grad_rgbs))

        return grad_sigmas, grad_rgbs, None, None, None

composite_rays_train = _composite_rays_train.apply


class _composite_rays_train_live(_composite_rays_train):
    """composite_rays_train for callers that only ever consume the sample rows owned by rays (the
    renderer's sync-free path sizes its buffers for the worst case, N*max_steps rows, and every consumer
    stops at the live count): identical kernels, but the gradient buffers are not zero-filled."""

    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        out = _composite_rays_train.forward.__wrapped__(ctx, sigmas, rgbs, deltas, rays, T_thresh)
        ctx.zero_fill = False
        return out

composite_rays_train_live = _composite_rays_train_live.apply


class _composite_rays_train_blend(Function):
    """composite_rays_train followed by the epilogue of NeRFRenderer.run_cuda's training branch
    (renderer_wtmk.py:298-303) in the same kernels:
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color          (scalar bg_color)
        depth = clamp(depth - nears, min=0) / (fars - nears)
    Returns (weights_sum, depth, image) with those final values; differentiable in sigmas and rgbs through image and
    weights_sum (depth carries no gradient, as in the reference's composite backward, raymarching.py:275).
    Replaces 7 element-wise launches forward and ~6 backward per render call."""

    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, nears, fars, bg_color, T_thresh=1e-4, zero_fill=True):
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        deltas = deltas.contiguous()
        rays = rays.contiguous()
        M = sigmas.shape[0]
        N = rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth_raw = torch.empty(N, dtype=torch.float32, device=dev)     # raw composite outputs (image_raw is saved)
        image_raw = torch.empty(N, 3, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        _lib.call("nsig_composite_rays_train_blend_forward", _P(sigmas), _P(rgbs), _P(deltas), _P(rays), M, N,
                  float(T_thresh), float(bg_color), _P(nears.contiguous()), _P(fars.contiguous()), _P(weights_sum),
                  _P(depth_raw), _P(image_raw), _P(image), _P(depth))
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, image_raw)
        ctx.dims = [M, N, float(T_thresh), float(bg_color)]
        ctx.zero_fill = bool(zero_fill)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(depth)
        return weights_sum, depth, image

    @staticmethod
    @custom_bwd
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        sigmas, rgbs, deltas, rays, weights_sum, image_raw = ctx.saved_tensors
        M, N, T_thresh, bg_color = ctx.dims
        if grad_image is None:
            grad_image = torch.zeros_like(image_raw)
        grad_image = grad_image.contiguous().view(N, 3)
        gws = grad_weights_sum.contiguous() if grad_weights_sum is not None else None
        alloc = torch.zeros_like if ctx.zero_fill else torch.empty_like
        grad_sigmas = alloc(sigmas)
        grad_rgbs = alloc(rgbs)
        _lib.call("nsig_composite_rays_train_blend_backward", _P(gws), _P(grad_image), _P(sigmas), _P(rgbs), _P(deltas),
                  _P(rays), _P(weights_sum), _P(image_raw), M, N, T_thresh, bg_color, _P(grad_sigmas), _P(grad_rgbs))
        return grad_sigmas, grad_rgbs, None, None, None, None, None, None, None

composite_rays_train_blend = _composite_rays_train_blend.apply

# ----------------------------------------
# infer functions
# ----------------------------------------

class _march_rays(Function):
    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        ''' march rays for inference (reference raymarching.py:297-348)
        Returns xyzs [n_alive*n_step (padded to align), 3], dirs [.., 3], deltas [.., 2], zero rows
        where a ray ran out of samples.
        '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()

        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)

        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)

        dev = rays_o.device
        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)

        noises = torch.rand(n_alive, dtype=rays_o.dtype, device=dev) if perturb else None

        _lib.call("nsig_march_rays", int(n_alive), int(n_step), _P(rays_alive), _P(rays_t), _P(rays_o), _P(rays_d),
                  float(bound), float(dt_gamma), int(max_steps), int(C), int(H), _P(density_bitfield.contiguous()),
                  _P(near.contiguous()), _P(far.contiguous()), _P(xyzs), _P(dirs), _P(deltas), _P(noises))

        return xyzs, dirs, deltas

march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)  # sigmas & rgbs may arrive as fp16 under autocast
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        ''' composite rays for inference, in place on rays_alive / rays_t / weights_sum / depth / image
        (reference raymarching.py:351-373) '''
        _lib.call("nsig_composite_rays", int(n_alive), int(n_step), float(T_thresh), _P(rays_alive), _P(rays_t),
                  _P(sigmas.contiguous()), _P(rgbs.contiguous()), _P(deltas.contiguous()), _P(weights_sum), _P(depth),
                  _P(image))
        return tuple()

composite_rays = _composite_rays.apply
