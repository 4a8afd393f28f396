"""`_backend` - the call surface of the reference's pybind11 extension (raymarching/src/bindings.cpp, declarations in
raymarching/src/raymarching.h:7-18) on top of libnsig_b200.so.

This is the file a reference maintainer drops in as `raymarching/backend.py` (the reference's own backend.py JIT-compiles
raymarching.cu + bindings.cpp): the reference's UNMODIFIED `raymarching/raymarching.py` - its autograd Functions, buffer
allocation, zero fills, `.item()` reads and AMP decorators - then runs on the sm_100a kernels, because every one of the
ten functions it calls (`_backend.near_far_from_aabb(...)`, ...) exists here with the same name, the same argument order
and the same in-place output semantics.  Tensors are handed to the C ABI (include/nsig.h) as raw device pointers together
with torch's current stream; a non-zero status raises.  No CPU path: CPU tensors are refused by `_lib.ptr`.

The package's own `raymarching.py` does not go through this module (it allocates less and keeps the sample count on the
device); both bind the same entry points.
"""
import torch

from .. import _lib

_P = _lib.ptr


def _dense(*tensors):
    """The pybind11 functions receive at::Tensor and read `.data_ptr()` of whatever layout arrives; the reference wrapper
    always passes contiguous tensors.  Checked here (by _lib.ptr as well) so that a strided view fails loudly."""
    for t in tensors:
        if t is not None and not t.is_contiguous():
            raise _lib.NsigError("raymarching backend: tensors must be contiguous")


class _Backend:
    """Namespace object: `from .backend import _backend` (raymarching.py:8-11 of the reference)."""

    # ---- raymarching.h:7-11 ---------------------------------------------------------------------------------------
    @staticmethod
    def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
        _dense(rays_o, rays_d, aabb, nears, fars)
        _lib.call("nsig_near_far_from_aabb", _P(rays_o), _P(rays_d), _P(aabb), int(N), float(min_near), _P(nears), _P(fars))

    @staticmethod
    def sph_from_ray(rays_o, rays_d, radius, N, coords):
        _dense(rays_o, rays_d, coords)
        _lib.call("nsig_sph_from_ray", _P(rays_o), _P(rays_d), float(radius), int(N), _P(coords))

    @staticmethod
    def morton3D(coords, N, indices):
        _dense(coords, indices)
        _lib.call("nsig_morton3D", _P(coords), int(N), _P(indices))

    @staticmethod
    def morton3D_invert(indices, N, coords):
        _dense(indices, coords)
        _lib.call("nsig_morton3D_invert", _P(indices), int(N), _P(coords))

    @staticmethod
    def packbits(grid, N, density_thresh, bitfield):
        _dense(grid, bitfield)
        _lib.call("nsig_packbits", _P(grid), int(N), float(density_thresh), _P(bitfield))

    # ---- raymarching.h:13-15 --------------------------------------------------------------------------------------
    @staticmethod
    def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays,
                         counter, noises):
        """The reference wrapper zero-fills xyzs / dirs / deltas before the call (raymarching.py:205-207), so the rows
        past the counter need no clearing here.  The per-ray scan state lives in a scratch buffer of this call."""
        _dense(rays_o, rays_d, grid, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
        nbytes = _lib.load().nsig_march_rays_train_scratch_bytes(int(N))
        scratch = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=rays_o.device)
        _lib.call("nsig_march_rays_train", _P(rays_o), _P(rays_d), _P(grid), float(bound), float(dt_gamma), int(max_steps),
                  int(N), int(C), int(H), int(M), _P(nears), _P(fars), _P(xyzs), _P(dirs), _P(deltas), _P(rays), _P(counter),
                  _P(noises), _P(scratch))

    @staticmethod
    def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image):
        _dense(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        _lib.call("nsig_composite_rays_train_forward", _P(sigmas), _P(rgbs), _P(deltas), _P(rays), int(M), int(N),
                  float(T_thresh), _P(weights_sum), _P(depth), _P(image))

    @staticmethod
    def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N,
                                      T_thresh, grad_sigmas, grad_rgbs):
        _dense(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, grad_sigmas, grad_rgbs)
        _lib.call("nsig_composite_rays_train_backward", _P(grad_weights_sum), _P(grad_image), _P(sigmas), _P(rgbs), _P(deltas),
                  _P(rays), _P(weights_sum), _P(image), int(M), int(N), float(T_thresh), _P(grad_sigmas), _P(grad_rgbs))

    # ---- raymarching.h:17-18 --------------------------------------------------------------------------------------
    @staticmethod
    def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars,
                   xyzs, dirs, deltas, noises):
        _dense(rays_alive, rays_t, rays_o, rays_d, grid, nears, fars, xyzs, dirs, deltas, noises)
        _lib.call("nsig_march_rays", int(n_alive), int(n_step), _P(rays_alive), _P(rays_t), _P(rays_o), _P(rays_d), float(bound),
                  float(dt_gamma), int(max_steps), int(C), int(H), _P(grid), _P(nears), _P(fars), _P(xyzs), _P(dirs), _P(deltas),
                  _P(noises))

    @staticmethod
    def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
        _dense(rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image)
        _lib.call("nsig_composite_rays", int(n_alive), int(n_step), float(T_thresh), _P(rays_alive), _P(rays_t), _P(sigmas),
                  _P(rgbs), _P(deltas), _P(weights_sum), _P(depth), _P(image))


_backend = _Backend()

__all__ = ['_backend']
