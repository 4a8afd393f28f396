"""ctypes/numpy front-end of the C oracle (oracle/raymarch_oracle.c, oracle/hash_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports it.
All functions take and return numpy arrays (fp32 / int32 / uint8, C-contiguous).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnsig_oracle.so")
_SRCS = [os.path.join(_HERE, f) for f in ("raymarch_oracle.c", "hash_oracle.c")]


def build(force=False):
    """Compile the C oracle with gcc (seconds)."""
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in _SRCS if os.path.exists(s))
    if stale:
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math",
               "-fopenmp", "-o", _SO] + _SRCS + ["-lm"]
        subprocess.check_call(cmd)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


_u32 = ctypes.c_uint32
_fl = ctypes.c_float


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3), _f(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().oracle_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), _u32(N), _fl(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().oracle_sph_from_ray(_p(rays_o), _p(rays_d), _fl(radius), _u32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i(coords)
    N = coords.shape[0]
    out = np.empty(N, np.int32)
    lib().oracle_morton3D(_p(coords), _u32(N), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i(indices)
    N = indices.shape[0]
    out = np.empty((N, 3), np.int32)
    lib().oracle_morton3D_invert(_p(indices), _u32(N), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().oracle_packbits(_p(grid), _u32(N), _fl(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises=None, M=None,
                     dt_gamma=0.0, max_steps=1024, counter=None):
    """Returns (xyzs[M,3], dirs[M,3], deltas[M,2], rays[N,3], counter[2]); buffers are zero-filled
    like the reference wrapper's (raymarching.py:205-207)."""
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    if M is None:
        M = N * max_steps
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    nears, fars = _f(nears), _f(fars)
    noises = np.zeros(N, np.float32) if noises is None else _f(noises)
    xyzs, dirs = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    rays = np.empty((N, 3), np.int32)
    counter = np.zeros(2, np.int32) if counter is None else _i(counter).copy()
    lib().oracle_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), _fl(bound), _fl(dt_gamma),
                                  _u32(max_steps), _u32(N), _u32(C), _u32(H), _u32(M), _p(nears), _p(fars),
                                  _p(xyzs), _p(dirs), _p(deltas), _p(rays), _p(counter), _p(noises))
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f(sigmas), _f(rgbs), _f(deltas), _i(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    lib().oracle_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), _u32(M), _u32(N),
                                              _fl(T_thresh), _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, T_thresh=1e-4):
    grad_ws, grad_image = _f(grad_ws), _f(grad_image)
    sigmas, rgbs, deltas, rays, ws, image = _f(sigmas), _f(rgbs), _f(deltas), _i(rays), _f(ws), _f(image)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gc = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().oracle_composite_rays_train_backward(_p(grad_ws), _p(grad_image), _p(sigmas), _p(rgbs), _p(deltas),
                                               _p(rays), _p(ws), _p(image), _u32(M), _u32(N), _fl(T_thresh),
                                               _p(gs), _p(gc))
    return gs, gc


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars,
               align=-1, noises=None, dt_gamma=0.0, max_steps=1024):
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    rays_alive, rays_t = _i(rays_alive), _f(rays_t)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    nears, fars = _f(nears), _f(fars)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    noises = np.zeros(max(n_alive, 1), np.float32) if noises is None else _f(noises)
    xyzs, dirs = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    lib().oracle_march_rays(_u32(n_alive), _u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d),
                            _fl(bound), _fl(dt_gamma), _u32(max_steps), _u32(C), _u32(H), _p(bitfield),
                            _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(noises))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image,
                   T_thresh=1e-2):
    """In-place on rays_alive, rays_t, weights_sum, depth, image (must be the right dtypes)."""
    for a, dt in ((rays_alive, np.int32), (rays_t, np.float32), (weights_sum, np.float32),
                  (depth, np.float32), (image, np.float32)):
        assert a.dtype == dt and a.flags["C_CONTIGUOUS"]
    sigmas, rgbs, deltas = _f(sigmas), _f(rgbs), _f(deltas)
    lib().oracle_composite_rays(_u32(n_alive), _u32(n_step), _fl(T_thresh), _p(rays_alive), _p(rays_t),
                                _p(sigmas), _p(rgbs), _p(deltas), _p(weights_sum), _p(depth), _p(image))


# ---------------------------------------------------------------------------------------
# hash encoders
# ---------------------------------------------------------------------------------------
def _ptr_array(tables):
    tabs = [_f(t) for t in tables]
    arr = (ctypes.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    return tabs, arr


def hash_encode_forward(x, tables, resolutions, log2_T=19, want_slots=False):
    x = _f(x).reshape(-1, 3)
    B, L = x.shape[0], len(tables)
    tabs, arr = _ptr_array(tables)
    res = _f(resolutions)
    out = np.empty((B, 2 * L), np.float32)
    slots = np.empty((B, L, 8), np.int32) if want_slots else None
    lib().oracle_hash_encode_forward(_p(x), _u32(B), arr, _p(res), _u32(L), _u32(log2_T), _p(out),
                                     _p(slots) if want_slots else None)
    return (out, slots) if want_slots else out


def hash_encode_backward(x, grad_out, resolutions, n_levels, log2_T=19):
    x, grad_out = _f(x).reshape(-1, 3), _f(grad_out)
    B = x.shape[0]
    grads = [np.zeros((1 << log2_T, 2), np.float32) for _ in range(n_levels)]
    arr = (ctypes.c_void_p * n_levels)(*[g.ctypes.data for g in grads])
    res = _f(resolutions)
    lib().oracle_hash_encode_backward(_p(x), _p(grad_out), _u32(B), arr, _p(res), _u32(n_levels), _u32(log2_T))
    return grads


def msg_encode_forward(x, tables, message, resolution=2048.0, log2_T=19):
    x = _f(x).reshape(-1, 3)
    B = x.shape[0]
    message = _f(message)
    tabs, arr = _ptr_array(tables)
    out = np.empty((B, 2), np.float32)
    lib().oracle_msg_encode_forward(_p(x), _u32(B), arr, _u32(message.shape[0]), _p(message), _fl(resolution),
                                    _u32(log2_T), _p(out))
    return out


def msg_encode_backward(x, grad_out, resolution=2048.0, log2_T=19):
    x, grad_out = _f(x).reshape(-1, 3), _f(grad_out)
    G = np.zeros((1 << log2_T, 2), np.float32)
    lib().oracle_msg_encode_backward(_p(x), _p(grad_out), _u32(x.shape[0]), _fl(resolution), _u32(log2_T), _p(G))
    return G
