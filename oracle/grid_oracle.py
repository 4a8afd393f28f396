"""CPU restatement (numpy, fp32 op by op) of the occupancy-grid side of the path and of get_rays.

TEST INFRASTRUCTURE ONLY (see oracle/cpu.py).  Restates
  * NeRFRenderer.update_extra_state      nerf/renderer_wtmk.py:445-538
  * NeRFRenderer.mark_untrained_grid     nerf/renderer_wtmk.py:380-442
  * get_rays                             nerf/utils_wtmk_disen.py:59-143
Pinned by tests/golden/grid_golden.npz = outputs of those reference functions themselves, executed on CPU from
their source text (tests/golden/make_golden_grid.py), see tests/test_oracle_cpu.py.
"""
import numpy as np

f32 = np.float32


def morton3D(coords):
    """raymarching.cu:56-72 (10-bit interleave)."""
    def expand(v):
        v = v.astype(np.uint64)
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v & 0xFFFFFFFF
    c = np.asarray(coords).astype(np.uint64)
    return (expand(c[:, 0]) | (expand(c[:, 1]) << 1) | (expand(c[:, 2]) << 2)).astype(np.int64)


def morton3D_invert(idx):
    """raymarching.cu:74-81."""
    def inv(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xc30c30c3
        x = (x | (x >> 4)) & 0x0f00f00f
        x = (x | (x >> 8)) & 0xff0000ff
        x = (x | (x >> 16)) & 0x0000ffff
        return x
    i = np.asarray(idx).astype(np.uint64)
    return np.stack([inv(i), inv(i >> 1), inv(i >> 2)], axis=-1).astype(np.int64)


def cascade_scalars(cas, H, bound):
    """bound_c = min(2**cas, bound); half = bound_c / H in python floats (renderer_wtmk.py:466-468)."""
    bound_c = min(2 ** cas, bound)
    half = bound_c / H
    return f32(bound_c - half), f32(half), f32(half * 2)


def cell_positions(coords, noise, cas, H, bound, cuda_div=True):
    """renderer_wtmk.py:470-479: xyzs = 2*coords.float()/(H-1) - 1; cas_xyzs = xyzs * (bound_c - half);
    cas_xyzs += (rand*2 - 1) * half.  torch divides a tensor by a python scalar as a multiply with fl(1/s) on CUDA
    (ATen BinaryDivTrueKernel.cu) but as a true division on CPU: cuda_div selects which; the reference runs on
    CUDA, the golden fixture was produced on CPU."""
    scale, half, _ = cascade_scalars(cas, H, bound)
    c = np.asarray(coords).astype(f32)
    if cuda_div:
        x = (f32(2.0) * c) * (f32(1.0) / f32(H - 1)) - f32(1.0)
    else:
        x = (f32(2.0) * c) / f32(H - 1) - f32(1.0)
    x = x * scale
    if noise is not None:
        x = x + (np.asarray(noise, dtype=f32) * f32(2.0) - f32(1.0)) * half
    return x.astype(f32)


def ema_update(density_grid, tmp_grid, decay=0.95):
    """renderer_wtmk.py:521-524: valid = (grid >= 0) & (tmp >= 0); grid[valid] = max(grid*decay, tmp);
    mean_density = mean(clamp(grid, min=0))."""
    g = np.array(density_grid, dtype=f32, copy=True)
    t = np.asarray(tmp_grid, dtype=f32)
    valid = (g >= 0) & (t >= 0)
    g[valid] = np.maximum(g[valid] * f32(decay), t[valid])
    mean = f32(np.clip(g, 0, None).astype(np.float64).mean())
    return g, mean


def packbits(grid, thresh):
    """raymarching.cu:268-289: bit i of byte n = grid[8n+i] > thresh."""
    bits = (np.asarray(grid, dtype=f32).reshape(-1, 8) > f32(thresh)).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)


def update_extra_state(density_grid, cells, sigmas, density_thresh, decay=0.95):
    """One update given the visited cells ([C,n] Morton indices) and the densities evaluated there ([C,n]).
    Duplicate cells keep the LARGEST sigma (the reference's index_put keeps an arbitrary one).
    Returns (grid, mean_density, thresh, bitfield)."""
    C = density_grid.shape[0]
    tmp = -np.ones_like(density_grid, dtype=f32)
    for cas in range(C):
        np.maximum.at(tmp[cas], np.asarray(cells[cas]).astype(np.int64), np.asarray(sigmas[cas], dtype=f32))
    g, mean = ema_update(density_grid, tmp, decay)
    thresh = f32(min(float(mean), density_thresh))
    return g, mean, thresh, packbits(g, thresh)


def mark_untrained_grid(poses, intrinsic, C, H, bound, cuda_div=True):
    """renderer_wtmk.py:380-442 -> bool [C, H^3] (True = seen by at least one camera), Morton-indexed."""
    fx, fy, cx, cy = intrinsic
    poses = np.asarray(poses, dtype=f32)
    idx = np.arange(H ** 3)
    coords = morton3D_invert(idx)
    seen = np.zeros((C, H ** 3), dtype=bool)
    for cas in range(C):
        _, _, margin = cascade_scalars(cas, H, bound)
        w = cell_positions(coords, None, cas, H, bound, cuda_div)             # [n,3]
        for p in poses:
            cam = (w - p[:3, 3][None, :]).astype(f32) @ p[:3, :3]             # (x - t) @ R
            cam = cam.astype(f32)
            z = cam[:, 2]
            ok = (z > 0) & (np.abs(cam[:, 0]) < f32(cx / fx) * z + margin) & (np.abs(cam[:, 1]) < f32(cy / fy) * z + margin)
            seen[cas] |= ok
    return seen


def get_rays(poses, intrinsics, H, W, inds=None, cuda_div=True):
    """utils_wtmk_disen.py:59-143 for a given pixel list -> rays_o, rays_d [B,N,3] fp32 (cuda_div: see
    cell_positions)."""
    fx, fy, cx, cy = intrinsics
    poses = np.asarray(poses, dtype=f32)
    B = poses.shape[0]
    if inds is None:
        inds = np.broadcast_to(np.arange(H * W), (B, H * W))
    inds = np.asarray(inds)
    if inds.ndim == 1:
        inds = np.broadcast_to(inds, (B, inds.shape[0]))
    i = (inds % W).astype(f32) + f32(0.5)
    j = (inds // W).astype(f32) + f32(0.5)
    if cuda_div:
        xs = (i - f32(cx)) * (f32(1.0) / f32(fx))
        ys = (j - f32(cy)) * (f32(1.0) / f32(fy))
    else:
        xs = (i - f32(cx)) / f32(fx)
        ys = (j - f32(cy)) / f32(fy)
    zs = np.ones_like(xs)
    d = np.stack([xs, ys, zs], axis=-1).astype(f32)
    d = d / np.sqrt((d * d).sum(axis=-1, keepdims=True, dtype=f32)).astype(f32)
    rays_d = np.einsum('bnj,bij->bni', d, poses[:, :3, :3]).astype(f32)
    rays_o = np.broadcast_to(poses[:, None, :3, 3], rays_d.shape).astype(f32)
    return rays_o, rays_d
