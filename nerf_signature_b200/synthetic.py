"""Seeded synthetic inputs shared by the tests, smoke() and bench.py:
Blender-shaped cameras (800x800, fl = 0.5*W/tan(0.5*0.6911112), orbit radius 4.0311*scale),
360-shaped cameras inside bound 2, and occupancy grids (solid sphere / Bernoulli)."""
import math

import numpy as np


def orbit_pose(theta, phi, radius):
    """cam2world looking at the origin (OpenGL-to-NGP style: camera z points away from the scene)."""
    c = np.array([radius * math.sin(theta) * math.sin(phi), radius * math.cos(theta),
                  radius * math.sin(theta) * math.cos(phi)], dtype=np.float64)
    fwd = -c / np.linalg.norm(c)
    up = np.array([0.0, -1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    up2 = np.cross(right, fwd)
    pose = np.eye(4)
    pose[:3, 0] = right
    pose[:3, 1] = up2
    pose[:3, 2] = fwd
    pose[:3, 3] = c
    return pose.astype(np.float32)


def camera_rays(pose, H, W, focal, pixel_ids):
    """Rays through the centres of the given flat pixel ids (get_rays, utils_wtmk_disen.py:59-143)."""
    i = (pixel_ids % W).astype(np.float32) + 0.5
    j = (pixel_ids // W).astype(np.float32) + 0.5
    dirs = np.stack([(i - W / 2) / focal, (j - H / 2) / focal, np.ones_like(i)], -1)
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    rays_d = dirs @ pose[:3, :3].T
    rays_o = np.broadcast_to(pose[:3, 3], rays_d.shape)
    return np.ascontiguousarray(rays_o, np.float32), np.ascontiguousarray(rays_d, np.float32)


def blender_rays(n_rays, seed=0, scale=0.8, H=800, W=800):
    rs = np.random.RandomState(seed)
    focal = 0.5 * W / math.tan(0.5 * 0.6911112)
    pose = orbit_pose(rs.uniform(math.pi / 3, 2 * math.pi / 3), rs.uniform(0, 2 * math.pi), 4.0311 * scale)
    ids = rs.randint(0, H * W, size=n_rays)
    return camera_rays(pose, H, W, focal, ids)


def rays_360(n_rays, seed=0, H=756, W=1008):
    rs = np.random.RandomState(seed)
    focal = 0.5 * W / math.tan(0.5 * 0.9)
    pose = orbit_pose(rs.uniform(math.pi / 3, 2 * math.pi / 3), rs.uniform(0, 2 * math.pi), 4.0 * 0.33)
    ids = rs.randint(0, H * W, size=n_rays)
    return camera_rays(pose, H, W, focal, ids)


def morton3D_np(x, y, z):
    def expand(v):
        v = v.astype(np.uint64)
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v
    return (expand(x) | (expand(y) << 1) | (expand(z) << 2)).astype(np.int64)


def sphere_grid(C, H=128, radius_frac=0.8, bound=None):
    """density grid [C, H^3] in Morton order: 1 inside a sphere of radius radius_frac * cascade bound."""
    ar = np.arange(H)
    xx, yy, zz = np.meshgrid(ar, ar, ar, indexing="ij")
    idx = morton3D_np(xx.ravel(), yy.ravel(), zz.ravel())
    pos = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], -1).astype(np.float32)
    pos = (pos + 0.5) / H * 2 - 1
    inside = (np.linalg.norm(pos, axis=-1) < radius_frac).astype(np.float32)
    grid = np.zeros((C, H ** 3), np.float32)
    for c in range(C):
        grid[c, idx] = inside
    return grid


def bernoulli_grid(C, H=128, p=0.5, seed=0):
    rs = np.random.RandomState(seed)
    return (rs.uniform(size=(C, H ** 3)) < p).astype(np.float32)


def packbits_np(grid, thresh):
    bits = (grid.reshape(-1, 8) > thresh).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(-1).astype(np.uint8)
