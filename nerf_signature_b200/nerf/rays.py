"""Ray generation with the reference's interface: get_rays (nerf/utils_wtmk_disen.py:59-143).

The reference materialises two full [B, H*W] pixel meshgrids every step just to gather N of them, then runs ~15
elementwise/matmul kernels.  Here the pixel choice stays in torch (same calls, same generator stream: randint /
multinomial) and the geometry is one kernel (nsig_get_rays): pixel id -> camera direction -> normalise -> rotate.
"""
import torch

from .. import _lib

_P = _lib.ptr


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1):
    """poses [B,4,4] cam2world (CUDA), intrinsics (fx, fy, cx, cy), -> dict(rays_o [B,N,3], rays_d [B,N,3],
    inds [B,N] (+ inds_coarse when error_map is given)); N <= 0 renders every pixel."""
    device = poses.device
    B = poses.shape[0]
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    results = {}

    if N > 0:
        N = min(N, H * W)
        if patch_size > 1:
            # random top-left corners, then the patch offsets (utils_wtmk_disen.py:86-103)
            num_patch = N // (patch_size ** 2)
            inds_x = torch.randint(0, H - patch_size, size=[num_patch], device=device)
            inds_y = torch.randint(0, W - patch_size, size=[num_patch], device=device)
            off = torch.arange(patch_size, device=device)
            px = (inds_x[:, None, None] + off[None, :, None]).expand(num_patch, patch_size, patch_size)
            py = (inds_y[:, None, None] + off[None, None, :]).expand(num_patch, patch_size, patch_size)
            inds = (px * W + py).reshape(-1)
            N = inds.shape[0]
            inds = inds.expand([B, N])
        elif error_map is None:
            inds = torch.randint(0, H * W, size=[N], device=device)  # may duplicate
            inds = inds.expand([B, N])
        else:
            # weighted sampling on the 128x128 error grid, then a random pixel inside the coarse cell (:110-121)
            inds_coarse = torch.multinomial(error_map.to(device), N, replacement=False)
            inds_x, inds_y = inds_coarse // 128, inds_coarse % 128
            sx, sy = H / 128, W / 128
            inds_x = (inds_x * sx + torch.rand(B, N, device=device) * sx).long().clamp(max=H - 1)
            inds_y = (inds_y * sy + torch.rand(B, N, device=device) * sy).long().clamp(max=W - 1)
            inds = inds_x * W + inds_y
            results['inds_coarse'] = inds_coarse
        results['inds'] = inds
    else:
        N = H * W
        inds = None
        results['inds'] = torch.arange(H * W, device=device).expand([B, H * W])

    poses_c = poses.detach().to(torch.float32).contiguous()
    rays_o = torch.empty(B, N, 3, dtype=torch.float32, device=device)
    rays_d = torch.empty(B, N, 3, dtype=torch.float32, device=device)
    if inds is None:
        ind_ptr, stride = None, 0
    else:
        shared = inds.stride(0) == 0 or B == 1       # the reference's expand([B, N]): one list for the whole batch
        flat = (inds[0] if shared else inds).to(torch.int64).contiguous()
        ind_ptr, stride = _P(flat), (0 if shared else N)
    _lib.call("nsig_get_rays", _P(poses_c), B, fx, fy, cx, cy, int(H), int(W), ind_ptr, stride, N, _P(rays_o), _P(rays_d))
    results['rays_o'] = rays_o
    results['rays_d'] = rays_d
    return results
