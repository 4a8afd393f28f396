"""Port of the reference's pure-PyTorch CPU path — TEST / BASELINE INFRASTRUCTURE ONLY.

This is what `bench.py --impl reference` and the `cpu_baseline` leg time on the GPU box's host
cores (kind = "port": /root/reference is Python and cannot travel to the GPU box, and even its
"CPU path" imports CUDA-only modules — SURVEY.md 8d).  It restates, op for op in torch fp32:

  * hash_encoding.py:11-46,78-111       HashEmbedder (16 hashed levels, trilinear)
  * hash_encoding_wtmk_bit.py:99-116    message-bit HashEmbedder (one gather+trilerp per bit, summed)
  * hash_encoding.py:114-195            SHEncoder degree 4 (stands in for tcnn SphericalHarmonics)
  * nerf/"network copy.py":33-68        bias-free Linear + ReLU MLPs (stand in for tcnn FullyFusedMLP)
  * nerf/network_wtmk_tcnn.py:97-176    forward / density / color wiring
  * nerf/renderer_wtmk.py:12-46,125-253 NeRFRenderer.run (non-cuda_ray: uniform samples + optional hierarchical resampling)
  * raymarching.cu:108-144              near_far_from_aabb restated in torch (the reference calls CUDA here)
  * nerf/utils_wtmk_disen.py:579-646    train_step losses (BCE x temp 10 on decoded bits + MSE)

Pinned by tests/test_oracle_cpu.py against outputs of the reference's own code: tests/golden/hash_golden.npz (the encoder
modules), tests/golden/run_golden.npz (NeRFRenderer.run + sample_pdf, incl. upsample_steps > 0) and
tests/golden/trainstep_golden.npz (Trainer.train_step: losses, block pixels, gradients).
"""
import math

import torch
import torch.nn.functional as F

_OFFSETS = torch.tensor([[[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)]])  # hash_encoding.py:8
_OFFSETS_ON = {}


def _offsets(device):
    """BOX_OFFSETS on `device` (the reference keeps its copy on 'cuda'; this port runs on the CPU and, for the
    reference-composed CUDA step, on the GPU)."""
    key = str(device)
    if key not in _OFFSETS_ON:
        _OFFSETS_ON[key] = _OFFSETS.to(device)
    return _OFFSETS_ON[key]
_PRIMES = (1, 2654435761, 805459861)


def _hash(coords, log2_T):
    h = torch.zeros_like(coords)[..., 0]
    for a in range(3):
        h ^= coords[..., a] * _PRIMES[a]
    return h & ((1 << log2_T) - 1)


def _voxel(x, resolution, log2_T):
    """hash_encoding.py:24-46 with bounding_box = (0, 1)."""
    xc = torch.clamp(x, min=0, max=1)
    grid_size = 1 / resolution
    idx = torch.floor(xc / grid_size).int()
    vmin = idx * grid_size
    vmax = vmin + grid_size
    slots = _hash(idx.unsqueeze(1) + _offsets(idx.device), log2_T)
    return vmin, vmax, slots


def _trilerp(x, vmin, vmax, e):
    """hash_encoding.py:78-104 (e: [B,8,F])."""
    w = (x - vmin) / (vmax - vmin)
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    c00 = e[:, 0] * (1 - wx) + e[:, 4] * wx
    c01 = e[:, 1] * (1 - wx) + e[:, 5] * wx
    c10 = e[:, 2] * (1 - wx) + e[:, 6] * wx
    c11 = e[:, 3] * (1 - wx) + e[:, 7] * wx
    c0 = c00 * (1 - wy) + c10 * wy
    c1 = c01 * (1 - wy) + c11 * wy
    return c0 * (1 - wz) + c1 * wz


def level_resolutions(base, finest, n_levels):
    base_t, finest_t = torch.tensor(base), torch.tensor(finest)
    b = torch.exp((torch.log(finest_t) - torch.log(base_t)) / (n_levels - 1))
    return [torch.floor(base_t * b ** i) for i in range(n_levels)]


def hash_embed(x, tables, resolutions, log2_T=19):
    """HashEmbedder.forward: tables = list of [T,2] tensors."""
    outs = []
    for tab, res in zip(tables, resolutions):
        vmin, vmax, slots = _voxel(x, res, log2_T)
        outs.append(_trilerp(x, vmin, vmax, F.embedding(slots, tab)))
    return torch.cat(outs, dim=-1)


def msg_embed(x, tables, message, resolution=2048.0, log2_T=19):
    """message-bit HashEmbedder.forward: table 2i+bit_i per bit, summed over bits."""
    res = torch.tensor(float(resolution))
    outs = []
    for i in range(message.shape[0]):
        vmin, vmax, slots = _voxel(x, res, log2_T)
        tab = tables[2 * i + int(message[i].item())]
        outs.append(_trilerp(x, vmin, vmax, F.embedding(slots, tab)))
    return torch.sum(torch.stack(outs, dim=-1), dim=-1)


_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def sh_degree4(d):
    x, y, z = d.unbind(-1)
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    r = torch.empty((*d.shape[:-1], 16), dtype=d.dtype)
    r[..., 0] = 0.28209479177387814
    r[..., 1] = -_C1 * y
    r[..., 2] = _C1 * z
    r[..., 3] = -_C1 * x
    r[..., 4] = _C2[0] * xy
    r[..., 5] = _C2[1] * yz
    r[..., 6] = _C2[2] * (2.0 * zz - xx - yy)
    r[..., 7] = _C2[3] * xz
    r[..., 8] = _C2[4] * (xx - yy)
    r[..., 9] = _C3[0] * y * (3 * xx - yy)
    r[..., 10] = _C3[1] * xy * z
    r[..., 11] = _C3[2] * y * (4 * zz - xx - yy)
    r[..., 12] = _C3[3] * z * (2 * zz - 3 * xx - 3 * yy)
    r[..., 13] = _C3[4] * x * (4 * zz - xx - yy)
    r[..., 14] = _C3[5] * z * (xx - yy)
    r[..., 15] = _C3[6] * x * (xx - 3 * yy)
    return r


class PortField:
    """Random-init watermark (or clean, message_dim=0) field with the reference's shapes, on CPU."""

    def __init__(self, bound=1.0, message_dim=32, log2_T=19, seed=0, train_msg=True):
        g = torch.Generator().manual_seed(seed)
        self.bound, self.log2_T, self.message_dim = float(bound), log2_T, message_dim
        T = 1 << log2_T
        u = lambda: (torch.rand(T, 2, generator=g) * 2 - 1) * 1e-4  # hash_encoding.py:65-66
        self.base_tables = [u() for _ in range(16)]
        self.base_res = level_resolutions(16, 2048, 16)
        self.msg_tables = [u().requires_grad_(train_msg) for _ in range(2 * message_dim)]

        def xavier(fo, fi):
            s = math.sqrt(6.0 / (fi + fo))
            return (torch.rand(fo, fi, generator=g) * 2 - 1) * s
        self.Ws = [xavier(64, 32), xavier(16, 64)]
        self.Wc = [xavier(64, 32), xavier(64, 64), xavier(16, 64)]

    def encode(self, x, message):
        x = (x + self.bound) / (2 * self.bound)
        feat = hash_embed(x, self.base_tables, self.base_res, self.log2_T)
        if message is not None and self.message_dim > 0:
            m = msg_embed(x, self.msg_tables, message, 2048.0, self.log2_T)
            feat = torch.cat([feat[:, :-2], feat[:, -2:] + m], dim=-1)
        return feat

    def density(self, x, message=None):
        h = torch.relu(self.encode(x, message) @ self.Ws[0].t()) @ self.Ws[1].t()
        return {'sigma': torch.exp(h[..., 0]), 'geo_feat': h[..., 1:]}

    def color(self, d, geo_feat):
        sh = sh_degree4(((d + 1) / 2) * 2 - 1)
        h = torch.cat([sh, geo_feat, torch.zeros_like(geo_feat[..., :1])], dim=-1)
        h = torch.relu(h @ self.Wc[0].t())
        h = torch.relu(h @ self.Wc[1].t())
        return torch.sigmoid((h @ self.Wc[2].t())[..., :3])


def near_far_from_aabb(rays_o, rays_d, bound, min_near=0.2):
    """raymarching.cu:108-144 in torch."""
    rd = 1 / rays_d
    t0 = (-bound - rays_o) * rd
    t1 = (bound - rays_o) * rd
    tmin, tmax = torch.minimum(t0, t1), torch.maximum(t0, t1)
    near = tmin.max(dim=-1).values.clamp(min=min_near)
    far = tmax.min(dim=-1).values
    miss = near > far
    big = torch.finfo(torch.float32).max
    return torch.where(miss, torch.full_like(near, big), near), torch.where(miss, torch.full_like(far, big), far)


def _alpha_weights(z, sigma, sample_dist):
    """Interval lengths (the last one is the nominal spacing), alphas and compositing weights of sorted depths z [N, T]
    (renderer_wtmk.py:204-208; the 1e-15 keeps the running product off exact zero)."""
    deltas = torch.cat([z[..., 1:] - z[..., :-1], sample_dist * torch.ones_like(z[..., :1])], dim=-1)
    alphas = 1 - torch.exp(-deltas * sigma)
    shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
    return deltas, alphas * torch.cumprod(shifted, dim=-1)[..., :-1]


def resample_depths(edges, weights, n_samples):
    """Deterministic inverse-CDF resampling (renderer_wtmk.py:12-46 with det=True, what run() uses in eval mode): the
    n_samples bin centres of [0, 1] pulled back through the piecewise-linear CDF that `weights` [N, B-1] define over the
    `edges` [N, B]."""
    pdf = weights + 1e-5
    pdf = pdf / pdf.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, n_samples).expand(cdf.shape[0], n_samples).contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = (hi - 1).clamp(min=0)
    hi = hi.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = cdf.gather(1, lo), cdf.gather(1, hi)
    e0, e1 = edges.gather(1, lo), edges.gather(1, hi)
    width = c1 - c0
    width = torch.where(width < 1e-5, torch.ones_like(width), width)
    return e0 + (u - c0) / width * (e1 - e0)


def render_run(field, rays_o, rays_d, message, num_steps=512, bg_color=1.0, min_near=0.2, upsample_steps=0,
               return_depth=False):
    """NeRFRenderer.run in eval mode, perturb=False (renderer_wtmk.py:125-253): uniform depths between near and far,
    optional hierarchical resampling (upsample_steps > 0: the extra points are evaluated WITHOUT the message, as the
    reference does at renderer_wtmk.py:185), colour only where the weight exceeds 1e-4, white-background blend.
    Returns (image, weights_sum[, depth])."""
    N = rays_o.shape[0]
    nears, fars = near_far_from_aabb(rays_o, rays_d, field.bound, min_near)
    nears, fars = nears.unsqueeze(-1), fars.unsqueeze(-1)
    z = torch.linspace(0.0, 1.0, num_steps).unsqueeze(0).expand(N, num_steps)
    z = nears + (fars - nears) * z
    sample_dist = (fars - nears) / num_steps

    def points(depths):
        return (rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * depths.unsqueeze(-1)).clamp(-field.bound, field.bound)

    xyzs = points(z)
    dens = field.density(xyzs.reshape(-1, 3), message)
    sigma, geo = dens['sigma'].view(N, num_steps), dens['geo_feat'].view(N, num_steps, -1)
    if upsample_steps > 0:
        with torch.no_grad():
            deltas, w = _alpha_weights(z, sigma, sample_dist)
            mids = z[..., :-1] + 0.5 * deltas[..., :-1]
            z_new = resample_depths(mids, w[:, 1:-1], upsample_steps)
        extra = field.density(points(z_new).reshape(-1, 3), None)
        z, order = torch.sort(torch.cat([z, z_new], dim=1), dim=1)
        sigma = torch.cat([sigma, extra['sigma'].view(N, upsample_steps)], dim=1).gather(1, order)
        geo = torch.cat([geo, extra['geo_feat'].view(N, upsample_steps, -1)], dim=1)
        geo = geo.gather(1, order.unsqueeze(-1).expand_as(geo))
        xyzs = points(z)
    T = z.shape[1]
    _, weights = _alpha_weights(z, sigma, sample_dist)
    mask = (weights > 1e-4).reshape(-1)
    dirs = rays_d.view(-1, 1, 3).expand_as(xyzs).reshape(-1, 3)
    rgbs = torch.zeros(N * T, 3)
    if mask.any():
        rgbs[mask] = field.color(dirs[mask], geo.reshape(N * T, -1)[mask])
    rgbs = rgbs.view(N, T, 3)
    weights_sum = weights.sum(dim=-1)
    image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2) + (1 - weights_sum).unsqueeze(-1) * bg_color
    if return_depth:
        depth = torch.sum(weights * ((z - nears) / (fars - nears)).clamp(0, 1), dim=-1)
        return image, weights_sum, depth
    return image, weights_sum


def train_step(field, decoder, batch, message, lambda_w=0.005, lambda_i=1.0, num_steps=512, return_terms=False):
    """One watermark training step (utils_wtmk_disen.py:579-646 + 1175): two render passes, decoder,
    losses, backward.  batch: rays_o_block/rays_d_block [md,pH,pW,3], rays_o/rays_d [n,3], gt [n,3]."""
    blk_o, blk_d = batch['rays_o_block'], batch['rays_d_block']
    shp = blk_o.shape[:-1]
    img_w, _ = render_run(field, blk_o.reshape(-1, 3), blk_d.reshape(-1, 3), message, num_steps)
    pred = img_w.view(*shp, 3).clamp(0, 1)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    decoded = decoder((pred.permute(0, 3, 1, 2) - mean) / std)
    img_c, _ = render_run(field, batch['rays_o'], batch['rays_d'], message, num_steps)
    lossi = F.mse_loss(img_c, batch['gt'], reduction='none').mean()
    lossw = F.binary_cross_entropy_with_logits(decoded * 10.0, message.unsqueeze(-1), reduction='mean')
    loss = lambda_w * lossw + lambda_i * lossi
    loss.backward()
    if return_terms:
        return float(loss.detach()), float(lossi.detach()), float(lossw.detach()), pred.detach()
    return float(loss.detach())
