"""Generate golden vectors for the raymarching kernels FROM THE UNMODIFIED REFERENCE CUDA EXTENSION.

Run on a GPU box (the reference kernels cannot execute in the GPU-less build container):
    gpurun -- python tests/golden/make_golden_raymarch.py gpurun_out/raymarch_golden.npz
and commit the result as tests/golden/raymarch_golden.npz.

The extension is oracle/_ref/_raymarching.so, compiled by oracle/build_ref.py from the sources where
they lie under /root/reference/raymarching/src (only -std=c++14 -> c++17).  Inputs are the
seeded fixtures of nerf_signature_b200/synthetic.py, so only the case parameters are stored; outputs
are stored in the canonical form of SURVEY F7 (rays sorted by id, each ray's samples in order), because
the reference assigns rows/offsets with atomicAdd.

Stored per case `c`:
  c_params  (n_rays, seed, bound, C, dt_gamma, perturb, kind)         [object -> json string]
  c_nears c_fars                       near_far_from_aabb
  c_counts [N] int32                   per-ray sample counts (rays[:,2] by ray id), c_total
  c_xyzs c_dirs c_deltas               canonical concatenation of every ray's samples
  c_ws c_depth c_image                 composite_rays_train_forward on the seeded sigmas/rgbs (T_thresh 1e-4)
  c_gsig c_grgb                        composite_rays_train_backward (seeded output gradients), canonical order
  c_inf_ws c_inf_depth c_inf_image c_inf_iters   the reference inference loop (march_rays/composite_rays,
                                       renderer_wtmk.py:336-367) driven with a closed-form field
plus `morton_*` / `packbits_*` fixtures.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from nerf_signature_b200 import synthetic as syn  # noqa: E402

CASES = [
    # name, rays, n_rays, seed, bound, C, grid kind, dt_gamma, perturb
    ("blender_sphere", "blender", 96, 3, 1.0, 1, "sphere", 0.0, False),
    ("blender_bernoulli_gamma", "blender", 64, 4, 1.0, 1, "bernoulli", 1.0 / 128, True),
    ("r360_sphere", "360", 64, 5, 2.0, 2, "sphere", 0.0, False),
    ("r360_bernoulli_gamma", "360", 64, 6, 2.0, 2, "bernoulli", 1.0 / 128, True),
    ("bound1p5_bernoulli", "360", 48, 7, 1.5, 2, "bernoulli", 0.0, True),
]


def case_inputs(case):
    """Seeded inputs of a case (shared with tests/test_oracle_cpu.py)."""
    name, cam, n, seed, bound, C, kind, dt_gamma, perturb = case
    rays_o, rays_d = (syn.blender_rays if cam == "blender" else syn.rays_360)(n, seed=seed)
    grid = syn.sphere_grid(C) if kind == "sphere" else syn.bernoulli_grid(C, p=0.3, seed=seed)
    bitfield = syn.packbits_np(grid, 0.5)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    noises = np.random.RandomState(seed + 1).uniform(size=n).astype(np.float32) if perturb else np.zeros(n, np.float32)
    return rays_o, rays_d, bitfield, aabb, noises


def seeded_field(M, seed):
    rs = np.random.RandomState(seed)
    sig = np.exp(rs.normal(0.0, 2.0, size=M)).astype(np.float32)
    sig[: M // 3] *= 50  # opaque early: exercises early termination
    return sig, rs.uniform(size=(M, 3)).astype(np.float32)


def closed_form_field(xyz):
    """Deterministic stand-in for the network in the inference loop (float32 numpy)."""
    s = (20.0 * np.abs(np.sin(7 * xyz[:, 0]) * np.cos(5 * xyz[:, 1]))).astype(np.float32)
    c = (0.5 + 0.5 * np.sin(xyz * 3)).astype(np.float32)
    return s, c


def canonical(rays, *bufs):
    rays = rays[np.argsort(rays[:, 0], kind="stable")]
    outs = [[] for _ in bufs]
    for rid, off, cnt in rays:
        for o, b in zip(outs, bufs):
            o.append(b[off:off + cnt])
    return rays[:, 2].astype(np.int32), [np.concatenate(o, 0) for o in outs]


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "_raymarching.so")
    spec = importlib.util.spec_from_file_location("_raymarching", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main(out_path):
    ref = load_ref()
    out = {}
    for case in CASES:
        name, cam, N, seed, bound, C, kind, dt_gamma, perturb = case
        rays_o, rays_d, bitfield, aabb, noises = case_inputs(case)
        t_o, t_d, t_b = cu(rays_o), cu(rays_d), cu(bitfield)
        nears = torch.empty(N, device="cuda"); fars = torch.empty(N, device="cuda")
        ref.near_far_from_aabb(t_o, t_d, cu(aabb), N, 0.2, nears, fars)
        M = N * 1024
        xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
        rays = torch.empty(N, 3, dtype=torch.int32, device="cuda")
        counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        ref.march_rays_train(t_o, t_d, t_b, bound, dt_gamma, 1024, N, C, 128, M, nears, fars, xyzs, dirs, deltas, rays,
                             counter, cu(noises))
        torch.cuda.synchronize()
        total = int(counter[0])
        assert int(counter[1]) == N
        r = rays.cpu().numpy()
        # composite on the reference's own (atomic-ordered) layout, stored canonically
        sig_c, rgb_c = seeded_field(total, seed + 11)       # defined in CANONICAL sample order
        counts, (cx, cd, cl) = canonical(r, xyzs.cpu().numpy(), dirs.cpu().numpy(), deltas.cpu().numpy())
        # scatter the canonical field values into the reference layout
        sig = np.zeros(M, np.float32); rgb = np.zeros((M, 3), np.float32)
        rs_sorted = r[np.argsort(r[:, 0], kind="stable")]
        pos = 0
        for rid, off, cnt in rs_sorted:
            sig[off:off + cnt] = sig_c[pos:pos + cnt]; rgb[off:off + cnt] = rgb_c[pos:pos + cnt]; pos += cnt
        ws = torch.empty(N, device="cuda"); dp = torch.empty(N, device="cuda"); im = torch.empty(N, 3, device="cuda")
        ref.composite_rays_train_forward(cu(sig), cu(rgb), deltas, rays, M, N, 1e-4, ws, dp, im)
        g = np.random.RandomState(seed + 12)
        gws = g.normal(size=N).astype(np.float32); gim = g.normal(size=(N, 3)).astype(np.float32)
        gs = torch.zeros(M, device="cuda"); gc = torch.zeros(M, 3, device="cuda")
        ref.composite_rays_train_backward(cu(gws), cu(gim), cu(sig), cu(rgb), deltas, rays, ws, im, M, N, 1e-4, gs, gc)
        torch.cuda.synchronize()
        _, (cgs, cgc) = canonical(r, gs.cpu().numpy(), gc.cpu().numpy())

        # inference loop (renderer_wtmk.py:336-367), no perturbation after step 0 and none here
        iws = torch.zeros(N, device="cuda"); idp = torch.zeros(N, device="cuda"); iim = torch.zeros(N, 3, device="cuda")
        alive = torch.arange(N, dtype=torch.int32, device="cuda"); rt = nears.clone()
        step = iters = 0
        while step < 1024:
            n_alive = alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            Mi = n_alive * n_step; Mi += 128 - Mi % 128
            ix = torch.zeros(Mi, 3, device="cuda"); idr = torch.zeros(Mi, 3, device="cuda"); il = torch.zeros(Mi, 2, device="cuda")
            ref.march_rays(n_alive, n_step, alive, rt, t_o, t_d, bound, dt_gamma, 1024, C, 128, t_b, nears, fars, ix, idr, il,
                           torch.zeros(n_alive, device="cuda"))
            s, c = closed_form_field(ix.cpu().numpy())
            ref.composite_rays(n_alive, n_step, 1e-4, alive, rt, cu(s), cu(c), il, iws, idp, iim)
            alive = alive[alive >= 0]
            step += n_step; iters += 1
        torch.cuda.synchronize()

        out[f"{name}_params"] = json.dumps(list(case))
        out[f"{name}_nears"] = nears.cpu().numpy(); out[f"{name}_fars"] = fars.cpu().numpy()
        out[f"{name}_counts"] = counts; out[f"{name}_total"] = np.int64(total)
        out[f"{name}_xyzs"] = cx; out[f"{name}_dirs"] = cd; out[f"{name}_deltas"] = cl
        out[f"{name}_ws"] = ws.cpu().numpy(); out[f"{name}_depth"] = dp.cpu().numpy(); out[f"{name}_image"] = im.cpu().numpy()
        out[f"{name}_gsig"] = cgs; out[f"{name}_grgb"] = cgc
        out[f"{name}_inf_ws"] = iws.cpu().numpy(); out[f"{name}_inf_depth"] = idp.cpu().numpy()
        out[f"{name}_inf_image"] = iim.cpu().numpy(); out[f"{name}_inf_iters"] = np.int64(iters)
        print(name, "rays", N, "samples", total, "inference iterations", iters)

    rs = np.random.RandomState(0)
    coords = rs.randint(0, 128, size=(4096, 3)).astype(np.int32)
    idx = torch.empty(4096, dtype=torch.int32, device="cuda")
    ref.morton3D(cu(coords), 4096, idx)
    back = torch.empty(4096, 3, dtype=torch.int32, device="cuda")
    ref.morton3D_invert(idx, 4096, back)
    out["morton_coords"] = coords; out["morton_idx"] = idx.cpu().numpy(); out["morton_back"] = back.cpu().numpy()
    grid = rs.uniform(-1, 1, size=(1, 8 * 4096)).astype(np.float32)
    grid[0, :64] = 0.25
    bits = torch.empty(4096, dtype=torch.uint8, device="cuda")
    ref.packbits(cu(grid), 4096, 0.25, bits)
    out["packbits_grid"] = grid; out["packbits_bits"] = bits.cpu().numpy()
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "raymarch_golden.npz"))
