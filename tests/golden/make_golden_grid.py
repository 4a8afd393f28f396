"""Golden vectors for the occupancy-grid update, mark_untrained_grid and get_rays FROM THE REFERENCE FUNCTIONS.

Run in the build container (needs /root/reference; not needed on the GPU box):
    python tests/golden/make_golden_grid.py

nerf/renderer_wtmk.py and nerf/utils_wtmk_disen.py cannot be imported here (trimesh, tensorboardX, lpips, ... are not
installed), so the three function definitions are cut out of the source text with `ast`, compiled unmodified and
run on CPU against a stub renderer object:
  * `raymarching.morton3D / morton3D_invert / packbits` are served by the numpy restatements in oracle/grid_oracle.py
    (themselves checked against the reference CUDA kernels' goldens in tests/test_oracle_cpu.py);
  * `self.density` is an analytic density field whose inputs and outputs are recorded, so the fixture pins the
    ORCHESTRATION (which cells, where inside them, EMA, mean, threshold, bitfield) independently of the network;
  * `torch.rand_like / torch.randint` are recorded so the oracle can be fed the same draws.
The grid size is 16 (the methods read self.grid_size) to keep the fixture small.  Nothing is copied into this repo.
"""
import ast
import os
import sys

import numpy as np
import torch

REF = os.environ.get("NSIG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import grid_oracle as go  # noqa: E402


def cut_functions(path, names):
    """Source text of the named (possibly nested in a class) function definitions, decorators included."""
    src = open(path).read()
    tree = ast.parse(src)
    lines = src.split("\n")
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            first = min([node.lineno] + [d.lineno for d in node.decorator_list])
            text = "\n".join(lines[first - 1:node.end_lineno])
            import textwrap
            out[node.name] = textwrap.dedent(text)
    return out


class TorchRecorder:
    """`torch` as the reference functions see it: everything forwards, random draws are recorded."""

    def __init__(self):
        self.rand_like_log, self.randint_log = [], []

    def __getattr__(self, name):
        return getattr(torch, name)

    def rand_like(self, t, **kw):
        r = torch.rand_like(t, **kw)
        self.rand_like_log.append(r.clone())
        return r

    def randint(self, *a, **kw):
        r = torch.randint(*a, **kw)
        self.randint_log.append(r.clone())
        return r


class Raymarching:
    @staticmethod
    def morton3D(coords):
        return torch.from_numpy(go.morton3D(coords.numpy()).astype(np.int32))

    @staticmethod
    def morton3D_invert(indices):
        return torch.from_numpy(go.morton3D_invert(indices.numpy()).astype(np.int32))

    @staticmethod
    def packbits(grid, thresh, bitfield=None):
        return torch.from_numpy(go.packbits(grid.numpy(), thresh))


def analytic_density(x):
    """A smooth blob + a ridge, fp32 torch ops; values span the threshold."""
    r2 = (x * x).sum(-1)
    return (12.0 * torch.exp(-3.0 * r2) + 2.0 * torch.exp(-8.0 * (x[:, 0] - 0.5 * x[:, 1]) ** 2)).float()


class Stub:
    def __init__(self, bound, H, density_thresh, density_scale=1.0):
        import math
        self.cuda_ray = True
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = H
        self.density_scale = density_scale
        self.density_thresh = density_thresh
        self.density_grid = torch.zeros([self.cascade, H ** 3])
        self.density_bitfield = torch.zeros(self.cascade * H ** 3 // 8, dtype=torch.uint8)
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter = torch.zeros(16, 2, dtype=torch.int32)
        self.mean_count = 0
        self.local_step = 0
        self.calls = []

    def density(self, x, message=None):
        s = analytic_density(x)
        self.calls.append((x.clone(), s.clone()))
        return {"sigma": s.clone()}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    out = {}

    # ---------------- update_extra_state + mark_untrained_grid (nerf/renderer_wtmk.py) ----------------
    fns = cut_functions(os.path.join(REF, "nerf", "renderer_wtmk.py"), {"update_extra_state", "mark_untrained_grid"})
    fns.update(cut_functions(os.path.join(REF, "nerf", "utils_wtmk_disen.py"), {"custom_meshgrid"}))
    from packaging import version as pver
    rec = TorchRecorder()
    ns = {"torch": rec, "raymarching": Raymarching, "np": np, "pver": pver}
    for name in ("custom_meshgrid", "update_extra_state", "mark_untrained_grid"):
        exec(compile(fns[name], f"<reference {name}>", "exec"), ns)

    H = 16
    for tag, bound in (("b1", 1), ("b2", 2)):
        stub = Stub(bound, H, density_thresh=10.0)
        C = stub.cascade
        # start from a grid with some untrained (-1) and some occupied cells
        g0 = torch.rand(C, H ** 3) * 3.0
        g0[:, ::7] = -1.0
        g0[:, 1::5] = 0.0
        stub.density_grid = g0.clone()
        stub.step_counter[:, 0] = torch.arange(16, dtype=torch.int32) * 1000 + 17
        stub.local_step = 5
        out[f"{tag}_grid0"] = g0.numpy()

        # ---- full update (iter_density < 16)
        rec.rand_like_log.clear(); rec.randint_log.clear(); stub.calls.clear()
        ns["update_extra_state"](stub, None)
        out[f"{tag}_full_noise"] = np.stack([r.numpy() for r in rec.rand_like_log])          # [C, H^3, 3] meshgrid order
        out[f"{tag}_full_xyz"] = np.stack([c[0].numpy() for c in stub.calls])
        out[f"{tag}_full_sigma"] = np.stack([c[1].numpy() for c in stub.calls])
        out[f"{tag}_full_grid"] = stub.density_grid.numpy().copy()
        out[f"{tag}_full_bitfield"] = stub.density_bitfield.numpy().copy()
        out[f"{tag}_full_mean"] = np.float32(stub.mean_density)
        out[f"{tag}_full_mean_count"] = np.int64(stub.mean_count)

        # ---- partial update (iter_density >= 16)
        stub.iter_density = 16
        stub.local_step = 3
        rec.rand_like_log.clear(); rec.randint_log.clear(); stub.calls.clear()
        g1 = stub.density_grid.clone()
        ns["update_extra_state"](stub, None)
        # per cascade the reference draws: randint coords [N,3], randint rand_mask [N]; rand_like noise [2N,3]
        out[f"{tag}_part_grid1"] = g1.numpy()
        out[f"{tag}_part_coords"] = np.stack([rec.randint_log[2 * c].numpy() for c in range(C)])
        out[f"{tag}_part_randmask"] = np.stack([rec.randint_log[2 * c + 1].numpy() for c in range(C)])
        out[f"{tag}_part_noise"] = np.stack([r.numpy() for r in rec.rand_like_log])
        out[f"{tag}_part_xyz"] = np.stack([c[0].numpy() for c in stub.calls])
        out[f"{tag}_part_sigma"] = np.stack([c[1].numpy() for c in stub.calls])
        out[f"{tag}_part_grid"] = stub.density_grid.numpy().copy()
        out[f"{tag}_part_bitfield"] = stub.density_bitfield.numpy().copy()
        out[f"{tag}_part_mean"] = np.float32(stub.mean_density)

        # ---- mark_untrained_grid
        stub2 = Stub(bound, H, density_thresh=10.0)
        rs = np.random.RandomState(5 + bound)
        poses = []
        for _ in range(3):
            th, ph = rs.uniform(np.pi / 3, 2 * np.pi / 3), rs.uniform(0, 2 * np.pi)
            eye = 1.6 * bound * np.array([np.sin(th) * np.sin(ph), np.cos(th), np.sin(th) * np.cos(ph)])
            fwd = -eye / np.linalg.norm(eye)
            right = np.cross(np.array([0.0, 1.0, 0.0]), fwd); right /= np.linalg.norm(right)
            up = np.cross(fwd, right)
            P = np.eye(4, dtype=np.float32)
            P[:3, 0], P[:3, 1], P[:3, 2], P[:3, 3] = right, up, fwd, eye
            poses.append(P)
        poses = np.stack(poses).astype(np.float32)
        intrinsic = (70.0, 66.0, 16.0, 15.0)  # narrow frustum so that part of the grid stays unseen
        ns["mark_untrained_grid"](stub2, poses, intrinsic)
        out[f"{tag}_mark_poses"] = poses
        out[f"{tag}_mark_intrinsic"] = np.array(intrinsic, np.float32)
        out[f"{tag}_mark_grid"] = stub2.density_grid.numpy().copy()

    # ---------------- get_rays (nerf/utils_wtmk_disen.py) ----------------
    fns = cut_functions(os.path.join(REF, "nerf", "utils_wtmk_disen.py"), {"get_rays", "custom_meshgrid"})
    rec2 = TorchRecorder()
    ns2 = {"torch": rec2, "np": np, "pver": pver}
    exec(compile(fns["custom_meshgrid"], "<reference custom_meshgrid>", "exec"), ns2)
    exec(compile(fns["get_rays"], "<reference get_rays>", "exec"), ns2)
    rs = np.random.RandomState(3)
    poses = out["b1_mark_poses"][:3]
    Hh, Ww = 36, 48
    intr = np.array([55.5, 54.25, 24.0, 18.0], np.float32)
    res = ns2["get_rays"](torch.from_numpy(poses), intr, Hh, Ww, 64)
    out["rays_poses"], out["rays_intr"], out["rays_HW"] = poses, intr, np.array([Hh, Ww])
    out["rays_inds"] = res["inds"].numpy().copy()
    out["rays_o"], out["rays_d"] = res["rays_o"].numpy().copy(), res["rays_d"].numpy().copy()
    res = ns2["get_rays"](torch.from_numpy(poses), intr, Hh, Ww, -1)
    out["rays_all_o"], out["rays_all_d"] = res["rays_o"].numpy().copy(), res["rays_d"].numpy().copy()
    res = ns2["get_rays"](torch.from_numpy(poses), intr, Hh, Ww, 64, None, 4)
    out["rays_patch_inds"] = res["inds"].numpy().copy()
    out["rays_patch_draws"] = np.stack([r.numpy() for r in rec2.randint_log[-2:]])

    path = os.path.join(HERE, "grid_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
