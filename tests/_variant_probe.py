"""Helper of tests/test_optin_variants_gpu.py (not a test): run one C-ABI entry point on seeded inputs and print sha256 digests
of everything it wrote.  The kernel variant is chosen by the environment of this process."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_signature_b200 import _lib, harness  # noqa: E402

P = _lib.ptr
dev = torch.device("cuda:0")


def digest(name, *tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    print("digest", name, h.hexdigest())


def march():
    from nerf_signature_b200 import raymarching
    from nerf_signature_b200.synthetic import packbits_np
    for cfg_name, n_rays, perturb in (("blender_wtmk", 4099, False), ("360_wtmk", 1537, True)):
        cfg = dict(harness.CONFIGS[cfg_name])
        cascade = 1 + int(np.ceil(np.log2(cfg["bound"])))
        grid = harness.occupancy(cfg, cascade, seed=3)
        bitfield = torch.from_numpy(packbits_np(grid, 0.5)).to(dev)
        b = harness.make_batch(cfg, seed=5, num_rays=n_rays)
        o = torch.from_numpy(b["rays_o"]).to(dev).reshape(-1, 3).contiguous()
        d = torch.from_numpy(b["rays_d"]).to(dev).reshape(-1, 3).contiguous()
        bound = float(cfg["bound"])
        aabb = torch.tensor([-bound, -bound, -bound, bound, bound, bound], dtype=torch.float32, device=dev)
        nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        torch.manual_seed(1)
        xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, bound, bitfield, cascade, 128, nears, fars, counter, 64,
                                                                perturb, 128, True, 0.0, 1024)
        torch.cuda.synchronize()
        m = int(counter[0])
        digest(f"march {cfg_name}", xyzs[:m], dirs[:m], deltas[:m], rays, counter)


def adam():
    md, log2_T = 6, 14
    n = 2 << log2_T
    gg = torch.Generator(device="cuda").manual_seed(11)
    tabs = [torch.randn(n, device=dev, generator=gg) * 1e-2 for _ in range(2 * md)]
    ms = [torch.randn(n, device=dev, generator=gg) * 1e-3 for _ in range(2 * md)]
    vs = [torch.rand(n, device=dev, generator=gg) * 1e-6 for _ in range(2 * md)]
    ptrs = torch.tensor([[t.data_ptr() for t in grp] for grp in (tabs, ms, vs)], dtype=torch.int64, device=dev)
    steps = torch.arange(2 * md, dtype=torch.float32, device=dev) + 3.0
    coef = torch.zeros(2 * md, 2, dtype=torch.float32, device=dev)
    G = torch.randn(n, device=dev, generator=gg) * 65.536
    msg = torch.tensor([1, 0, 1, 1, 0, 0], dtype=torch.float32, device=dev)
    scale = torch.tensor([65536.0], device=dev)
    finf = torch.zeros(1, device=dev)
    for lo, cnt in ((0, 0), (4096, 8192), (1000, 1004)):   # whole tables, an aligned slice, a ragged slice (tail chunk)
        _lib.call("nsig_msg_adam_step", P(ptrs), 2 * md, md, P(msg), P(G), P(steps), P(coef), P(scale), P(finf), 1e-2, 0.9,
                  0.99, 1e-15, log2_T, None, lo, cnt, 0)
    torch.cuda.synchronize()
    digest("adam state", *(tabs + ms + vs), steps, coef)


if __name__ == "__main__":
    {"march": march, "adam": adam}[sys.argv[1]]()
