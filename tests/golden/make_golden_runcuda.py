"""Golden vectors for the cuda_ray renderer FROM THE REFERENCE'S OWN NeRFRenderer.run_cuda BODY - both branches.

Run in the build container (needs /root/reference; not needed on the GPU box):
    python tests/golden/make_golden_runcuda.py

nerf/renderer_wtmk.py cannot be imported here (its `import raymarching` builds a CUDA extension), so `run_cuda`
(renderer_wtmk.py:256-377) is cut out of the source text with `ast` and run unmodified on CPU:
  * the `raymarching` module it calls is served by the C oracle (oracle/raymarch_oracle.c, pinned bit-exactly to the reference
    CUDA kernels' outputs: tests/golden/raymarch_golden.npz), behind wrappers with the buffer and padding conventions of
    raymarching/raymarching.py:161-373 (zero-filled sample buffers cut to the counter rounded up to `align`, in-place
    alive-ray state);
  * `self(xyzs, dirs, message)` is the field oracle (C hash oracles + oracle/field_oracle.py; wiring pinned by
    tests/golden/field_golden.npz).
The fixture holds what the reference's orchestration returns for the networks and rays of
tests/test_render_gpu.py::test_fused_renderer_matches_cpu_oracle:
  * eval branch - the host-driven alive-ray loop (n_step schedule, compaction, composite_rays' kill rule, background blend,
    depth normalisation): the frames `nsig_render_rays` must reproduce, INCLUDING the dense early-termination case;
  * training branch (force_all_rays, perturb off): what the oracle chain of __graft_entry__.smoke() stands for.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden_grid import REF  # noqa: E402
from make_golden_field import cut_methods  # noqa: E402
from oracle import cpu as oc  # noqa: E402
from oracle import field_oracle as fo  # noqa: E402
from nerf_signature_b200 import synthetic as syn  # noqa: E402

CASES = {"gain0": (0.0, 1e-4), "gain0.6": (0.6, 1e-4), "gain12_dense": (12.0, 1e-2)}   # sigma_gain, T_thresh
MD, N_RAYS, BOUND, RAY_SEED, MSG_SEED = 8, 600, 1.0, 23, 4


def make_net_cpu(bound, md, table_scale=300.0, seed=0, sigma_gain=0.0):
    """The network tests/test_render_gpu.py::_net builds, kept on the CPU (same seed => same parameters)."""
    from nerf_signature_b200.nerf.network_wtmk_tcnn import NeRFNetwork
    torch.manual_seed(seed)
    net = NeRFNetwork(bound=bound, cuda_ray=True, message_dim=md)
    with torch.no_grad():
        for e in list(net.encoder.embeddings) + list(net.msg_encoder.embeddings):
            e.weight.mul_(table_scale)
        if sigma_gain:
            net.sigma_net.params[2048:2048 + 64] += sigma_gain
    grid = syn.sphere_grid(net.cascade)
    net.density_grid.copy_(torch.from_numpy(grid))
    net.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, 0.5)))
    return net


def case_inputs(name):
    gain, T_thresh = CASES[name]
    net = make_net_cpu(BOUND, MD, sigma_gain=gain)
    rays_o, rays_d = syn.blender_rays(N_RAYS, seed=RAY_SEED)
    msg = np.random.RandomState(MSG_SEED).randint(0, 2, size=MD).astype(np.float32)
    return net, rays_o, rays_d, msg, T_thresh


class FieldOracle:
    """`self(xyzs, dirs, message)` of the renderer: sigma, rgb from the oracle chain; counts the samples it is asked for."""

    def __init__(self, net):
        self.bound = float(net.bound)
        self.base = [e.weight.detach().numpy() for e in net.encoder.embeddings]
        self.msgt = [e.weight.detach().numpy() for e in net.msg_encoder.embeddings]
        self.res, self.msg_res = net.encoder.resolutions, net.msg_encoder.resolution
        self.sp, self.cp = net.sigma_net.params.detach(), net.color_net.params.detach()
        self.evaluated = 0

    def __call__(self, xyzs, dirs, message):
        x = xyzs.numpy()
        self.evaluated += x.shape[0]
        xn = ((x + np.float32(self.bound)) * np.float32(0.5 / self.bound)).astype(np.float32)
        feat = oc.hash_encode_forward(xn, self.base, self.res, 19)
        feat[:, 30:32] += oc.msg_encode_forward(xn, self.msgt, message.numpy(), self.msg_res, 19)
        sig, rgb, _, _ = fo.mlp_forward(torch.from_numpy(feat), dirs, self.sp, self.cp)
        return sig, rgb


class Raymarching:
    """The call surface of raymarching/raymarching.py that run_cuda uses, over the C oracle (CPU tensors)."""

    @staticmethod
    def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
        n, f = oc.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb.numpy(), float(min_near))
        return torch.from_numpy(n), torch.from_numpy(f)

    @staticmethod
    def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False,
                         align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        assert force_all_rays and not perturb
        xyzs, dirs, deltas, rays, cnt = oc.march_rays_train(rays_o.numpy(), rays_d.numpy(), float(bound), bitfield.numpy(), C, H,
                                                            nears.numpy(), fars.numpy(), dt_gamma=float(dt_gamma),
                                                            max_steps=max_steps)
        step_counter.copy_(torch.from_numpy(cnt))
        m = int(cnt[0])
        if align > 0:
            m += align - m % align                       # raymarching.py:226-228
        return (torch.from_numpy(xyzs[:m].copy()), torch.from_numpy(dirs[:m].copy()), torch.from_numpy(deltas[:m].copy()),
                torch.from_numpy(rays))

    @staticmethod
    def composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        ws, depth, img = oc.composite_rays_train_forward(sigmas.numpy(), rgbs.numpy(), deltas.numpy(), rays.numpy(), float(T_thresh))
        return torch.from_numpy(ws), torch.from_numpy(depth), torch.from_numpy(img)

    @staticmethod
    def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, near, far, align=-1, perturb=False,
                   dt_gamma=0, max_steps=1024):
        assert not perturb
        xyzs, dirs, deltas = oc.march_rays(n_alive, n_step, rays_alive.numpy(), rays_t.numpy(), rays_o.numpy(), rays_d.numpy(),
                                           float(bound), bitfield.numpy(), C, H, near.numpy(), far.numpy(), align=align,
                                           dt_gamma=float(dt_gamma), max_steps=max_steps)
        return torch.from_numpy(xyzs), torch.from_numpy(dirs), torch.from_numpy(deltas)

    @staticmethod
    def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        oc.composite_rays(n_alive, n_step, rays_alive.numpy(), rays_t.numpy(), sigmas.numpy(), rgbs.numpy(), deltas.numpy(),
                          weights_sum.numpy(), depth.numpy(), image.numpy(), float(T_thresh))


class StubRenderer:
    def __init__(self, net, training):
        b = float(net.bound)
        self.bound, self.cascade, self.grid_size = b, int(net.cascade), 128
        self.aabb_train = torch.tensor([-b, -b, -b, b, b, b], dtype=torch.float32)
        self.aabb_infer = self.aabb_train.clone()
        self.min_near, self.bg_radius, self.density_scale, self.training = 0.2, 0, 1, training
        self.density_bitfield = net.density_bitfield.detach().clone()
        self.step_counter = torch.zeros(16, 2, dtype=torch.int32)
        self.local_step, self.mean_count = 0, 0
        self.field = FieldOracle(net)

    def __call__(self, xyzs, dirs, message):
        return self.field(xyzs, dirs, message)


def reference_run_cuda():
    fns = cut_methods(os.path.join(REF, "nerf", "renderer_wtmk.py"), {"run_cuda"})
    env = {"torch": torch, "raymarching": Raymarching}
    exec(compile(fns["run_cuda"], "ref:renderer_wtmk", "exec"), env)
    return env["run_cuda"]


def main():
    oc.build()
    torch.set_num_threads(4)
    run_cuda = reference_run_cuda()
    out = {}
    for name in CASES:
        net, rays_o, rays_d, msg, T_thresh = case_inputs(name)
        o, d, m = torch.from_numpy(rays_o)[None], torch.from_numpy(rays_d)[None], torch.from_numpy(msg)
        for mode, training in (("eval", False), ("train", True)):
            stub = StubRenderer(net, training)
            with torch.no_grad():
                r = run_cuda(stub, o, d, m, dt_gamma=0.0, bg_color=1, perturb=False, force_all_rays=True, max_steps=1024,
                             T_thresh=T_thresh)
            out[f"{name}_{mode}_image"] = r["image"].reshape(-1, 3).numpy()
            out[f"{name}_{mode}_depth"] = r["depth"].reshape(-1).numpy()
            out[f"{name}_{mode}_samples_evaluated"] = np.int64(stub.field.evaluated)
            if training:
                out[f"{name}_{mode}_weights_sum"] = r["weights_sum"].numpy()
                out[f"{name}_{mode}_counter"] = stub.step_counter[0].numpy().copy()
            print(name, mode, "samples evaluated", stub.field.evaluated, "image min", float(r["image"].min()))
    np.savez_compressed(os.path.join(HERE, "runcuda_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "runcuda_golden.npz"), os.path.getsize(os.path.join(HERE, "runcuda_golden.npz")))


if __name__ == "__main__":
    main()
