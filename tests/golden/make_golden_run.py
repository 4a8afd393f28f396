"""Golden vectors for the non-cuda_ray renderer FROM THE REFERENCE'S OWN NeRFRenderer.run.

Run in the build container (needs /root/reference; not needed on the GPU box):
    python tests/golden/make_golden_run.py

oracle/torch_port.render_run restates NeRFRenderer.run (nerf/renderer_wtmk.py:125-253); it is the CPU baseline of
bench.py (`cpu_baseline`, `--impl reference`, BASELINE configs[0]) and the oracle of the product's `run`
(tests/test_render_gpu.py).  nerf/renderer_wtmk.py cannot be imported here (it imports the CUDA extension, trimesh,
...), so `run` and `sample_pdf` are cut out of its source text with `ast`, compiled unmodified and run on CPU against a
stub renderer object:
  * `raymarching.near_far_from_aabb` is served by the C oracle (pinned to the reference CUDA kernel's goldens in
    tests/test_oracle_cpu.py);
  * `self.density` / `self.color` are oracle/torch_port.PortField's (encoders pinned bit-exactly to the reference modules;
    `color` wrapped with the reference's mask semantics, network_wtmk_tcnn.py:150-176), rebuilt in the test from the seed.
What the fixture pins is therefore the ORCHESTRATION of run(): sample placement, the manual clip, deltas and the last
interval, alpha compositing with the 1e-15 guard, the 1e-4 colour mask, depth normalisation, the background blend - with
and without a message, and with the hierarchical resampling (sample_pdf) of upsample_steps > 0 for reference.
Nothing is copied into this repo.
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

from make_golden_grid import REF, cut_functions  # noqa: E402
from oracle import cpu as oc  # noqa: E402
from oracle import torch_port as tp  # noqa: E402
from nerf_signature_b200 import synthetic as syn  # noqa: E402

LOG2_T = 14   # small tables: the fixture pins run(), the encoders have their own goldens


class Raymarching:
    @staticmethod
    def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
        n, f = oc.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb.numpy(), float(min_near))
        return torch.from_numpy(n), torch.from_numpy(f)


def stub_renderer(field, training=False):
    self = types.SimpleNamespace()
    b = field.bound
    self.aabb_train = torch.tensor([-b, -b, -b, b, b, b], dtype=torch.float32)
    self.aabb_infer = self.aabb_train.clone()
    self.min_near, self.density_scale, self.bg_radius, self.training = 0.2, 1, 0, training
    self.density = lambda x, message=None: field.density(x, message)

    def color(x, d, mask=None, geo_feat=None, **kwargs):   # mask semantics of network_wtmk_tcnn.py:150-176
        if mask is None:
            return field.color(d, geo_feat)
        rgbs = torch.zeros(mask.shape[0], 3, dtype=x.dtype)
        if mask.any():
            rgbs[mask] = field.color(d[mask], geo_feat[mask])
        return rgbs
    self.color = color
    return self


CASES = {
    # name: (bound, message_dim, n_rays, num_steps, upsample_steps, field seed, ray seed)
    "clean": (1.0, 0, 48, 64, 0, 3, 5),
    "wtmk": (1.0, 4, 48, 64, 0, 4, 6),
    "wtmk_bound2": (2.0, 4, 40, 48, 0, 5, 7),
    "wtmk_upsample": (1.0, 4, 32, 48, 24, 6, 8),
}


def make_field(bound, md, seed, table_scale=3e3):
    """PortField with tables scaled so the random-init scene is visible (sigma, colours away from their init values)."""
    f = tp.PortField(bound=bound, message_dim=md, log2_T=LOG2_T, seed=seed, train_msg=False)
    f.base_tables = [t * table_scale for t in f.base_tables]
    f.msg_tables = [t * table_scale for t in f.msg_tables]
    return f


def case_inputs(name):
    bound, md, n, steps, up, fseed, rseed = CASES[name]
    o, d = syn.blender_rays(n, seed=rseed)
    if bound > 1:
        o = o * np.float32(1.5)
    msg = torch.from_numpy(np.random.RandomState(rseed).randint(0, 2, size=md).astype(np.float32)) if md else None
    return make_field(bound, md, fseed), torch.from_numpy(o), torch.from_numpy(d), msg, steps, up


def pdf_inputs():
    """[N, B] bin edges and [N, B-1] weights: generic rays, an all-zero ray (uniform fallback through the 1e-5 guard), a ray
    with all its weight in one bin (degenerate spans), a ray with weight in the first and last bin only."""
    rs = np.random.RandomState(9)
    N, B = 12, 17
    bins = np.sort(rs.uniform(0.2, 4.0, size=(N, B)).astype(np.float32), axis=1)
    w = rs.uniform(0, 1, size=(N, B - 1)).astype(np.float32) ** 4
    w[0] = 0
    w[1] = 0
    w[1, 5] = 1
    w[2] = 0
    w[2, 0] = w[2, -1] = 0.5
    return torch.from_numpy(bins), torch.from_numpy(w)


def main():
    oc.build()
    torch.set_num_threads(4)
    fns = cut_functions(os.path.join(REF, "nerf", "renderer_wtmk.py"), {"run", "sample_pdf"})
    env = {"torch": torch, "raymarching": Raymarching}
    exec(compile(fns["sample_pdf"], "ref:sample_pdf", "exec"), env)
    exec(compile(fns["run"], "ref:run", "exec"), env)
    out = {}
    for name in CASES:
        field, o, d, msg, steps, up = case_inputs(name)
        with torch.no_grad():
            r = env["run"](stub_renderer(field), o[None], d[None], msg, num_steps=steps, upsample_steps=up, bg_color=1,
                           perturb=False)
        out[f"{name}_image"] = r["image"].reshape(-1, 3).numpy()
        out[f"{name}_depth"] = r["depth"].reshape(-1).numpy()
        out[f"{name}_weights_sum"] = r["weights_sum"].reshape(-1).numpy()
        print(name, "image range", float(r["image"].min()), float(r["image"].max()), "ws max", float(r["weights_sum"].max()))
    # ---- sample_pdf alone (the product's nerf.renderer_wtmk.sample_pdf is plain torch and is compared on CPU) ----
    bins, weights = pdf_inputs()
    out["pdf_bins"], out["pdf_weights"] = bins.numpy(), weights.numpy()
    out["pdf_det"] = env["sample_pdf"](bins, weights, 24, det=True).numpy()
    torch.manual_seed(1234)
    out["pdf_rand_seed1234"] = env["sample_pdf"](bins, weights, 24, det=False).numpy()
    np.savez_compressed(os.path.join(HERE, "run_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "run_golden.npz"))


if __name__ == "__main__":
    main()
