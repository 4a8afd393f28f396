"""Golden vectors for one watermark training step FROM THE REFERENCE'S OWN Trainer.train_step BODY.

Run in the build container (needs /root/reference; not needed on the GPU box):
    python tests/golden/make_golden_trainstep.py

nerf/utils_wtmk_disen.py cannot be imported (tensorboardX, lpips, trimesh, ... are not installed), so `train_step` and
`distortion_layer` are cut out of its source text with `ast` and run unmodified on CPU on a stub Trainer:
  * `self.loss_w` is compiled from the source segment of the reference's own `bce` lambda (utils_wtmk_disen.py:441),
    `self.criterion` is what main_nerf_wtmk.py:106 passes (MSELoss(reduction='none')), distortion 'none' (CLI default);
  * `self.model.msg_decoder` / `normalization` are the reference's nerf/hidden_models.py, imported unmodified;
  * `self.model.render` is served by oracle/torch_port.render_run, itself pinned to the reference's NeRFRenderer.run
    (tests/golden/run_golden.npz).
The fixture pins the step's ORCHESTRATION (utils_wtmk_disen.py:579-646): two render passes with the same message, clamp of
the block pixels, 'b h w c -> b c h w', normalisation, decoder, lossi = mean of the element-wise MSE over the content rays,
lossw = BCE on 10 x logits against message[:, None], loss = lambda_w lossw + lambda_i lossi - and, through autograd, the
gradients every implementation of the step must deliver to the message tables and the decoder.
"""
import argparse
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
from einops import rearrange

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden_grid import REF  # noqa: E402
from make_golden_field import cut_methods  # noqa: E402
from oracle import torch_port as tp  # noqa: E402
from nerf_signature_b200 import synthetic as syn  # noqa: E402

LOG2_T, NUM_STEPS = 12, 32
CASES = {"md4": (4, 3, 3, 24, 0.005, 1.0, 51), "md8_lw1": (8, 2, 3, 16, 1.0, 1.0, 52)}   # md, pH, pW, content rays, lambda_w, lambda_i, seed


def case_inputs(name):
    md, ph, pw, n, lw, li, seed = CASES[name]
    field = tp.PortField(bound=1.0, message_dim=md, log2_T=LOG2_T, seed=seed, train_msg=True)
    field.base_tables = [t * 3e3 for t in field.base_tables]
    field.msg_tables = [(t.detach() * 3e3).requires_grad_(True) for t in field.msg_tables]
    bo, bd = syn.blender_rays(md * ph * pw, seed=seed + 1)
    co, cd = syn.blender_rays(n, seed=seed + 2)
    g = torch.Generator().manual_seed(seed)
    batch = {"rays_o_block": torch.from_numpy(bo).view(md, ph, pw, 3), "rays_d_block": torch.from_numpy(bd).view(md, ph, pw, 3),
             "rays_o": torch.from_numpy(co), "rays_d": torch.from_numpy(cd), "gt": torch.rand(n, 3, generator=g)}
    message = torch.randint(0, 2, (md,), generator=g).float()
    return field, batch, message, lw, li, seed


def summarize(field, decoder, loss_terms):
    out = {"losses": np.array([float(v) for v in loss_terms], np.float64)}
    out["table_grad_sums"] = np.array([0.0 if t.grad is None else float(t.grad.double().abs().sum()) for t in field.msg_tables])
    out["table_grad_0"] = next(t.grad.numpy() for t in field.msg_tables if t.grad is not None)
    out["decoder_grad_norms"] = np.array([float(p.grad.norm()) for p in decoder.parameters()], np.float64)
    return out


def main():
    torch.set_num_threads(1)
    src_path = os.path.join(REF, "nerf", "utils_wtmk_disen.py")
    fns = cut_methods(src_path, {"train_step", "distortion_layer"})
    env = {"torch": torch, "F": F, "rearrange": rearrange}
    for f in fns.values():
        exec(compile(f, "ref:utils_wtmk_disen", "exec"), env)
    # the reference's own bce lambda (defined inside Trainer.__init__)
    src = open(src_path).read()
    lam = [ast.get_source_segment(src, n) for n in ast.walk(ast.parse(src))
           if isinstance(n, ast.Lambda) and "binary_cross_entropy_with_logits" in ast.get_source_segment(src, n)]
    assert len(lam) == 1
    loss_w = eval(lam[0], {"torch": torch, "F": F})
    spec = importlib.util.spec_from_file_location("ref_hidden_models", os.path.join(REF, "nerf", "hidden_models.py"))
    ref_hidden = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_hidden)

    out = {}
    for name in CASES:
        field, batch, message, lw, li, seed = case_inputs(name)
        torch.manual_seed(seed)
        decoder = ref_hidden.get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=8, input_ch=3, channels=64)

        def render(rays_o, rays_d, msg, staged=False, bg_color=None, perturb=False, force_all_rays=False, **kw):
            prefix = rays_o.shape[:-1]
            img, _ = tp.render_run(field, rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), msg, num_steps=NUM_STEPS, bg_color=bg_color)
            return {"image": img.view(*prefix, 3)}

        self = types.SimpleNamespace()
        self.model = types.SimpleNamespace(bg_radius=0, render=render, msg_decoder=decoder, normalization=ref_hidden.normalize_img)
        self.opt = argparse.Namespace(color_space="srgb")
        self.distortion, self.lambda_w, self.lambda_i = "none", lw, li
        self.criterion = torch.nn.MSELoss(reduction="none")
        self.loss_w = loss_w
        self.distortion_layer = lambda x: env["distortion_layer"](self, x)
        md, ph, pw = batch["rays_o_block"].shape[:3]
        data = {"watermark": {"images": torch.zeros(md, ph, pw, 3), "rays_o_block": batch["rays_o_block"],
                              "rays_d_block": batch["rays_d_block"]},
                "content": {"rays_o": batch["rays_o"][None], "rays_d": batch["rays_d"][None], "images": batch["gt"][None].clone()}}
        pred_rgb, gt_rgb, content_pred, lossi, lossw, loss = env["train_step"](self, data, message)
        loss.backward()
        res = summarize(field, decoder, (loss, lossi, lossw))
        res["pred_rgb"] = pred_rgb.detach().numpy()
        for k, v in res.items():
            out[f"{name}_{k}"] = v
        print(name, "loss / lossi / lossw", res["losses"], "tables with gradient", int((res["table_grad_sums"] > 0).sum()))
    np.savez_compressed(os.path.join(HERE, "trainstep_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "trainstep_golden.npz"), os.path.getsize(os.path.join(HERE, "trainstep_golden.npz")))


if __name__ == "__main__":
    main()
