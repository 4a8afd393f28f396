"""Golden vectors for the HiDDeN message decoder FROM THE REFERENCE MODULE (nerf/hidden_models.py, imported unmodified).

Run in the build container (needs /root/reference and torchvision; not needed on the GPU box):
    python tests/golden/make_golden_decoder.py

nerf_signature_b200/nerf/hidden_models.py restates the decoder as a plain PyTorch module; it is the oracle the fused
decoder kernels (csrc/decoder.cu) are tested against in tests/test_decoder_gpu.py, so it is pinned here to the
reference's own module: same parameter names and shapes, same initial values under the same seed, and - on seeded
inputs of the training step's shape (message_dim blocks of 12x12 pixels, utils_wtmk_disen.py:592-595) - the same
logits, the same loss (BCE on 10 x logits, utils_wtmk_disen.py:640) and the same gradients.
"""
import importlib.util
import os

import numpy as np
import torch

REF = os.environ.get("NSIG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {"md32_12x12": (32, 12, 12, 0), "md48_12x12": (48, 12, 12, 1), "md8_23x31": (8, 23, 31, 2)}   # 360 blocks: 756/32 x 1008/32


def case_inputs(name):
    md, h, w, seed = CASES[name]
    g = torch.Generator().manual_seed(100 + seed)
    pred = torch.rand(md, h, w, 3, generator=g)                     # rendered block pixels, clamped to [0, 1]
    message = torch.randint(0, 2, (md,), generator=g).float()
    return seed, pred, message


def run_case(module, name):
    """logits, loss and gradients of `module` (a hidden_models namespace) on the case's inputs."""
    seed, pred, message = case_inputs(name)
    torch.manual_seed(seed)
    dec = module.get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=8, input_ch=3, channels=64)
    pred = pred.clone().requires_grad_(True)
    logits = dec(module.normalize_img(pred.permute(0, 3, 1, 2)))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits * 10.0, message.unsqueeze(-1), reduction="mean")
    loss.backward()
    return dec, {"logits": logits.detach().numpy(), "loss": np.float32(loss.item()), "dpred": pred.grad.numpy(),
                 "grad_norms": np.array([float(p.grad.norm()) for p in dec.parameters()], np.float32),
                 "grad_linear_w": dec.linear.weight.grad.numpy(), "grad_conv0_w": dec.layers[0].layers[0].weight.grad.numpy(),
                 "param_sums": np.array([float(p.detach().double().sum()) for p in dec.parameters()], np.float64)}


def main():
    torch.set_num_threads(1)
    spec = importlib.util.spec_from_file_location("ref_hidden_models", os.path.join(REF, "nerf", "hidden_models.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for name in CASES:
        dec, res = run_case(ref, name)
        for k, v in res.items():
            out[f"{name}_{k}"] = v
        out[f"{name}_keys"] = np.array(list(dec.state_dict().keys()))
        out[f"{name}_shapes"] = np.array([str(tuple(v.shape)) for v in dec.state_dict().values()])
        print(name, "loss", float(res["loss"]), "logit range", float(res["logits"].min()), float(res["logits"].max()))
    np.savez_compressed(os.path.join(HERE, "decoder_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "decoder_golden.npz"))


if __name__ == "__main__":
    main()
