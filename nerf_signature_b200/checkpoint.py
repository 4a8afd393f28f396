"""Checkpoint compatibility with the reference Trainer (nerf/utils_wtmk_disen.py:1385-1517; README.md:33,40: a CLEAN
torch-ngp model is trained first, its `.pth` is then loaded into the watermark network with strict=False and frozen).

File layout kept (torch.save of a dict):
    {'epoch', 'global_step', 'stats', 'mean_count', 'mean_density', 'model': state_dict,
     ['optimizer', 'lr_scheduler', 'scaler' when full=True]}
or a bare state_dict.  Model keys are the reference's: `encoder.embeddings.{l}.weight`, `sigma_net.params`,
`color_net.params`, `density_grid`, `density_bitfield`, `step_counter`, `aabb_train/infer`, and for the watermark
network `msg_encoder.embeddings.{i}.weight`, `msg_decoder.*`.

tiny-cuda-nn parameter layout -> this repo's (ASSUMED: tiny-cuda-nn is neither vendored nor installed, no real checkpoint
is available offline - SURVEY 8c/8f; the assumption is stated here and pinned by tests/test_checkpoint_*.py against a
torch emulation of it):
  * `FullyFusedMLP.params` is ONE flat fp32 (or fp16) vector: the bias-free weight matrices in layer order, each stored
    row-major as [out, in], input width padded to a multiple of 16, output width padded to 16:
        sigma_net: [64,32] [16,64]            = 3072 values
        color_net: [64,32] [64,64] [16,64]    = 7168 values
    - identical to FusedMLP.params here, so the PERMUTATION IS THE IDENTITY;
  * padding: tiny-cuda-nn feeds the constant 1 into padded INPUT slots (the colour net's 32nd input: 16 SH + 15 geo + pad),
    which makes weight column 31 of its first matrix a bias.  The kernels here define that slot as 0 (and ignore column
    31).  The conversion folds the bias into SH component 0, which is the constant 0.28209479 for every direction:
        W0'[:, 0] = W0[:, 0] + W0[:, 31] / fp16(0.28209479),   W0'[:, 31] = 0
    (exact in real arithmetic; the merged weight is rounded to fp16 once more, <= 2^-11 relative).  Padded OUTPUT rows
    (colour rows 3..15) produce values nobody reads in either implementation.
"""
import torch

SH_C0_FP16 = float(torch.tensor(0.28209479177387814).half())   # the value the kernels multiply column 0 with
SIGMA_PARAMS, COLOR_PARAMS = 3072, 7168


def convert_tcnn_params(flat, kind):
    """tiny-cuda-nn flat `params` -> this repo's FusedMLP.params (fp32).  kind: 'sigma' | 'color'."""
    want = SIGMA_PARAMS if kind == "sigma" else COLOR_PARAMS
    flat = flat.detach().reshape(-1).float()
    if flat.numel() != want:
        raise ValueError(f"{kind}_net.params has {flat.numel()} values, expected {want} "
                         "(64-wide FullyFusedMLP of nerf/network_wtmk_tcnn.py:52-88)")
    out = flat.clone()
    if kind == "color":
        w0 = out[:2048].view(64, 32)
        w0[:, 0] += w0[:, 31] / SH_C0_FP16     # constant-1 padding input == bias == multiple of the constant SH band
        w0[:, 31] = 0.0
    return out


def _convert_state_dict(sd, tcnn):
    sd = dict(sd)
    if tcnn:
        for key, kind in (("sigma_net.params", "sigma"), ("color_net.params", "color")):
            if key in sd:
                sd[key] = convert_tcnn_params(sd[key], kind)
    return sd


def load_checkpoint(model, checkpoint, optimizer=None, lr_scheduler=None, scaler=None, model_only=False,
                    map_location=None, tcnn=True):
    """Trainer.load_checkpoint (utils_wtmk_disen.py:1455-1517) for this repo's networks.

    checkpoint: path or already-loaded dict.  tcnn=True converts `sigma_net.params` / `color_net.params` from
    tiny-cuda-nn's convention (see module docstring); pass tcnn=False for checkpoints written by save_checkpoint here.
    Model weights load with strict=False (a clean checkpoint leaves msg_encoder / msg_decoder at their initial values).
    Returns {'missing_keys', 'unexpected_keys', 'epoch', 'global_step', 'stats'}."""
    ckpt = torch.load(checkpoint, map_location=map_location, weights_only=False) if isinstance(checkpoint, (str, bytes)) \
        or hasattr(checkpoint, "__fspath__") else checkpoint
    info = {"missing_keys": [], "unexpected_keys": [], "epoch": None, "global_step": None, "stats": None}
    if "model" not in ckpt:                       # bare state dict (reference L1469-1472): strict load
        model.load_state_dict(_convert_state_dict(ckpt, tcnn))
        _invalidate(model)
        return info
    native = bool(ckpt.get("nsig_native_params", False))
    res = model.load_state_dict(_convert_state_dict(ckpt["model"], tcnn and not native), strict=False)
    info["missing_keys"], info["unexpected_keys"] = list(res.missing_keys), list(res.unexpected_keys)
    _invalidate(model)
    if getattr(model, "cuda_ray", False):
        if "mean_count" in ckpt:
            model.mean_count = ckpt["mean_count"]
        if "mean_density" in ckpt:
            model.mean_density = ckpt["mean_density"]
    if model_only:
        return info
    info.update(epoch=ckpt.get("epoch"), global_step=ckpt.get("global_step"), stats=ckpt.get("stats"))
    for obj, key in ((optimizer, "optimizer"), (lr_scheduler, "lr_scheduler"), (scaler, "scaler")):
        if obj is not None and key in ckpt:
            obj.load_state_dict(ckpt[key])
    return info


def _invalidate(model):
    """Derived device state (fp16 weight copies, half2 shadow tables, cached summed message table) follows the
    parameters' versions; loading bumps them, dropping the caches here makes that explicit."""
    for m in model.modules():
        if hasattr(m, "_half"):
            m._half, m._half_key = None, None
        if hasattr(m, "_shadow_key"):
            m._shadow_key = None
    if hasattr(model, "_S_cache"):
        model._S_cache = None


def save_checkpoint(model, path, epoch=0, global_step=0, stats=None, optimizer=None, lr_scheduler=None, scaler=None,
                    full=False):
    """Trainer.save_checkpoint (utils_wtmk_disen.py:1385-1431).  MLP parameters are written in this repo's native
    parameterisation (padding input = 0) and flagged, so load_checkpoint does not convert them again."""
    state = {"epoch": epoch, "global_step": global_step, "stats": stats if stats is not None else {},
             "nsig_native_params": True}
    if getattr(model, "cuda_ray", False):
        state["mean_count"] = model.mean_count
        state["mean_density"] = model.mean_density
    if full:
        for obj, key in ((optimizer, "optimizer"), (lr_scheduler, "lr_scheduler"), (scaler, "scaler")):
            if obj is not None:
                state[key] = obj.state_dict()
    state["model"] = model.state_dict()
    torch.save(state, path)
    return path
