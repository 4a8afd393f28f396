"""Training-time distortion of the rendered watermark blocks (Trainer.distortion_layer, utils_wtmk_disen.py:551-577;
CLI `--distortion`, default 'none'): applied to the clamped block pixels [B, H, W, 3] between the renderer and the HiDDeN
decoder so that the embedded message survives the attack.  Plain torch / torchvision ops on whatever device the pixels
live on, differentiable where the reference's are; same random draws as the reference for the same generator state.

Inside a captured CUDA graph only 'none' is accepted (harness.Scene): four of the five attacks read a random parameter
on the host (`.item()` inside torchvision's get_params / the scale factor), exactly like the reference, and the captured
step was never measured with the fifth ('noise'); all of them run in the eager step."""
import torch
import torch.nn.functional as F

KINDS = ("none", "noise", "rotation", "scaling", "blurring", "brightness")
CAPTURABLE = ("none",)


def _channels_first(fn, pixels):
    return fn(pixels.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)


def _noise(pixels):
    std = torch.sqrt(torch.tensor(0.1))                      # the reference's own expression for the standard deviation
    return pixels + torch.normal(0, std, size=pixels.shape, device=pixels.device)


def _rotation(pixels):
    from torchvision import transforms as T
    rotate = T.RandomRotation(degrees=(-30, 30))             # one angle per block, nearest-neighbour, zero fill
    return _channels_first(lambda x: torch.stack([rotate(block) for block in x]), pixels)


def _scaling(pixels):
    factor = torch.empty(1).uniform_(0.75, 1.25).item()      # one factor per call, drawn on the host
    # a [3, H, W] block handed to 1-D linear interpolation is read as (batch 3, channels H, length W): only the WIDTH is
    # resampled, to floor(W * factor) - the reference's behaviour, kept as it is
    return _channels_first(lambda x: torch.stack([F.interpolate(block, scale_factor=factor, mode="linear") for block in x]),
                           pixels)


def _blurring(pixels):
    from torchvision import transforms as T
    return _channels_first(T.GaussianBlur(kernel_size=3, sigma=(0.01, 0.5)), pixels)    # one sigma for the whole batch


def _brightness(pixels):
    from torchvision import transforms as T
    return _channels_first(T.ColorJitter(brightness=0.5), pixels)                       # one factor for the whole batch


_LAYERS = {"none": lambda p: p, "noise": _noise, "rotation": _rotation, "scaling": _scaling, "blurring": _blurring,
           "brightness": _brightness}


def distortion_layer(pixels, kind="none"):
    """pixels: [B, H, W, 3] in [0, 1] -> distorted pixels, [B, H, W', 3] (W' != W only for 'scaling')."""
    try:
        layer = _LAYERS[kind]
    except KeyError:
        raise ValueError(f"distortion must be one of {KINDS}, got {kind!r}") from None
    return layer(pixels)
