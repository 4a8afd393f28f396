"""HiDDeN message decoder (plain PyTorch; BASELINE north_star allows it to stay so).

Architecture and parameter names follow the reference's nerf/hidden_models.py:16-35,104-137 so
that its checkpoints load: `layers.{k}.layers.{0,1}.*` conv/BN blocks and `linear.*`.
  ConvBNRelu      = Conv2d(3x3, pad 1) -> BatchNorm2d(eps=1e-3, track_running_stats=False) -> GELU
  decoder         = ConvBNRelu(input_ch, C), (num_blocks-1) x ConvBNRelu(C, C), ConvBNRelu(C, bits*red),
                    AdaptiveAvgPool2d(1), Linear(bits*red, bits*red), sum over redundancy
BatchNorm always uses batch statistics (over the message_dim image blocks): sharding the blocks
across ranks would change results, so the multi-GPU path all-gathers pixels first.
"""
import torch
import torch.nn as nn

_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


_CONST = {}


def _mean_std(x):
    """Device-resident normalisation constants, built once per (device, dtype): creating them from Python
    lists on every call is a host-to-device copy per step and cannot be captured in a CUDA graph."""
    key = (x.device, x.dtype)
    if key not in _CONST:
        _CONST[key] = (torch.tensor(_MEAN, dtype=x.dtype, device=x.device).view(-1, 1, 1),
                       torch.tensor(_STD, dtype=x.dtype, device=x.device).view(-1, 1, 1))
    return _CONST[key]


def normalize_img(x):
    """(x - mean) / std over the channel dim of a [..., 3, H, W] tensor (torchvision Normalize)."""
    mean, std = _mean_std(x)
    return (x - mean) / std


def unnormalize_img(x):
    mean, std = _mean_std(x)
    return x * std + mean


class ConvBNRelu(nn.Module):
    def __init__(self, channels_in, channels_out):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Conv2d(channels_in, channels_out, 3, stride=1, padding=1),
            nn.BatchNorm2d(channels_out, eps=1e-3, track_running_stats=False),
            nn.GELU(),
        )

    def forward(self, x):
        return self.layers(x)


class HiddenDecoder_multi_views(nn.Module):
    def __init__(self, num_blocks, num_bits, input_ch, channels, redundancy=1):
        super().__init__()
        layers = [ConvBNRelu(input_ch, channels)]
        for _ in range(num_blocks - 1):
            layers.append(ConvBNRelu(channels, channels))
        layers.append(ConvBNRelu(channels, num_bits * redundancy))
        layers.append(nn.AdaptiveAvgPool2d(output_size=(1, 1)))
        self.layers = nn.Sequential(*layers)
        self.linear = nn.Linear(num_bits * redundancy, num_bits * redundancy)
        self.num_bits = num_bits
        self.redundancy = redundancy

    def forward(self, img_w):
        x = self.layers(img_w)            # b d 1 1
        x = x.squeeze(-1).squeeze(-1)     # b d
        x = self.linear(x)
        x = x.view(-1, self.num_bits, self.redundancy)
        return torch.sum(x, dim=-1)       # b k


def get_hidden_decoder_multi_views(num_bits, redundancy=1, num_blocks=7, input_ch=3, channels=64):
    return HiddenDecoder_multi_views(num_blocks=num_blocks, num_bits=num_bits, input_ch=input_ch,
                                     channels=channels, redundancy=redundancy)
