"""trunc_exp with the reference's semantics (activation.py:5-18): exp in fp32 forward, gradient
g * exp(clamp(x, -15, 15)).  The fused field kernels apply both inline; this autograd function is
kept for API compatibility with code that imports `activation.trunc_exp`."""
import torch
from torch.autograd import Function


class _trunc_exp(Function):
    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _trunc_exp.apply
