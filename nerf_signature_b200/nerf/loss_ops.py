"""Host side of the watermark loss-head kernels (csrc/wtmk_loss.cu): the element-wise / reduction glue of the
reference training step (nerf/utils_wtmk_disen.py:592-593, 636-644) as three launches forward and two backward.

    pred, content = split_clamp(image_all, n_block_pixels)       # clamp(image[:n_block], 0, 1), image[n_block:]
    loss, lossi, lossw = wtmk_loss(content, gt, logits, message, lambda_w, lambda_i)

Both are ordinary autograd Functions; the plain PyTorch expressions they replace stay available in the harness
(`Scene(fused_losses=False)`) and are what the parity tests compare against.
"""
import torch
from torch.autograd import Function

from .. import _lib

_P = _lib.ptr


class _split_clamp(Function):
    @staticmethod
    def forward(ctx, image, n_block):
        image = image.contiguous().float()
        flat = image.view(-1, 3)
        n_total = flat.shape[0]
        pred = torch.empty(n_block, 3, dtype=torch.float32, device=image.device)
        content = torch.empty(n_total - n_block, 3, dtype=torch.float32, device=image.device)
        _lib.call("nsig_split_clamp_forward", _P(flat), 3 * n_block, 3 * n_total, _P(pred), _P(content))
        ctx.save_for_backward(flat)
        ctx.n_block = n_block
        ctx.shape = image.shape
        ctx.set_materialize_grads(False)
        return pred, content

    @staticmethod
    def backward(ctx, g_pred, g_content):
        (flat,) = ctx.saved_tensors
        n_total = flat.shape[0]
        g = torch.empty_like(flat)
        _lib.call("nsig_split_clamp_backward", _P(flat), _P(g_pred.contiguous() if g_pred is not None else None),
                  _P(g_content.contiguous() if g_content is not None else None), 3 * ctx.n_block, 3 * n_total, _P(g))
        return g.view(ctx.shape), None


def split_clamp(image, n_block):
    """image [..., 3] holding n_block watermark-block pixels followed by the content pixels ->
    (clamp(block pixels, 0, 1) [n_block, 3], content pixels [n - n_block, 3])."""
    return _split_clamp.apply(image, int(n_block))


class _wtmk_loss(Function):
    @staticmethod
    def forward(ctx, image, gt, logits, message, lambda_w, lambda_i, temp):
        image = image.contiguous().float()
        gt = gt.contiguous().float()
        logits = logits.contiguous().float()
        message = message.contiguous().float()
        if image.numel() != gt.numel() or logits.numel() != message.numel():
            raise ValueError("wtmk_loss: image/gt and logits/message must have matching sizes")
        dev = image.device
        n, md = image.numel(), logits.numel()
        out = torch.empty(3, dtype=torch.float32, device=dev)
        g_image = torch.empty_like(image)
        g_logits = torch.empty_like(logits)
        _lib.call("nsig_wtmk_loss_forward", _P(image), _P(gt), n, _P(logits), _P(message), md, float(lambda_w),
                  float(lambda_i), float(temp), _P(out), _P(g_image), _P(g_logits))
        ctx.save_for_backward(g_image, g_logits)
        loss, lossi, lossw = out[0], out[1], out[2]
        ctx.mark_non_differentiable(lossi, lossw)
        return loss, lossi, lossw

    @staticmethod
    def backward(ctx, grad_loss, _gi, _gw):
        g_image, g_logits = ctx.saved_tensors
        d_image = torch.empty_like(g_image)
        d_logits = torch.empty_like(g_logits)
        grad_loss = grad_loss.contiguous().float().reshape(1)
        _lib.call("nsig_wtmk_loss_backward", _P(g_image), _P(g_logits), g_image.numel(), g_logits.numel(), _P(grad_loss),
                  _P(d_image), _P(d_logits))
        return d_image, None, d_logits, None, None, None, None


def wtmk_loss(image, gt, logits, message, lambda_w, lambda_i, temp=10.0):
    """(loss, lossi, lossw) = (lambda_w*lossw + lambda_i*lossi, mean((image-gt)^2),
    mean BCE-with-logits(logits*temp, message)); differentiable in image and logits through `loss`."""
    return _wtmk_loss.apply(image, gt, logits, message, lambda_w, lambda_i, temp)
