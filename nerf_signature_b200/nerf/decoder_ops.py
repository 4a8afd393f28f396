"""Host side of the fused HiDDeN decoder kernels (csrc/decoder.cu).

`decode(decoder, image)` evaluates `decoder(normalize_img(image.permute(0, 3, 1, 2)))` under float16 autocast
semantics for an [B, H, W, 3] block image (hidden_models.py:104-137; call sites utils_wtmk_disen.py:592-595), and
its backward produces the gradient of the image and of every decoder parameter.  The module itself
(`HiddenDecoder_multi_views`) is unchanged - same parameters, same state dict - and remains the path for shapes
the kernels do not cover.
"""
import torch
from torch.autograd import Function

from .. import _lib

_P = _lib.ptr


def decoder_params(decoder):
    """Parameters in the order nsig_decoder_forward expects, or None when the architecture is not the
    reference's ConvBNRelu(3x3) x (num_blocks+1) -> AdaptiveAvgPool2d -> Linear with 64 channels."""
    blocks = [m for m in decoder.layers if hasattr(m, "layers")]
    if len(blocks) < 2 or len(blocks) != len(decoder.layers) - 1:
        return None
    ps = []
    nb = decoder.num_bits * decoder.redundancy
    for i, blk in enumerate(blocks):
        conv, bn = blk.layers[0], blk.layers[1]
        cin = 3 if i == 0 else 64
        cout = nb if i == len(blocks) - 1 else 64
        if tuple(conv.weight.shape) != (cout, cin, 3, 3) or conv.bias is None or bn.weight is None or abs(bn.eps - 1e-3) > 1e-12:
            return None
        ps += [conv.weight, conv.bias, bn.weight, bn.bias]
    if nb > 8 or tuple(decoder.linear.weight.shape) != (nb, nb):
        return None
    ps += [decoder.linear.weight, decoder.linear.bias]
    return ps


_deferred_keep = []   # workspaces / gradient scratch of backward calls whose weight-gradient tail is still pending


def finish_backward():
    """Join the deferred tail of the last fused-decoder backward (decode(..., defer_weight_grads=True)) into the current stream:
    after this call the decoder's conv weight / bias gradients are complete in stream order.  No-op when nothing is pending."""
    if _deferred_keep:
        _lib.call("nsig_decoder_finish_backward")
        for g, tmp in [f for item in _deferred_keep for f in item[1]]:
            g.add_(tmp)
        _deferred_keep.clear()


class _fused_decode(Function):
    @staticmethod
    def forward(ctx, image, meta, prepared, *params):
        num_blocks, num_bits, redundancy, ctx.defer = meta
        image = image.contiguous().float()
        B, H, W, _ = image.shape
        dev = image.device
        ws = torch.empty(_lib.load().nsig_decoder_workspace_bytes(B, H, W, num_blocks), dtype=torch.uint8, device=dev)
        logits = torch.empty(B, num_bits, dtype=torch.float32, device=dev)
        flat = [p.detach().contiguous() for p in params]
        _lib.call("nsig_decoder_forward", _P(image), B, H, W, num_blocks, num_bits, redundancy, _lib.pointer_array(flat),
                  _P(ws), _P(logits), _P(prepared))
        ctx.prepared = prepared
        ctx.meta = (B, H, W, num_blocks, num_bits, redundancy)
        ctx.ws = ws
        ctx.params = params
        ctx.need_image = image.requires_grad or ctx.needs_input_grad[0]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        """RESTRICTION (deliberate): parameter gradients are ACCUMULATED INTO `p.grad` here and `None` is returned for the
        parameters, instead of handing tensors to autograd's accumulator - that is what lets every decoder gradient land
        directly in the flat all-reduce bucket with zero extra kernels.  Consequences: `torch.autograd.grad`, per-parameter
        hooks and DistributedDataParallel's reducer do not see these gradients (use the plain module, harness
        fused_decoder=False, for those), and the saved workspace is released after the first backward."""
        if ctx.ws is None:
            raise RuntimeError("fused HiDDeN decoder: backward called twice on the same graph (retain_graph is not "
                               "supported: the activation workspace is released after the first backward)")
        B, H, W, num_blocks, num_bits, redundancy = ctx.meta
        dev = dlogits.device
        params = ctx.params
        # parameter gradients are accumulated straight into .grad (the data-parallel harness points them at its
        # flat all-reduce bucket): no per-parameter zeros / add kernels.  A missing .grad is created as zeros.
        grads, fixups = [], []
        for p in params:
            if not p.requires_grad:
                grads.append(torch.zeros(p.shape, dtype=torch.float32, device=dev))   # scratch sink, discarded
                continue
            if p.grad is None:
                p.grad = torch.zeros(p.shape, dtype=torch.float32, device=dev)
            if p.grad.is_contiguous():
                grads.append(p.grad)
            else:  # e.g. a channels_last weight whose .grad kept that layout: go through a dense scratch
                tmp = torch.zeros(p.shape, dtype=torch.float32, device=dev)
                grads.append(tmp)
                fixups.append((p.grad, tmp))
        dimage = torch.empty(B, H, W, 3, dtype=torch.float32, device=dev) if ctx.need_image else None
        if ctx.defer:
            # the call returns with the weight gradients still running on the library's side streams: the workspace (their
            # partials) and the gradient buffers must outlive it - parked until finish_backward()
            finish_backward()
            _lib.load().nsig_decoder_defer_weight_grads(1)
        try:
            _lib.call("nsig_decoder_backward", _P(dlogits.contiguous().float()), B, H, W, num_blocks, num_bits, redundancy,
                      _lib.pointer_array([p.detach().contiguous() for p in params]), _lib.pointer_array(grads), _P(ctx.ws),
                      _P(dimage), _P(ctx.prepared))
        finally:
            if ctx.defer:
                _lib.load().nsig_decoder_defer_weight_grads(0)
        if ctx.defer:
            _deferred_keep.append((ctx.ws, fixups, grads))
        else:
            for g, tmp in fixups:
                g.add_(tmp)
        ctx.ws = None
        return (dimage, None, None) + (None,) * len(params)


class PreparedWeights:
    """fp16 copies of a decoder's conv weights in the two layouts the kernels consume (nsig_decoder_prepare_weights), kept
    in one persistent buffer so the conversion runs once per optimizer update instead of inside every forward.  The owner
    must call `refresh()` after ANY change of the decoder's conv weights (optim.WatermarkAdam does after its flat-bucket
    step; checkpoint loads and manual edits must do it themselves) - nothing here can detect a write through a raw pointer."""

    def __init__(self, decoder):
        ps = decoder_params(decoder)
        if ps is None:
            raise ValueError("decoder does not have the architecture the fused kernels cover")
        self.decoder = decoder
        self.num_blocks = len(ps) // 4 - 1
        self.buffer = torch.empty(_lib.load().nsig_decoder_weights_bytes(self.num_blocks), dtype=torch.uint8, device=ps[0].device)
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        ps = decoder_params(self.decoder)
        _lib.call("nsig_decoder_prepare_weights", _lib.pointer_array([p.detach().contiguous() for p in ps]), self.num_blocks,
                  self.decoder.num_bits, self.decoder.redundancy, _P(self.buffer))


def decode(decoder, image, prepared=None, defer_weight_grads=False):
    """logits [B, num_bits] of `image` [B,H,W,3] (fp32, in [0,1]); fused kernels when the decoder has the
    reference architecture, the plain module (under autocast) otherwise.  prepared: optional PreparedWeights of `decoder`.
    defer_weight_grads: the backward returns once the image gradient is enqueued; the decoder's weight gradients are complete
    only after finish_backward() (a training loop calls it once the rest of its backward is issued, so the renderer's
    backward kernels run next to the decoder's weight gradients)."""
    ps = decoder_params(decoder) if image.is_cuda else None
    if ps is None:
        from .hidden_models import normalize_img
        with torch.autocast("cuda", dtype=torch.float16, enabled=image.is_cuda):
            return decoder(normalize_img(image.permute(0, 3, 1, 2)))
    num_blocks = len(ps) // 4 - 1
    return _fused_decode.apply(image, (num_blocks, decoder.num_bits, decoder.redundancy, bool(defer_weight_grads)),
                               prepared.buffer if prepared is not None else None, *ps)
