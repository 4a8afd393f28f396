"""Watermark field network with the reference's interface (nerf/network_wtmk_tcnn.py:15-194).

Same constructor arguments, sub-module names and state-dict keys (`encoder.embeddings.{i}.weight`,
`msg_encoder.embeddings.{i}.weight`, `sigma_net.params`, `color_net.params`, `msg_decoder.*` and the
renderer buffers), same freezing policy (base encoder and both MLPs frozen; message tables and
decoder trainable; `finetune_decoder` also freezes the message tables), same
forward / density / color / get_params.  tiny-cuda-nn is replaced by the fused sm_100a kernels in
csrc/field.cu; the per-bit message encoder runs in its pre-summed form (hash_encoding_wtmk_bit.py).
"""
import os

import numpy as np
import torch

from .renderer_wtmk import NeRFRenderer
from .field_ops import FusedMLP, SHEncoding, FieldConfig, field_forward, field_density, color_forward
from .hidden_models import get_hidden_decoder_multi_views, normalize_img
from ..hash_encoding import HashEmbedder
from ..hash_encoding_wtmk_bit import HashEmbedder as HashEmbedder_msg, message_bits
from .. import hash_encoding_wtmk_bit as _hmsg
from .. import _lib


class NeRFNetwork(NeRFRenderer):
    def __init__(self,
                 num_layers=2,
                 hidden_dim=64,
                 geo_feat_dim=15,
                 num_layers_color=3,
                 hidden_dim_color=64,
                 bound=1,
                 message_dim=16,
                 n_views=1,
                 finetune_decoder=False,
                 **kwargs
                 ):
        super().__init__(bound, **kwargs)
        if (num_layers, hidden_dim, geo_feat_dim, num_layers_color, hidden_dim_color) != (2, 64, 15, 3, 64):
            raise NotImplementedError("the fused kernels implement the reference's only configuration: "
                                      "sigma 32-64-16, colour 32-64-64-16 (main_nerf_wtmk.py never overrides it)")

        self.finetune_decoder = finetune_decoder
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.geo_feat_dim = geo_feat_dim
        self.message_dim = message_dim

        self.per_level_scale = np.exp2(np.log2(2048 * bound / 16) / (16 - 1))

        self.encoder = HashEmbedder(bounding_box=(0, 1), n_levels=16, n_features_per_level=2,
                                    log2_hashmap_size=19, base_resolution=16, finest_resolution=2048)
        self.msg_encoder = HashEmbedder_msg(bounding_box=(0, 1), n_levels=message_dim * 2, n_features_per_level=2,
                                            log2_hashmap_size=19, base_resolution=2048, finest_resolution=2048,
                                            message_dim=message_dim)

        self.msg_decoder = get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=8,
                                                          input_ch=n_views * 3, channels=64)
        self.normalization = normalize_img

        self.sigma_net = FusedMLP(n_input_dims=32, n_output_dims=1 + self.geo_feat_dim, n_neurons=hidden_dim,
                                  n_hidden_layers=num_layers - 1, seed=1337)

        self.num_layers_color = num_layers_color
        self.hidden_dim_color = hidden_dim_color
        self.encoder_dir = SHEncoding(n_input_dims=3, degree=4)
        self.in_dim_color = self.encoder_dir.n_output_dims + self.geo_feat_dim
        self.color_net = FusedMLP(n_input_dims=self.in_dim_color, n_output_dims=3, n_neurons=hidden_dim_color,
                                  n_hidden_layers=num_layers_color - 1, seed=1338)

        frozen = [*self.encoder.parameters(), *self.color_net.parameters(), *self.sigma_net.parameters()]
        if finetune_decoder:
            frozen += [*self.msg_encoder.parameters()]
        for param in frozen:
            param.requires_grad = False

        self._S_cache = None

    # ---- helpers -----------------------------------------------------------------------------
    def _cfg(self, density_scale=1.0):
        shadow = self.encoder.half_tables() if self.half2_tables else None
        sink = self.msg_encoder.grad_sink if (torch.is_grad_enabled() and _hmsg.grad_reducer is None
                                              and os.environ.get("NSIG_NO_DIRECT_SINK") != "1") else None
        return FieldConfig(self.bound, self.encoder.resolutions, self.encoder.log2_hashmap_size,
                           self.msg_encoder.resolution, density_scale, shadow, sink)

    def _summed_table(self, message):
        """S for this message; cached across the two render passes of a training step (keyed on the
        message object/version and the versions of the message tables)."""
        if message is None:
            return None
        tabs = self.msg_encoder.tables()
        key = (id(message), getattr(message, "_version", None), torch.is_grad_enabled(),
               tuple(t._version for t in tabs), tabs[0].data_ptr())
        if self._S_cache is not None and self._S_cache[0] == key:
            return self._S_cache[1]
        sink_mode = self.msg_encoder.shard is not None or (self.msg_encoder.grad_sink is not None and torch.is_grad_enabled())
        S = self.msg_encoder.summed_table(message, None if sink_mode else message_bits(message))
        self._S_cache = (key, S, message)  # keep `message` alive so id() stays unique
        return S

    def prefetch_summed_table(self, message):
        """renderer hook, called before the march: start the table sum (message_dim x 4 MiB streamed, independent of the
        rays) on a side stream so it overlaps the march kernels; `field` joins the stream before the first use of S."""
        if message is None or not message.is_cuda:
            return
        side = _lib.side_stream(message.device)
        if side is None:
            return
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            S = self._summed_table(message)
        S.record_stream(cur)
        self._S_pending = side

    def _join_prefetch(self):
        side = getattr(self, "_S_pending", None)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            self._S_pending = None

    def field(self, xyzs, dirs, message, count=None):
        """renderer hook: (density_scale * sigma, rgb), fused."""
        self._join_prefetch()
        return field_forward(xyzs, dirs, self._summed_table(message), count, self._cfg(self.density_scale),
                             self.sigma_net, self.color_net, self.encoder.tables())

    def field_args(self, message):
        """renderer hook for the fused frame renderer: (S, cfg, sigma_mlp, color_mlp, base tables)."""
        S = self._summed_table(message)
        return (S.detach() if S is not None else None, self._cfg(self.density_scale), self.sigma_net, self.color_net,
                self.encoder.tables())

    # ---- reference API -------------------------------------------------------------------------
    def forward(self, x, d, message=None):
        # x: [N, 3] in [-bound, bound]; d: [N, 3] normalised.  Returns sigma [N] fp32, color [N,3] fp32.
        return field_forward(x, d, self._summed_table(message), None, self._cfg(1.0), self.sigma_net,
                             self.color_net, self.encoder.tables())

    def density(self, x, message=None):
        S = self._summed_table(message)
        if S is not None:
            S = S.detach()
        sigma, geo_feat = field_density(x, S, self._cfg(1.0), self.sigma_net, self.encoder.tables())
        return {
            'sigma': sigma,
            'geo_feat': geo_feat,
        }

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
        # x is unused by the colour branch (as in the reference); mask selects the rows to evaluate.
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], 3, dtype=torch.float32, device=d.device)
            if not mask.any():
                return rgbs
            d = d[mask]
            geo_feat = geo_feat[mask]
        h = color_forward(d, geo_feat, self.color_net)
        if mask is not None:
            rgbs[mask] = h.to(rgbs.dtype)
        else:
            rgbs = h
        return rgbs

    def decode_blocks(self, image):
        """msg_decoder(normalization(image.permute(0, 3, 1, 2))) for rendered blocks `image` [B,H,W,3] in [0,1], with
        float16-autocast arithmetic (utils_wtmk_disen.py:592-595), through the fused decoder kernels."""
        from .decoder_ops import decode
        return decode(self.msg_decoder, image, getattr(self, "_dec_prepared", None),
                      defer_weight_grads=getattr(self, "defer_decoder_weight_grads", False) and torch.is_grad_enabled())

    def prepare_decoder_weights(self, enable=True):
        """Keep the decoder's fp16 weight copies in a persistent buffer (decoder_ops.PreparedWeights) instead of converting
        them inside every decode_blocks call.  Whoever changes the conv weights afterwards must call
        `refresh_decoder_weights()`; optim.WatermarkAdam does so after each of its steps."""
        from .decoder_ops import PreparedWeights
        self._dec_prepared = PreparedWeights(self.msg_decoder) if enable else None
        if enable and not getattr(self, "_dec_prepared_hook", False):   # state-dict loads change the weights too
            self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.refresh_decoder_weights())
            self._dec_prepared_hook = True
        return self._dec_prepared

    def refresh_decoder_weights(self):
        if getattr(self, "_dec_prepared", None) is not None:
            self._dec_prepared.refresh()

    def get_params(self, lr):
        if self.finetune_decoder:
            params = [
                {'params': self.msg_decoder.parameters(), 'lr': lr},
            ]
        else:
            params = [
                {'params': self.msg_encoder.parameters(), 'lr': lr},
                {'params': self.msg_decoder.parameters(), 'lr': lr},
            ]
        if self.bg_radius > 0:
            params.append({'params': self.encoder_bg.parameters(), 'lr': lr})
            params.append({'params': self.bg_net.parameters(), 'lr': lr})
        return params
