"""Clean (un-watermarked) field network with the reference's interface (nerf/network_hash.py:13-165):
hash encoder + sigma/colour MLPs, no message branch.  forward(x, d) / density(x) / color / get_params.
"""
import numpy as np
import torch

from .renderer import NeRFRenderer
from .field_ops import FusedMLP, SHEncoding, FieldConfig, field_forward, field_density, color_forward
from ..hash_encoding import HashEmbedder


class NeRFNetwork(NeRFRenderer):
    def __init__(self,
                 encoding="HashGrid",
                 encoding_dir="SphericalHarmonics",
                 num_layers=2,
                 hidden_dim=64,
                 geo_feat_dim=15,
                 num_layers_color=3,
                 hidden_dim_color=64,
                 bound=1,
                 **kwargs
                 ):
        super().__init__(bound, **kwargs)
        if (num_layers, hidden_dim, geo_feat_dim, num_layers_color, hidden_dim_color) != (2, 64, 15, 3, 64):
            raise NotImplementedError("the fused kernels implement sigma 32-64-16 / colour 32-64-64-16 only")
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.geo_feat_dim = geo_feat_dim
        self.per_level_scale = np.exp2(np.log2(2048 * bound / 16) / (16 - 1))

        self.encoder = HashEmbedder(bounding_box=(0, 1), n_levels=16, n_features_per_level=2,
                                    log2_hashmap_size=19, base_resolution=16, finest_resolution=2048)
        self.sigma_net = FusedMLP(n_input_dims=32, n_output_dims=1 + self.geo_feat_dim, n_neurons=hidden_dim,
                                  n_hidden_layers=num_layers - 1, seed=1337)
        self.num_layers_color = num_layers_color
        self.hidden_dim_color = hidden_dim_color
        self.encoder_dir = SHEncoding(n_input_dims=3, degree=4)
        self.in_dim_color = self.encoder_dir.n_output_dims + self.geo_feat_dim
        self.color_net = FusedMLP(n_input_dims=self.in_dim_color, n_output_dims=3, n_neurons=hidden_dim_color,
                                  n_hidden_layers=num_layers_color - 1, seed=1338)

    def _cfg(self, density_scale=1.0):
        shadow = self.encoder.half_tables() if self.half2_tables else None
        return FieldConfig(self.bound, self.encoder.resolutions, self.encoder.log2_hashmap_size, 0.0, density_scale,
                           shadow)

    def field(self, xyzs, dirs, message=None, count=None):
        return field_forward(xyzs, dirs, None, count, self._cfg(self.density_scale), self.sigma_net,
                             self.color_net, self.encoder.tables())

    def field_args(self, message=None):
        return (None, self._cfg(self.density_scale), self.sigma_net, self.color_net, self.encoder.tables())

    def forward(self, x, d):
        return field_forward(x, d, None, None, self._cfg(1.0), self.sigma_net, self.color_net,
                             self.encoder.tables())

    def density(self, x, message=None):
        sigma, geo_feat = field_density(x, None, self._cfg(1.0), self.sigma_net, self.encoder.tables())
        return {
            'sigma': sigma,
            'geo_feat': geo_feat,
        }

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
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

    def get_params(self, lr):
        params = [
            {'params': self.encoder.parameters(), 'lr': lr},
            {'params': self.sigma_net.parameters(), 'lr': lr},
            {'params': self.encoder_dir.parameters(), 'lr': lr},
            {'params': self.color_net.parameters(), 'lr': lr},
        ]
        if self.bg_radius > 0:
            params.append({'params': self.encoder_bg.parameters(), 'lr': lr})
            params.append({'params': self.bg_net.parameters(), 'lr': lr})
        return params
