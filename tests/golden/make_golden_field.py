"""Golden vectors for the WIRING of NeRFNetwork.forward / density / color FROM THE REFERENCE'S OWN METHOD BODIES.

Run in the build container (needs /root/reference; not needed on the GPU box):
    python tests/golden/make_golden_field.py

nerf/network_wtmk_tcnn.py cannot be imported: it needs tiny-cuda-nn, an unvendored, unpinned dependency that is not
installed anywhere offline.  What CAN be run unmodified is everything around the three tcnn modules:
`forward`, `density` and `color` are cut out of the source text with `ast`, compiled as they are and called on a stub
`self` whose
  * `encoder` / `msg_encoder` are the reference's own HashEmbedder modules (hash_encoding.py, hash_encoding_wtmk_bit.py),
  * `trunc_exp` is the reference's activation.py,
  * `encoder_dir` stands in for tcnn's SphericalHarmonics(degree 4): the reference's own SHEncoder (hash_encoding.py:114-195)
    on 2 d - 1 (tcnn maps its [0, 1] input back to [-1, 1]),
  * `sigma_net` / `color_net` stand in for tcnn FullyFusedMLP: bias-free matrices taken from a flat parameter vector in
    layer order, row-major [out, in], input padded to a multiple of 16 and output to 16 (tcnn's published layout), ReLU
    hidden layers, no output activation, n_output_dims columns returned - evaluated with the storage rounding the oracle
    states (fp16 weights and activations, fp32 accumulation).
The fixture therefore pins what oracle/field_oracle.py + the hash oracles claim about the reference's call sites
(network_wtmk_tcnn.py:97-176): coordinate mapping, the message feature added to the LAST TWO encoder channels, column 0 ->
trunc_exp, geo_feat = columns 1..15, (d + 1) / 2, concatenation order [SH | geo_feat], sigmoid, the masked colour query.
The arithmetic INSIDE tcnn's kernels (fp16 accumulation) stays unpinned, as DESIGN.md 5 says.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden_grid import REF  # noqa: E402
from make_golden_hash import load_ref_module, make_tables  # noqa: E402

LOG2_T = 12
CASES = {"md4_bound1": (1.0, 4, 300, 41), "md8_bound2": (2.0, 8, 200, 42), "clean_bound1": (1.0, 0, 128, 43)}


def cut_methods(path, names):
    """Like make_golden_grid.cut_functions, for bodies that hold comment lines starting in column 0 (textwrap.dedent
    gives up on those): the `def` line's indentation is removed from every line that has it."""
    import ast
    src = open(path).read()
    lines = src.split("\n")
    out = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            body = lines[node.lineno - 1:node.end_lineno]
            pad = body[0][:len(body[0]) - len(body[0].lstrip())]
            out[node.name] = "\n".join(l[len(pad):] if l.startswith(pad) else l for l in body)
    return out


def q(t):
    return t.half().float()


class FusedMLPStandIn(torch.nn.Module):
    """tcnn.Network(otype=FullyFusedMLP, ReLU, no output activation) as far as its interface goes."""

    def __init__(self, params, n_in, n_out, n_hidden_layers, width=64):
        super().__init__()
        pad_in = -(-n_in // 16) * 16
        shapes = [(width, pad_in)] + [(width, width)] * (n_hidden_layers - 1) + [(16, width)]
        self.n_in, self.pad_in, self.n_out = n_in, pad_in, n_out
        self.mats, o = [], 0
        for r, c in shapes:
            self.mats.append(q(params[o:o + r * c].view(r, c)))
            o += r * c
        assert o == params.numel()

    def forward(self, x):
        assert x.shape[-1] == self.n_in
        h = q(torch.cat([x, torch.zeros(x.shape[0], self.pad_in - self.n_in)], -1))
        for m in self.mats[:-1]:
            h = q(torch.relu(h @ m.t()))
        return (h @ self.mats[-1].t())[:, :self.n_out]


def case_inputs(name):
    bound, md, n, seed = CASES[name]
    rs = np.random.RandomState(seed)
    x = rs.uniform(-bound, bound, size=(n, 3)).astype(np.float32)
    x[0], x[1] = -bound, bound                                   # box corners
    d = rs.standard_normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    msg = rs.randint(0, 2, size=md).astype(np.float32) if md else None
    base = [t * np.float32(2e4) for t in make_tables(seed, 16, LOG2_T)]          # features O(1): a visible field
    msgt = [t * np.float32(2e4) for t in make_tables(seed + 1, 2 * md, LOG2_T)] if md else []
    g = torch.Generator().manual_seed(seed)
    sigma_params = (torch.rand(3072, generator=g) * 2 - 1) * 0.5
    color_params = (torch.rand(7168, generator=g) * 2 - 1) * 0.5
    mask = torch.from_numpy(rs.uniform(size=n) < 0.6)
    return bound, md, x, d, msg, base, msgt, sigma_params, color_params, mask


def main():
    torch.set_num_threads(4)
    he = load_ref_module("hash_encoding.py", "ref_hash_encoding")
    hm = load_ref_module("hash_encoding_wtmk_bit.py", "ref_hash_encoding_wtmk_bit")
    act = load_ref_module("activation.py", "ref_activation")
    fns = cut_methods(os.path.join(REF, "nerf", "network_wtmk_tcnn.py"), {"forward", "density", "color"})
    env = {"torch": torch, "trunc_exp": act.trunc_exp}
    for f in fns.values():
        exec(compile(f, "ref:network_wtmk_tcnn", "exec"), env)
    sh = he.SHEncoder(input_dim=3, degree=4)
    out = {}
    for name in CASES:
        bound, md, x, d, msg, base, msgt, sp, cp, mask = case_inputs(name)
        self = types.SimpleNamespace(bound=bound)
        self.encoder = he.HashEmbedder(bounding_box=(0, 1), n_levels=16, n_features_per_level=2, log2_hashmap_size=LOG2_T,
                                       base_resolution=16, finest_resolution=2048)
        with torch.no_grad():
            for e, t in zip(self.encoder.embeddings, base):
                e.weight.copy_(torch.from_numpy(t))
        if md:
            self.msg_encoder = hm.HashEmbedder(bounding_box=(0, 1), n_levels=2 * md, n_features_per_level=2,
                                               log2_hashmap_size=LOG2_T, base_resolution=2048, finest_resolution=2048,
                                               message_dim=md)
            with torch.no_grad():
                for e, t in zip(self.msg_encoder.embeddings, msgt):
                    e.weight.copy_(torch.from_numpy(t))
        self.encoder_dir = lambda u: sh(u * 2 - 1)
        self.sigma_net = FusedMLPStandIn(sp, 32, 16, 1)
        self.color_net = FusedMLPStandIn(cp, 16 + 15, 3, 2)
        xt, dt = torch.from_numpy(x), torch.from_numpy(d)
        mt = torch.from_numpy(msg) if md else None
        with torch.no_grad():
            sigma, color = env["forward"](self, xt, dt, mt)
            dens = env["density"](self, xt, mt)
            rgb_masked = env["color"](self, xt, dt, mask=mask, **dens)
            rgb_all = env["color"](self, xt, dt, geo_feat=dens["geo_feat"])
        assert torch.equal(dens["sigma"], sigma) and torch.equal(rgb_all, color)
        out[f"{name}_sigma"], out[f"{name}_color"] = sigma.numpy(), color.numpy()
        out[f"{name}_geo_feat"], out[f"{name}_rgb_masked"] = dens["geo_feat"].numpy(), rgb_masked.numpy()
        print(name, "sigma", float(sigma.min()), float(sigma.max()), "rgb", float(color.min()), float(color.max()))
    np.savez_compressed(os.path.join(HERE, "field_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "field_golden.npz"), os.path.getsize(os.path.join(HERE, "field_golden.npz")))


if __name__ == "__main__":
    main()
