"""Generate golden vectors for the hash encoders FROM THE REFERENCE MODULES THEMSELVES.

Run in the build container (needs /root/reference; not needed on the GPU box):
    python tests/golden/make_golden_hash.py

The reference modules create BOX_OFFSETS on device='cuda' at import (hash_encoding.py:8-9,
hash_encoding_wtmk_bit.py:9-10), which fails on a GPU-less host, so their source text is
exec'd in memory with that one literal replaced by 'cpu'.  Nothing is copied into this repo.

Tables are filled from numpy's legacy RandomState (bit-stable across versions) so the big
2^19 fixtures only need to store the seed; tests regenerate the tables with `make_tables`.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NSIG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref_module(fname, name):
    src = open(os.path.join(REF, fname)).read().replace("device='cuda'", "device='cpu'")
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF, fname)
    sys.path.insert(0, REF)  # hash_encoding_wtmk_bit imports msgencoder
    try:
        exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    finally:
        sys.path.pop(0)
    return mod


def make_tables(seed, n_tables, log2_T):
    """U(-1e-4, 1e-4) like hash_encoding.py:65-66, from a reproducible generator."""
    rs = np.random.RandomState(seed)
    return [rs.uniform(-1e-4, 1e-4, size=(1 << log2_T, 2)).astype(np.float32) for _ in range(n_tables)]


def make_points(seed, B):
    rs = np.random.RandomState(seed)
    x = rs.uniform(0.0, 1.0, size=(B, 3)).astype(np.float32)
    # edge cases: box corners, exact cell boundaries, 1.0 (idx == resolution)
    x[0] = (0.0, 0.0, 0.0)
    x[1] = (1.0, 1.0, 1.0)
    x[2] = (0.5, 0.25, 0.125)
    x[3] = (1.0, 0.0, 0.5)
    x[4] = np.float32(1.0) - np.float32(2.0 ** -24)
    x[5] = (3.0 / 2048, 17.0 / 2048, 2047.0 / 2048)
    return x


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    he = load_ref_module("hash_encoding.py", "ref_hash_encoding")
    hm = load_ref_module("hash_encoding_wtmk_bit.py", "ref_hash_encoding_wtmk_bit")

    out = {}
    # ---- base encoder, as instantiated by nerf/network_wtmk_tcnn.py:40-41 -------------------
    for tag, log2_T, B, seed in (("small", 10, 192, 11), ("full", 19, 256, 12)):
        enc = he.HashEmbedder(bounding_box=(0, 1), n_levels=16, n_features_per_level=2,
                              log2_hashmap_size=log2_T, base_resolution=16, finest_resolution=2048)
        tabs = make_tables(seed, 16, log2_T)
        with torch.no_grad():
            for i in range(16):
                enc.embeddings[i].weight.copy_(torch.from_numpy(tabs[i]))
        x = make_points(seed + 100, B)
        xt = torch.from_numpy(x)
        res = [float(torch.floor(enc.base_resolution * enc.b ** i)) for i in range(16)]
        slots = []
        for i in range(16):
            r = torch.floor(enc.base_resolution * enc.b ** i)
            _, _, hv, _ = he.get_voxel_vertices(xt, enc.bounding_box, r, log2_T)
            slots.append(hv.numpy().astype(np.int32))
        feats = enc(xt)
        out[f"base_{tag}_seed"] = np.int64(seed)
        out[f"base_{tag}_log2T"] = np.int64(log2_T)
        out[f"base_{tag}_x"] = x
        out[f"base_{tag}_res"] = np.asarray(res, np.float32)
        out[f"base_{tag}_slots"] = np.stack(slots, 1)          # [B,16,8]
        out[f"base_{tag}_feat"] = feats.detach().numpy()
        if tag == "small":  # autograd of the tables (dense grads are small at 2^10)
            rs = np.random.RandomState(seed + 7)
            g = rs.standard_normal(size=feats.shape).astype(np.float32)
            feats.backward(torch.from_numpy(g))
            out["base_small_gout"] = g
            out["base_small_gtab"] = np.stack([e.weight.grad.numpy() for e in enc.embeddings])

    # ---- message-bit encoder, nerf/network_wtmk_tcnn.py:43-44 -------------------------------
    for tag, log2_T, md, B, seed in (("small", 10, 4, 192, 21), ("md32", 19, 32, 128, 22), ("md48", 19, 48, 64, 23)):
        enc = hm.HashEmbedder(bounding_box=(0, 1), n_levels=md * 2, n_features_per_level=2,
                              log2_hashmap_size=log2_T, base_resolution=2048, finest_resolution=2048,
                              message_dim=md)
        tabs = make_tables(seed, 2 * md, log2_T)
        with torch.no_grad():
            for i in range(2 * md):
                enc.embeddings[i].weight.copy_(torch.from_numpy(tabs[i]))
        x = make_points(seed + 100, B)
        msg = np.random.RandomState(seed + 1).randint(0, 2, size=(md,)).astype(np.float32)
        res = [float(torch.floor(enc.base_resolution * enc.b ** i)) for i in range(md)]
        assert all(r == 2048.0 for r in res), res
        feats = enc(torch.from_numpy(x), torch.from_numpy(msg))
        out[f"msg_{tag}_seed"] = np.int64(seed)
        out[f"msg_{tag}_log2T"] = np.int64(log2_T)
        out[f"msg_{tag}_md"] = np.int64(md)
        out[f"msg_{tag}_x"] = x
        out[f"msg_{tag}_message"] = msg
        out[f"msg_{tag}_feat"] = feats.detach().numpy()
        if tag == "small":
            rs = np.random.RandomState(seed + 7)
            g = rs.standard_normal(size=feats.shape).astype(np.float32)
            feats.backward(torch.from_numpy(g))
            out["msg_small_gout"] = g
            out["msg_small_gtab"] = np.stack([
                (e.weight.grad.numpy() if e.weight.grad is not None else np.zeros((1 << log2_T, 2), np.float32))
                for e in enc.embeddings])

    # ---- SH degree 4 (hash_encoding.py:114-195) — the restatement of tcnn's SphericalHarmonics --
    sh = he.SHEncoder(input_dim=3, degree=4)
    d = np.random.RandomState(31).standard_normal(size=(64, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out["sh_dirs"] = d
    out["sh_out"] = sh(torch.from_numpy(d)).numpy()

    path = os.path.join(HERE, "hash_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
