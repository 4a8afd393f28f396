"""CPU tests (-m "not gpu"): the oracle of the fused field kernel (oracle/hash_oracle.c + oracle/field_oracle.py, the chain
tests/test_field_gpu.py and __graft_entry__.smoke() compare the kernels with) against the REFERENCE'S OWN forward / density /
color method bodies (network_wtmk_tcnn.py:97-176), run unmodified on CPU around stand-ins for the three tiny-cuda-nn modules
(tests/golden/make_golden_field.py).  Pins the wiring at the reference's call sites; tcnn's internal arithmetic stays unpinned."""
import os

import numpy as np
import pytest
import torch

import make_golden_field as mg
from nerf_signature_b200.hash_encoding import level_resolutions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "field_golden.npz"))


def _oracle(oracle_cpu, name, with_message=True):
    from oracle import field_oracle as fo
    bound, md, x, d, msg, base, msgt, sp, cp, mask = mg.case_inputs(name)
    xn = ((x + np.float32(bound)) * np.float32(0.5 / bound)).astype(np.float32)
    base_r, finest_r = torch.tensor(16), torch.tensor(2048)                     # hash_encoding.py:56-60
    res = np.asarray(level_resolutions(base_r, torch.exp((torch.log(finest_r) - torch.log(base_r)) / 15), 16), np.float32)
    feat = oracle_cpu.hash_encode_forward(xn, base, res, mg.LOG2_T)
    if md and with_message:
        feat[:, 30:32] += oracle_cpu.msg_encode_forward(xn, msgt, msg, 2048.0, mg.LOG2_T)
    sig, rgb, logit, geo = fo.mlp_forward(torch.from_numpy(feat), torch.from_numpy(d), sp, cp)
    return sig.numpy(), rgb.numpy(), geo.numpy(), mask.numpy()


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("case", list(mg.CASES))
def test_field_oracle_matches_reference_forward_wiring(golden, oracle_cpu, case):
    sig, rgb, geo, mask = _oracle(oracle_cpu, case)
    g_sig, g_rgb, g_geo = golden[f"{case}_sigma"], golden[f"{case}_color"], golden[f"{case}_geo_feat"]
    assert np.ptp(g_sig) > 0.5 * g_sig.mean() and np.ptp(g_rgb) > 0.3          # a field that varies: wiring errors would show
    # fp16 storage rounding sits at the same places on both sides: observed difference 0 on the generating machine; the
    # tolerance leaves room for another host's fp32 matmul summation order (an fp16 rounding flip of one activation)
    assert _rel(sig, g_sig) < 1e-4 and _rel(geo, g_geo) < 1e-4 and np.abs(rgb - g_rgb).max() < 1e-4
    # masked colour query (renderer's run(): colour only where the weight matters): zeros outside the mask
    masked = golden[f"{case}_rgb_masked"]
    assert not masked[~mask].any() and np.abs(masked[mask] - rgb[mask]).max() < 1e-4


def test_message_feature_enters_the_last_two_encoder_channels(golden, oracle_cpu):
    """Sanity of the fixture's discriminating power: leaving the message feature out (or the wiring wrong) is far outside
    the tolerance above."""
    case = "md4_bound1"
    sig, rgb, _, _ = _oracle(oracle_cpu, case, with_message=False)
    assert _rel(sig, golden[f"{case}_sigma"]) > 2e-2
