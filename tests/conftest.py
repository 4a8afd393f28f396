import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_hash():
    return np.load(os.path.join(ROOT, "tests", "golden", "hash_golden.npz"))


@pytest.fixture(scope="session")
def oracle_cpu():
    from oracle import cpu
    cpu.build()
    return cpu


def record_parity(name, metrics):
    """Append measured parity errors to gpurun_out/parity_metrics.jsonl (when that directory exists: GPU runs under
    gpurun) so the tolerances written in the tests can be audited against what was actually observed; the summary is
    committed under profiles/."""
    import json
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    with open(os.path.join(out, "parity_metrics.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **metrics}) + "\n")
