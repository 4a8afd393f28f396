"""-m gpu: the opt-in kernel variants that an environment switch selects when the library is first used (so they need their own
process) produce the SAME BITS as the default kernels:
  NSIG_MARCH_FUSED=1   march_rays_train as one kernel with decoupled look-back offsets  vs  count / scan / write
  NSIG_ADAM_TMA=2      message-table Adam through a shared-memory ring of bulk asynchronous copies  vs  the register kernel
Each variant runs tests/_variant_probe.py in a subprocess and prints sha256 digests of its outputs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _probe(what, **env):
    e = dict(os.environ)
    for k in ("NSIG_MARCH_FUSED", "NSIG_ADAM_TMA"):
        e.pop(k, None)
    e.update(env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_variant_probe.py"), what], env=e, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("digest ")]
    assert lines, out.stdout[-2000:]
    return lines


def test_single_launch_march_is_bit_identical_to_three_kernel_march():
    assert _probe("march") == _probe("march", NSIG_MARCH_FUSED="1")


def test_tma_staged_adam_is_bit_identical_to_register_adam():
    base = _probe("adam")
    assert base == _probe("adam", NSIG_ADAM_TMA="2")
    assert base == _probe("adam", NSIG_ADAM_TMA="1")
