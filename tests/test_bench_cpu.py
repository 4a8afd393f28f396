"""CPU tests (-m "not gpu") of bench.py's reference arm (`--impl reference`): the driver launches it like our own arm
(torchrun included) and computes the headline ratio from its JSON line, so the line's contract is checked here where
it can run - the arm is the oracle port of the reference's CPU path and needs no GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *argv):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    env.update(extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *argv], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600)


@pytest.mark.timeout(900)
def test_reference_arm_prints_one_contract_line():
    out = _run({}, "--steps", "1", "--warmup", "1", "--cpu-rays", "4")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["metric"].split(" (")[0] in base["metric"]                       # "train rays/s" is BASELINE's metric
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["vs_baseline"] is None and not base["published"]                 # nothing published for this metric
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and "rays per step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "blender_wtmk" and d["config"]["message_dim"] == 32
    # rays/s follows from the bounded sample it names: (content + block rays) / step time
    rays = 4 + 32 * 1 * 1
    assert abs(d["value"] - rays / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    c0 = d["configs0_cpu_render"]                                             # BASELINE configs[0] exactly as stated
    assert c0["rays"] == 4096 and c0["samples_per_ray"] == 512 and c0["ms_per_render"] > 0


def test_reference_arm_is_rank_0_only():
    """Under torchrun the other ranks exit 0 without work and without output."""
    out = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--steps", "1", "--warmup", "1")
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
