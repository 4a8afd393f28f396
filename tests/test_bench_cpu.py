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


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_roofline_arithmetic_follows_from_its_own_inputs():
    """`frac` must be sum(algorithmic bytes) / sum(launch time) over the SAME replays, against the peak
    of MEASURED_PEAKS.json.  rooflines() is a pure function of time_scene's accounting: checked on synthetic numbers."""
    b = _bench_module()
    assert b.ALG_FWD == 24 + 1024 + 64 + 16 and b.ALG_BWD == 32 + 16 + 16 + 128 and b.FLOP_BWD == 20480   # SURVEY 8d, DESIGN 4
    n, S, fwd_ms, bwd_ms, step_ms = 8, 8 * 1_050_000.0, 8 * 0.25, 8 * 0.1, 0.9
    res = {"kernel_ms": {"nsig_field_forward": {"n": n, "ms": fwd_ms},
                         "nsig_field_backward_masks": {"n": n, "ms": bwd_ms}, "nsig_field_backward": {"n": 0, "ms": 0.0},
                         "nsig_field_backward_tc": {"n": 0, "ms": 0.0}, "nsig_field_backward_tc_masks": {"n": 0, "ms": 0.0}},
           "kernel_samples": S, "kernel_steps": n}
    main, other = b.rooflines(res, step_ms)
    peak_gbs, peak_tf, src = b._peaks()
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        pk = json.load(open(peaks_file))
        assert peak_gbs == pk["hbm_gbs"] and src.startswith("measured")
        assert peak_tf == pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))   # a kernel timed inside a long step
    want = b.ALG_FWD * S / (fwd_ms * 1e-3) / 1e9
    assert main["bound"] == "hbm" and main["unit"] == "GB/s" and main["peak"] == peak_gbs
    assert abs(main["achieved"] - want) < 1e-9 * want and abs(main["frac"] - want / peak_gbs) < 1e-12
    assert abs(main["avg_launch_ms"] - 0.25) < 1e-12 and main["samples_per_launch"] == 1_050_000.0
    assert abs(main["achieved"] - main["alg_bytes_per_sample"] * main["samples_per_launch"] / (main["avg_launch_ms"] * 1e-3) / 1e9) < 1e-6
    assert abs(main["share_of_step"] - 0.25 / 0.9) < 1e-12
    o = other[0]
    assert o["bound"] == "tensor" and "k_field_bwd_masks" in o["kernel"] and o["peak"] == peak_tf
    assert abs(o["achieved"] - b.FLOP_BWD * S / (bwd_ms * 1e-3) / 1e12) < 1e-9 * o["achieved"]
    assert abs(o["frac"] - o["achieved"] / peak_tf) < 1e-12 and abs(o["hbm_frac"] - o["hbm_achieved_gbs"] / peak_gbs) < 1e-12


def test_committed_bench_lines_are_self_consistent():
    """Every round-2 bench line under profiles/: the printed roofline fraction follows from the line's own per-launch
    figures, value = rays / step time, e2e is not a copy of value and declares its copies."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_n*_v*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r02_bench_n1_final_short.json")))
    assert paths
    checked = 0
    for p in paths:
        lines = [l for l in open(p) if l.startswith("{")]
        if not lines:
            continue
        d = json.loads(lines[-1])
        if d.get("impl") == "reference" or "roofline" not in d:
            continue
        r = d["roofline"]
        ach = r["alg_bytes_per_sample"] * r["samples_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9
        assert abs(ach - r["achieved"]) < 1e-6 * ach, p
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9, p
        rays = d["config"]["rays_per_step_per_gpu"] * d["n_gpus"]
        assert abs(d["value"] - rays / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"], p
        assert d["unit"] == "rays/s" and d["higher_is_better"] is True and d["warmup"] >= 3 and d["gpu_launches"] > 0
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"], p
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}, p
        checked += 1
    assert checked >= 4
