#!/usr/bin/env python
"""bench.py — headline benchmark of the ray-batch render/train hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

A "step" is one watermark training step (utils_wtmk_disen.py:1164-1181): fresh random message, watermark
render pass over message_dim image blocks + content render pass over num_rays rays (both forward and
backward), HiDDeN decoder, losses, Adam.  metric = rays rendered (and back-propagated) per second, whole job.
N=1 workload = BASELINE.json configs[1] ("blender_wtmk").  For N>1 (torchrun) every rank runs the same
per-GPU batch on its own rays and gradients are all-reduced once per step: weak scaling.

`--impl reference` times the oracle port of the reference's pure-PyTorch CPU path (oracle/torch_port.py)
on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train rays/s (fwd+bwd)"
UNIT = "rays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="blender_wtmk")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (debug)")
    ap.add_argument("--no-render", action="store_true", help="skip the full-frame inference leg")
    ap.add_argument("--no-graph", action="store_true", help="eager step instead of the CUDA-graph-captured one")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused = optim.WatermarkAdam; torch = torch.optim.Adam over get_params (implies --no-graph)")
    ap.add_argument("--render-mode", default="merged", choices=["overlap", "merged", "split"],
                    help="merged = one render call over [block rays | content rays] (fastest: 1.16 ms); split = two render "
                         "calls like the reference trainer (block rays, then content rays), 1.30 ms; overlap = split with "
                         "the decoder chain on a side stream next to the content pass, 1.20 ms (the persistent field "
                         "kernels hold every SM's registers, so only march/composite really overlap)")
    ap.add_argument("--split-render", action="store_true", help="same as --render-mode split")
    ap.add_argument("--torch-decoder", action="store_true",
                    help="run the HiDDeN decoder as the plain PyTorch module under autocast instead of the fused kernels")
    ap.add_argument("--torch-losses", action="store_true",
                    help="clamp / MSE / BCE / weighting as plain torch expressions instead of the loss-head kernels")
    ap.add_argument("--cpu-rays", type=int, default=0, help="override the CPU sample size (rays per pass)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line), through NVML
    from a background thread every 5 ms (nvidia-smi -lms starts too slowly for a timed region of ~0.1 s);
    falls back to the nvidia-smi query loop when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.samples, self.stop_flag, self.thread, self.nvml = [], False, None, None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:
                return self.index
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self._visible_index()), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.samples.append((mhz, reasons))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            n = self.nvml
            masks = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            sm = [s[0] for s in self.samples]
            reasons = sorted(k for k, m in masks.items() if any(s[1] & m for s in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": reasons, "source": "nvml, 5 ms period, during the timed regions"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


# ---------------------------------------------------------------------------------------------------
# CPU port (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------
def cpu_port_run(cfg, steps, warmup, rays_per_pass, num_steps=512):
    """Time the oracle port of the reference CPU path on a bounded sample: `rays_per_pass` content rays +
    as many watermark-block rays (message_dim blocks of pH x pW pixels, pH*pW*md ~ rays_per_pass)."""
    import numpy as np
    import torch
    from nerf_signature_b200 import harness
    from nerf_signature_b200.nerf.hidden_models import get_hidden_decoder_multi_views
    from oracle import torch_port as tp

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    md = cfg["message_dim"]
    px = max(1, int(round((rays_per_pass / md) ** 0.5)))
    sub = dict(cfg)
    sub["H"], sub["W"] = px * cfg["num_rows"], px * cfg["num_cols"]   # pH = pW = px
    batch = {k: torch.from_numpy(v) for k, v in harness.make_batch(sub, seed=123, num_rays=rays_per_pass).items()}
    batch["rays_o"], batch["rays_d"], batch["gt"] = batch["rays_o"][0], batch["rays_d"][0], batch["gt"][0]
    field = tp.PortField(bound=cfg["bound"], message_dim=md, seed=0)
    torch.manual_seed(0)
    decoder = get_hidden_decoder_multi_views(num_bits=1, redundancy=1, num_blocks=8, input_ch=3, channels=64)
    params = [t for t in field.msg_tables] + list(decoder.parameters())
    opt = torch.optim.Adam(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    n_rays = rays_per_pass + md * px * px
    gen = torch.Generator().manual_seed(1)
    times = []
    for it in range(warmup + steps):
        message = torch.randint(0, 2, (md,), generator=gen).float()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        tp.train_step(field, decoder, batch, message, num_steps=num_steps)
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": n_rays * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{rays_per_pass} content + {md}x{px}x{px} watermark-block rays per step, {num_steps} uniform "
                      f"samples/ray (non-cuda_ray NeRFRenderer.run), {len(times)} timed steps, torch {cores} threads",
            "ms_per_step": 1e3 * total / len(times), "rays_per_step": n_rays}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nerf_signature_b200 import harness
    cfg = harness.CONFIGS[args.config]
    rays = args.cpu_rays or 16
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    # keep the whole run within minutes: ~1.5 s per step at 16+32 rays on 8 cores
    steps = min(steps, 40)
    r = cpu_port_run(cfg, steps, warmup, rays)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.config, **{k: cfg[k] for k in ("bound", "message_dim", "num_rays", "dt_gamma")},
                       "sample": r["sample"]},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: nerf_signature_b200 has no CPU path "
                         "(use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from nerf_signature_b200 import _build, _lib, harness
    if rank == 0 and _build.is_stale():
        _build.build_library()
    if world > 1:
        dist.barrier()
    cfg = dict(harness.CONFIGS[args.config])
    if args.config.startswith("shard"):
        cfg["num_rays"] = cfg["num_rays"] // world
    if args.split_render:
        args.render_mode = "split"
    use_graph = (not args.no_graph) and args.optimizer == "fused"
    scene = harness.Scene(cfg, dev, seed=0, optimizer=args.optimizer, graph=use_graph,
                          merged_render=args.render_mode == "merged", overlap_decoder=args.render_mode == "overlap",
                          fused_decoder=not args.torch_decoder,
                          fused_losses=not args.torch_losses)
    md = cfg["message_dim"]
    n_pool = 4  # distinct host batches cycled through (fresh rays every step)
    seed_rank = 0 if os.environ.get("NSIG_DIAG_SAME_RAYS") == "1" else rank  # diagnosis: identical work on every rank
    host_batches = [harness.make_batch(cfg, seed=1000 * seed_rank + i) for i in range(n_pool)]
    pinned = {k: torch.empty(v.shape, dtype=torch.float32).pin_memory() for k, v in host_batches[0].items()}
    pinned_msg = torch.empty(md, dtype=torch.float32).pin_memory()
    dev_batches = [scene.to_device(b) for b in host_batches]
    rays_per_step = host_batches[0]["rays_o"].shape[1] + int(np.prod(host_batches[0]["rays_o_block"].shape[:-1]))
    h2d_bytes = sum(v.nbytes for v in host_batches[0].values()) + md * 4
    gen = torch.Generator().manual_seed(7)  # identical message stream on every rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- phase A: device-resident inputs -> `value` --------------------------------------------------
    W, K = max(args.warmup, 3), max(args.steps, 1)
    _lib.timing_enable(["nsig_field_forward", "nsig_field_backward"])  # external events when captured
    for i in range(W):
        scene.train_step(dev_batches[i % n_pool], scene.new_message(gen))
    if not use_graph:
        _lib.timing_enable(["nsig_field_forward", "nsig_field_backward"])  # drop the warm-up events
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # NVML start-up (tens of ms, rank 0 only) BEFORE the barrier: ranks enter the timed region together
    barrier()
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        scene.train_step(dev_batches[i % n_pool], scene.new_message(gen))
    e1.record()
    barrier()
    launches = (scene.launches_per_step * K) if use_graph else (_lib.launch_count - launches0)
    ktimes = _lib.timing_read()
    n_samples_step, _ = scene.samples_per_step()
    spr = n_samples_step / rays_per_step
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if use_graph:
        # events inside a graph are re-recorded by every replay: the read above is the LAST timed step; average a
        # few more synchronised replays of the same graph for a steadier per-launch figure
        acc = {}
        for i in range(8):
            scene.train_step(dev_batches[i % n_pool], scene.new_message(gen))
            for name, d in _lib.timing_read().items():
                a = acc.setdefault(name, {"ms": 0.0, "n": 0})
                a["ms"] += d["ms"]; a["n"] += d["n"]
        for name, d in ktimes.items():
            acc[name]["ms"] += d["ms"]; acc[name]["n"] += d["n"]
        ktimes_avg, steps_timed = acc, 9
    else:
        ktimes_avg, steps_timed = ktimes, K

    # ---- phase B: host inputs through the public API, H2D in, loss D2H out -> `e2e` ----------------------
    # the host batches wait in pinned memory, packed like the captured step's input buffer (a pin_memory data loader):
    # a step is then ONE H2D copy (rays, ground truth and the fresh message) + graph replay + D2H read of the loss
    pool = [scene.pinned_batch(b) for b in host_batches] if use_graph else None
    if pool is not None:
        h2d_bytes = pool[0]["_flat"].numel() * 4

    def host_step(i):
        if pool is not None:
            pb = pool[i % n_pool]
            pb["message"].copy_(scene.new_message(gen))
            loss, _, _ = scene.train_step(pb, pb["message"])
            return float(loss)
        for k, v in host_batches[i % n_pool].items():
            pinned[k].copy_(torch.from_numpy(v))
        pinned_msg.copy_(scene.new_message(gen))
        loss, _, _ = scene.train_step(pinned, pinned_msg)   # H2D of the batch + message from pinned memory
        return float(loss)                                   # D2H read of the step's result

    for i in range(3):
        host_step(i)
    barrier()
    e0.record()
    for i in range(K):
        loss_host = host_step(i)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    _lib.timing_collect()

    # ---- full-frame inference (second half of BASELINE's metric): the 10 test views sharded over the ranks ------
    render = {}
    if not args.no_render:
        n_views = 10
        mine = [v for v in range(n_views) if v % world == rank] or [rank % n_views]
        for name in ("blender_800x800", "llff_1008x756"):
            fms, fsamples = harness.time_frames(name, dev, mine)
            t = torch.tensor([fms], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fms_max = float(t.item())
            render[name] = {"ms_per_frame": fms_max, "views": n_views, "frames_per_s_all_gpus": world * 1e3 / fms_max,
                            "samples_per_frame": fsamples,
                            "path": "NeRFRenderer.render(eval) -> nsig_render_rays (one persistent kernel per frame)"}

    if rank != 0:
        _finish(world)
        return

    # ---- roofline of the dominant kernel (fused field forward) -------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json, copy burst)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback (B200_PROFILING.md)")
    fwd = ktimes_avg.get("nsig_field_forward", {"ms": 0.0, "n": 0})
    bwd = ktimes_avg.get("nsig_field_backward", {"ms": 0.0, "n": 0})
    samples_per_step = n_samples_step  # both render passes, read from the march counters
    # SURVEY 8d algorithmic bytes per sample of the field forward: 24 (xyz, dir in) + 1024 (16 levels x 8 corners
    # x 8 B) + 64 (pre-summed message table gather) + 16 (sigma, rgb out)
    alg_bytes_per_sample = 24 + 1024 + 64 + 16
    calls_per_step = max(fwd["n"] / steps_timed, 1e-9)
    avg_ms = fwd["ms"] / max(fwd["n"], 1)
    samples_per_launch = samples_per_step / calls_per_step
    achieved = alg_bytes_per_sample * samples_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    step_ms = ms / K
    traffic = None
    try:  # ncu dram__bytes_read+write per launch of this kernel, from the committed capture of this round
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["k_field_fwd_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_field_fwd (nsig_field_forward)", "achieved": achieved, "peak": peak_gbs,
                "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": avg_ms, "samples_per_launch": samples_per_launch,
                "alg_bytes_per_sample": alg_bytes_per_sample,
                "share_of_step": (fwd["ms"] / steps_timed) / step_ms if step_ms > 0 else None,
                "field_backward_avg_launch_ms": bwd["ms"] / max(bwd["n"], 1),
                "field_backward_share_of_step": (bwd["ms"] / steps_timed) / step_ms if step_ms > 0 else None,
                "timed": "CUDA events around the launches on the launching stream" +
                         (" (external event nodes inside the captured step graph; mean of 9 replays)" if use_graph else "")}

    total_rays = rays_per_step * world
    line = {
        "metric": METRIC, "value": total_rays * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 MMA operands / f32 accumulate, f32 encoder+march+composite", "data": "synthetic",
        "config": {"workload": args.config, "bound": cfg["bound"], "scale": cfg["scale"], "dt_gamma": cfg["dt_gamma"],
                   "message_dim": md, "codebook": f'{cfg["num_rows"]}x{cfg["num_cols"]}',
                   "content_rays_per_gpu": cfg["num_rays"], "watermark_rays_per_gpu": rays_per_step - cfg["num_rays"],
                   "rays_per_step_per_gpu": rays_per_step, "mean_samples_per_ray": spr, "occupancy": cfg["occupancy"],
                   "weights": "random-init (tables U(+-1e-4), Xavier MLPs)",
                   "optimizer": ("WatermarkAdam (fused message-table Adam + torch fused Adam for the decoder)"
                                 if args.optimizer == "fused" else "torch.optim.Adam(fused)") + " + GradScaler",
                   "step": ("one CUDA graph replay per step" if use_graph else "eager") +
                           {"merged": "; both render passes in one call over [block rays | content rays]",
                            "split": "; two render calls (block rays, content rays), decoder in sequence",
                            "overlap": "; two render calls (block rays, content rays), decoder chain on a side stream "
                                       "next to the content pass"}[args.render_mode],
                   "l2": "inputs larger than L2: 64 MiB base tables + %d MiB message tables selected by a fresh message "
                         "each step + per-step sample buffers vs 126 MB L2" % (4 * md),
                   "decoder": "fused kernels (csrc/decoder.cu)" if scene.fused_decoder else "plain PyTorch module (autocast)",
                   "losses": "loss-head kernels (csrc/wtmk_loss.cu)" if scene.fused_losses else "plain torch expressions",
                   "parallelism": f"ray-sharded dp{world}", "exchange": scene.sync.exchange},
        "e2e": {"value": total_rays * K / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "clocks": clk, "roofline": roofline, "render": render,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_port_run(cfg, steps=12, warmup=1, rays_per_pass=args.cpu_rays or 64)  # ~12 s of CPU work
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """Leave without tearing NCCL down: destroy_process_group() can block on communicators that were captured
    into the step's CUDA graph; every rank has already passed the final barrier, so a hard exit is safe."""
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
