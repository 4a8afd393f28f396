#!/usr/bin/env python
"""Side-by-side of selected raw metrics of the first kernel in several .ncu-rep files.
    python tools/ncu_compare.py a.ncu-rep b.ncu-rep ..."""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "memory_l1_tag_requests_global", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
cols = []
idxs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [0]
files = [a for a in sys.argv[1:] if not a.isdigit()]
if len(files) == 1 and len(idxs) > 1:
    files = files * len(idxs)
else:
    idxs = idxs * len(files) if len(idxs) == 1 else idxs
for f, idx in zip(files, idxs):
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, r = rows[0], rows[2 + idx]
    cols.append(dict(zip(hdr, r)))
    print("#", f, "->", cols[-1].get("Kernel Name", "")[:60])
for k in KEYS:
    print(f"{k[:78]:78s} " + " ".join(f"{c.get(k, '-'):>16s}" for c in cols))
