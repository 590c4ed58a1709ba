"""Turns ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/profile_summary.py launches <launches.csv> <out.md-fragment>
    python tools/profile_summary.py full <report.ncu-rep> <out.csv>
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_shared_loads",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.avg.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.avg.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip(), 1e-6)
        k = r[ki].split("(")[0].replace("void ", "").replace("gpmpc::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(t for _, t in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {c} | {t:.3f} | {100 * t / tot:.2f} % | {t / c:.4f} |\n")
    print(open(out).read())


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w") as f:
        f.write("metric,unit,value\n")
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or h == "Kernel Name":
                f.write(f'{h},{u},"{v}"\n')
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
