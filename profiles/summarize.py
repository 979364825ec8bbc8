"""Turn gpurun_out/*.ncu-rep (full-set captures) and launch lists into the tracked summaries under profiles/.

  python profiles/summarize.py <tag>        e.g. r01
Writes profiles/<name>_<tag>.summary.csv (selected raw metrics per captured launch) and
profiles/launches_*_<tag>.summary.csv (per-kernel totals and shares of the launch list).
"""
import csv
import glob
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def summarize_rep(path, tag):
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    if len(rows) < 3:
        print("no data in", path)
        return
    hdr, units = rows[0], rows[1]
    name = os.path.basename(path).replace(".ncu-rep", "")
    dst = os.path.join(PROF, name + ".summary.csv")
    kcol = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "metric", "unit", "value"])
        for li, vals in enumerate(rows[2:]):
            kern = vals[kcol] if kcol is not None else ""
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP:
                    w.writerow([li, kern[:90], h, u, v])
    print("wrote", dst)


def summarize_launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next((i for i, r in enumerate(rows) if r and r[0] == "ID"), None)
    if start is None:
        print("no launch table in", path)
        return
    hdr = rows[start]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            ns = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0][:100]
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + ns)
    total = sum(t for _, t in agg.values()) or 1.0
    dst = os.path.join(PROF, os.path.basename(path).replace(".csv", ".summary.csv"))
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_listed_gpu_time"])
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, c, f"{t / 1e3:.2f}", f"{t / 1e3 / c:.2f}", f"{t / total:.4f}"])
    print("wrote", dst)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    for p in sorted(glob.glob(os.path.join(OUT, f"*_{tag}.ncu-rep"))):
        summarize_rep(p, tag)
    for p in sorted(glob.glob(os.path.join(OUT, f"launches_*_{tag}.csv"))):
        summarize_launches(p)


if __name__ == "__main__":
    main()
