#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/launches_rNN.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep > profiles/kernel_rNN.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_src_fp64.sum.per_second", "sm__ops_path_tensor_src_fp64.sum.peak_sustained",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("".join(lines))):
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        agg[row["Kernel Name"][:90]][0] += 1
        agg[row["Kernel Name"][:90]][1] += val
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time "
          f"(ncu: cold caches, serialised -- compare shares)")
    print(f"{'ms':>10} {'launches':>9} {'share':>7} {'us/launch':>10}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:10.2f} {v[0]:9d} {100 * v[1] / tot:6.1f}% {v[1] / v[0]:10.1f}  {k}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    print(f"# {path}")
    for r in rows[2:]:
        print(f"## {r[name_col][:100]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:90s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
