#!/usr/bin/env python
"""Summarises ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv           > profiles/rNN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [kernel-regex] [alg_bytes=N] > profiles/rNN_kernel.txt

alg_bytes=N adds the line `algorithmic_bytes byte N` (the algorithmic bytes of the captured launch), which
bench.py reads together with the dram__bytes lines for `roofline.traffic`.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__cycles_elapsed.avg.per_second",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warps_issue_stalled_lg_throttle_per_warp_active.pct",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        tot[r[ki]] += v
        cnt[r[ki]] += 1
    s = sum(tot.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-70s %6s %14s %8s %12s" % ("kernel", "n", "total_ns", "share", "avg_ns"))
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print("%-70s %6d %14.0f %8.4f %12.0f" % (k[:70], cnt[k], v, v / s, v / cnt[k]))


def full(path, pattern=None, alg_bytes=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pattern and pattern not in r[ni]:
            continue
        print("## %s  grid %s block %s" % (r[ni], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("%-75s %-14s %s" % (k, units[i], r[i]))
        if alg_bytes is not None:
            print("%-75s %-14s %s" % ("algorithmic_bytes", "byte", alg_bytes))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        rest = sys.argv[3:]
        alg = next((a.split("=", 1)[1] for a in rest if a.startswith("alg_bytes=")), None)
        pat = next((a for a in rest if not a.startswith("alg_bytes=")), None)
        full(sys.argv[2], pat, alg)
