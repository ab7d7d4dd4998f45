#!/usr/bin/env python
"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep files stay in gpurun_out/).
  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof_x.ncu-rep > profiles/r01_x.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[h]
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        d = dict(zip(H, r))
        name = d["Kernel Name"].split("(")[0]
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
        a = agg.setdefault(name, [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += v
    ours = {k: v for k, v in agg.items() if "at::" not in k}
    tot = sum(v[1] for v in ours.values()) or 1.0
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})")
    print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's live event times")
    print(f"{'kernel':46s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}  grid / block (last)")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:46s} {v[0]:8d} {v[1]:10.3f} {v[1] / v[0]:9.4f} {v[1] / tot:7.3f}  {v[2]} / {v[3]}")
    oth = sum(v[1] for k, v in agg.items() if "at::" in k)
    print(f"# torch data-generation kernels (not the product, outside the timed region): {oth:.3f} ms total")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none  ({path}); one column per captured launch")
    kn = H.index("Kernel Name")
    print("kernel:", [r[kn].split("(")[0] for r in rows[2:]])
    for k in KEYS:
        if k in H:
            i = H.index(k)
            print(f"{k:78s} {U[i]:16s} " + "  ".join(r[i] for r in rows[2:]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
