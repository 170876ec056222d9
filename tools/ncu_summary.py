#!/usr/bin/env python
"""Text summary of one `ncu --set full --import-source on` capture for profiles/: the counters the
design argues from (time, DRAM and L2 bytes / hit rates, issue utilisation, pipe utilisation, lanes
per instruction, local-memory instructions, stall reasons per issue, registers / occupancy limits)
and the per-function instruction / sample shares (tools/ncu_byfunc.py).
usage: ncu_summary.py <x.ncu-rep> <csrc dir> > profiles/<name>.txt"""
import csv
import io
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main():
    rep, csrc = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, vals = rows[0], rows[1], rows[2]
    name = vals[h.index("Kernel Name")] if "Kernel Name" in h else "?"
    print(f"== {os.path.basename(rep)}: {name}")
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"{k:92s} {units[i]:18s} {vals[i]}")
    t_ms = float(vals[h.index("gpu__time_duration.sum")].replace(",", "")) * {"ms": 1, "us": 1e-3, "msecond": 1, "usecond": 1e-3, "ns": 1e-6}.get(units[h.index("gpu__time_duration.sum")], 1)
    if "lts__t_bytes.sum" in h:
        i = h.index("lts__t_bytes.sum")
        b = float(vals[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
        print(f"{'derived: L2 throughput':92s} {'GB/s':18s} {b / (t_ms * 1e-3) / 1e9:.1f}")
    if "lts__t_sectors_srcunit_tex_op_read.sum" in h:
        sec = sum(float(vals[h.index(k)].replace(",", "")) for k in
                  ("lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum") if k in h)
        print(f"{'derived: L2 <-> SM throughput (sectors x 32 B)':92s} {'GB/s':18s} {sec * 32 / (t_ms * 1e-3) / 1e9:.1f}")
    dr = sum(float(vals[h.index(k)].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[h.index(k)], 1)
             for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in h)
    print(f"{'derived: DRAM throughput':92s} {'GB/s':18s} {dr / (t_ms * 1e-3) / 1e9:.1f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_summary_src.csv"
    open(tmp, "w").write(src)
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "ncu_byfunc.py"), tmp, csrc], capture_output=True, text=True).stdout
    print(out)


if __name__ == "__main__":
    main()
