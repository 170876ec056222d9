#!/usr/bin/env python
"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` output (first kernel row)."""
import csv
import sys

DEFAULT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__grid_size', 'launch__shared_mem',
           'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum ', 'smsp__thread_inst_executed_per_inst_executed.ratio',
           'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum ', 'lts__throughput.avg.pct',
           'l1tex__throughput.avg.pct', 'smsp__average_warp', 'smsp__warps_eligible.avg.per_cycle_active',
           'sm__inst_executed_pipe_fp64', 'sm__pipe_fp64', 'sm__inst_executed_pipe_', 'smsp__inst_executed_pipe_',
           'local_load', 'local_store', 'lmem', 'sm__cycles_active.avg ', 'smsp__pcsamp_warps_issue_stalled',
           'smsp__average_warps_issue_stalled', 'thread_inst_executed_true', 'sm__sass_thread_inst_executed_op_d',
           'sm__sass_thread_inst_executed_op_f', 'smsp__sass_thread_inst_executed_op']


def main():
    path = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[h.index('Kernel Name')][:80] if 'Kernel Name' in h else '')
        for i, name in enumerate(h):
            if any(p.strip() in name for p in pats):
                print(f"{name:95s} {u[i]:18s} {r[i]}")


if __name__ == '__main__':
    main()
