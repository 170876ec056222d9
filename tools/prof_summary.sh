#!/bin/bash
# usage: tools/prof_summary.sh <report-basename> <mangled-kernel-substring> [topN lines]
# reads gpurun_out/<report>.ncu-rep; joins with nvdisasm line info of the current libb2sense.so
REP=$1; K=$2
cd /root/repo/gpurun_out
ncu -i $REP.ncu-rep --page raw --csv 2>/dev/null > raw_$REP.csv
python ../tools/ncu_pick.py raw_$REP.csv gpu__time_duration.sum smsp__issue_active.avg.pct smsp__inst_executed.sum smsp__thread_inst_executed_per_inst sass__thread_inst_executed_true_per_opcode smsp__average_warps_issue_stalled sass__inst_executed_local sm__warps_active.avg.pct l1tex__t_sector_hit_rate sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active launch__occupancy_limit launch__registers_per_thread dram__bytes_read.sum dram__bytes_write.sum lts__t_bytes.sum 2>&1 | grep -v "TriageC\|modifier\|per_second\|pct_of_peak_sustained_elapsed\|stalled_\(drain\|membar\|misc\|sleeping\|tex\|lg_\|barrier\|dispatch\|mio\)"
ncu -i $REP.ncu-rep --page source --csv 2>/dev/null > src_$REP.csv
mkdir -p /tmp/elf && cd /tmp/elf && rm -f *.cubin && cuobjdump -xelf all /root/repo/mujoco_ros2_simulation_b200/libb2sense.so >/dev/null 2>&1
cat /dev/null > k.dis; for f in *.cubin; do nvdisasm -g -c $f >> k.dis 2>/dev/null; done
python /root/repo/tools/ncu_funcs.py /root/repo/gpurun_out/src_$REP.csv k.dis $K /root/repo/mujoco_ros2_simulation_b200/csrc | head -${3:-26}
