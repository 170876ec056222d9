#!/bin/bash
# usage: tools/prof_summary.sh <tag> [kernel-substring]   (reads gpurun_out/cast_<tag>.ncu-rep)
TAG=$1; K=${2:-cast_kernelILi0ELb1}
cd /root/repo/gpurun_out
ncu -i cast_$TAG.ncu-rep --page raw --csv 2>/dev/null > cast_raw_$TAG.csv
python ../tools/ncu_pick.py cast_raw_$TAG.csv gpu__time_duration.sum smsp__issue_active.avg.pct smsp__inst_executed.sum smsp__thread_inst_executed_per_inst sass__thread_inst_executed_true_per_opcode smsp__average_warps_issue_stalled sass__inst_executed_local sm__warps_active.avg.pct l1tex__t_sector_hit_rate sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active launch__occupancy_limit launch__registers_per_thread dram__bytes_read.sum dram__bytes_write.sum lts__t_bytes.sum 2>&1 | grep -v "TriageC\|modifier\|per_second\|pct_of_peak_sustained_elapsed"
ncu -i cast_$TAG.ncu-rep --page source --csv 2>/dev/null > cast_src_$TAG.csv
mkdir -p /tmp/elf && cd /tmp/elf && rm -f *.cubin && cuobjdump -xelf all /root/repo/mujoco_ros2_simulation_b200/libb2sense.so >/dev/null 2>&1
nvdisasm -g -c kernels.sm_100a.cubin > k.dis 2>/dev/null
python /root/repo/tools/ncu_lines.py /root/repo/gpurun_out/cast_src_$TAG.csv k.dis $K ${3:-36}
