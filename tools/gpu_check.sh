#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, then the ncu launch list and one full capture of the
# cast kernel (each ncu pass only after the same command exited 0 without ncu).
# usage: tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke_$TAG.log
python bench.py $BENCH_ARGS > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench_$TAG.json
if [ "${REF:-1}" = "1" ]; then python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"; fi
if [ "${NCU:-1}" = "1" ]; then
  BCMD="python bench.py --steps 2 --warmup 3 --scans 256 --e2e-scans 32 --no-cpu-baseline"
  $BCMD > $OUT/plain_$TAG.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv $BCMD > $OUT/ncu_launches_$TAG.log 2>&1
  echo "ncu launches rc=$?"
  $BCMD > $OUT/plain2_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:drain_kernel -s 4 -c 1 -o $OUT/drain_$TAG $BCMD > $OUT/ncu_drain_$TAG.log 2>&1
  echo "ncu drain rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:collect_kernel -s 4 -c 1 -o $OUT/collect_$TAG $BCMD > $OUT/ncu_collect_$TAG.log 2>&1
  echo "ncu collect rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:mesh_pairs -s 2 -c 1 -o $OUT/mesh_$TAG $BCMD > $OUT/ncu_mesh_$TAG.log 2>&1
  echo "ncu mesh rc=$?"
  if [ "${NCU_DEPTH:-1}" = "1" ]; then ncu --set full --clock-control none --import-source on -k regex:depth_to_cloud -s 3 -c 1 -o $OUT/depth_$TAG $BCMD > $OUT/ncu_depth_$TAG.log 2>&1; echo "ncu depth rc=$?"; fi
fi
ls -la $OUT | tail -20
