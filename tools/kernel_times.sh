#!/bin/bash
# per-kernel device times of ONE cast step (tools/traffic_run.py) for a given libb2sense.so build:
#   tools/kernel_times.sh <tag> <variant.so | default> [scans]
# ncu times are cold-cache and serialised: compare SHARES and per-kernel deltas between builds.
tag=$1; v=$2; scans=${3:-256}
if [ "$v" = default ]; then unset B2_LIB_PATH; else export B2_LIB_PATH=$PWD/$v; fi
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/${tag}_times.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum \
    python tools/traffic_run.py --scans $scans > gpurun_out/${tag}_times.log 2>&1
python tools/ncu_traffic.py gpurun_out/${tag}_times.csv $scans "$tag" > gpurun_out/${tag}_traffic.json
python - gpurun_out/${tag}_traffic.json "$tag" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[2], "cast ms (serialised)", round(d["cast_ms_serialised"], 3), "DRAM B/ray", round(d["cast_dram_bytes_per_ray"], 2),
      "write B/ray", round(d["cast_dram_write_bytes_per_ray"], 2))
for k, v in d["kernels"].items():
    print(f"   {k:22s} x{v['launches']} {v['ms']:8.3f} ms  {v['warp_inst']/1e6:9.1f} Minst  dram {v['dram_bytes_per_ray']:6.2f} B/ray  L2 {v['l2_gbs']:8.1f} GB/s")
PY
unset B2_LIB_PATH
