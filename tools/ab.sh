#!/bin/bash
# A/B of libb2sense.so builds on one GPU box: tools/ab.sh <tag> <variant.so>... ("default" = in-tree build)
# writes gpurun_out/<tag>_<name>.json (+ .err) per build and prints value / ms per step / kernel ms
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  name=$(basename "$v" .so)
  if [ "$v" = default ]; then unset B2_LIB_PATH; else export B2_LIB_PATH=$PWD/$v; fi
  python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - "$name" gpurun_out/${tag}_${name}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:14s} value {d['value']:.3f} Grays/s  ms/step {d['ms_per_step']:.3f}  kernel_ms {d['roofline']['kernel_ms']:.3f}  e2e {d['e2e']['value']:.3f}  ok {d['config']['e2e_matches_device_path']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
unset B2_LIB_PATH
