#!/bin/bash
# A/B of library variants built into tests/emu/_build/libb2_<name>.so (same box, same session)
OUT=gpurun_out/sweep_libs.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" python bench.py --no-cpu-baseline --no-secondary --steps 5 --e2e-scans 256 2>>$OUT | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3f e2e %.3f kernel_ms %.2f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches']))" >> $OUT; }
run B2_DUMMY=1
for f in tests/emu/_build/libb2_*.so; do run B2_LIB_PATH=$PWD/$f; done
run B2_DUMMY=2
cat $OUT
