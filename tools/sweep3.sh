#!/bin/bash
OUT=gpurun_out/sweep3.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" python bench.py --no-cpu-baseline --no-secondary --steps 5 $EXTRA 2>>$OUT | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3f e2e %.3f kernel_ms %.2f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches']))" >> $OUT; }
run B2_DUMMY=1
for v in rf2 rf4 rf16; do run B2_LIB_PATH=$PWD/tests/emu/_build/libb2_$v.so; done
EXTRA="--e2e-scans 1024"
run B2_E2E_CHUNKS=4
run B2_E2E_CHUNKS=8
run B2_E2E_CHUNKS=16
run B2_E2E_CHUNKS=32
cat $OUT
