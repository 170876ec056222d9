#!/bin/bash
OUT=gpurun_out/sweep_libs.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" python bench.py --no-cpu-baseline --no-secondary --steps 3 2>>$OUT | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3f e2e %.3f kernel_ms %.2f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches']))" >> $OUT; }
run B2_DUMMY=1
for v in c3m4 c2m6 c3m6 c2m5; do run B2_LIB_PATH=$PWD/tests/emu/_build/libb2_$v.so; done
run B2_CHUNK_MRAYS=64
run B2_CHUNK_MRAYS=140
cat $OUT
