#!/bin/bash
# env-var sweeps of the cast launch shape on the bench workload (no rebuild needed)
OUT=gpurun_out/sweep_${1:-a}.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" python bench.py --no-cpu-baseline --no-secondary --steps 3 2>>$OUT | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3f e2e %.3f kernel_ms %.2f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches']))" >> $OUT; }
run B2_DUMMY=1
run B2_STAGE=0
run B2_TILES_PER_BLOCK=16
run B2_TILES_PER_BLOCK=32
run B2_TILES_PER_BLOCK=128
run B2_CHUNK_MRAYS=4
run B2_CHUNK_MRAYS=16
run B2_CHUNK_MRAYS=32
run B2_SPLIT=0
run B2_E2E_CHUNKS=2
run B2_E2E_CHUNKS=16
cat $OUT
