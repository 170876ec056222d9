#!/bin/bash
OUT=gpurun_out/sweep2.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" python bench.py --no-cpu-baseline --no-secondary --steps 5 2>>$OUT | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3f e2e %.3f kernel_ms %.2f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches']))" >> $OUT; }
run B2_FAT_CLUSTERS=1
run B2_FAT_CLUSTERS=0
run B2_FAT_CLUSTERS=1
run B2_FAT_CLUSTERS=0
run B2_MESH_LEAF=1
run B2_MESH_LEAF=4
cat $OUT
