#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'])"; }
run "704 auto" A=1
run "640" GG_LIB=$PWD/gargammel_b200/libgg_t640.so
run "672" GG_LIB=$PWD/gargammel_b200/libgg_t672.so
run "640 w19" GG_LIB=$PWD/gargammel_b200/libgg_t640.so GG_FUSED_WARPS=19
