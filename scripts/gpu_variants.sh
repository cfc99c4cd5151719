#!/bin/bash
# quick tuning loop: parity tests, then bench (no e2e / cpu baseline) under a few launch settings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'])"; }
run "auto" A=1
run "warps 18" GG_FUSED_WARPS=18
