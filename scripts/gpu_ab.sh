#!/bin/bash
# A/B timing of library builds on the headline workload: bash scripts/gpu_ab.sh <lib.so> [<lib.so> ...]; "cur" = the in-tree build
mkdir -p gpurun_out
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'])"; }
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" == "cur" ]; then run "cur" A=1; else run "$lib" GG_LIB=$PWD/$lib; fi
done
done
