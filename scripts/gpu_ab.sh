#!/bin/bash
# A/B timing of library builds on the headline workload: bash scripts/gpu_ab.sh <lib.so|cur>[:ENV=V[,ENV=V]] ...; "cur" = the in-tree build
mkdir -p gpurun_out
run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'])"; }
for rep in 1 2; do
for spec in "$@"; do
  lib=${spec%%:*}; envs=""; [ "$spec" != "$lib" ] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  if [ "$lib" == "cur" ]; then run "$spec" A=1 $envs; else run "$spec" GG_LIB=$PWD/$lib $envs; fi
done
done
