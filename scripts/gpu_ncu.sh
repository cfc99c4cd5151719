#!/bin/bash
# One full ncu capture of k_fused on the headline workload (after a plain run exited 0).  usage: bash scripts/gpu_ncu.sh <tag> [env...]
TAG=${1:-n}; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
env "$@" $CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/${TAG}_fused $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_plain.log | cut -c1-300
