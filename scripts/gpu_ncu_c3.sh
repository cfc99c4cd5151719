#!/bin/bash
# C3-shaped capture (Briggs, genomes larger than L2) at a reduced size: bash scripts/gpu_ncu_c3.sh <tag> [scale]
TAG=${1:-c3}; SC=${2:-0.1}
mkdir -p gpurun_out
CMD="python bench.py --config C3 --scale $SC --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/${TAG}_fused $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
