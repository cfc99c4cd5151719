#!/bin/bash
TAG=${1:-c1}
mkdir -p gpurun_out
CMD="python bench.py --config C1 --frags 16000000 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_deam_stream -s 3 -c 1 -o gpurun_out/${TAG}_deam $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-200
