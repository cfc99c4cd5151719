#!/bin/bash
# one full ncu capture of the BGZF encoder with matches on the headline stream.  usage: bash scripts/gpu_ncu_lz.sh <tag> [fragments]
TAG=${1:-lz}; N=${2:-4000000}
mkdir -p gpurun_out
timeout 300 python scripts/bgzf_probe.py $N > gpurun_out/${TAG}_probe.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_lz -s 3 -c 1 -o gpurun_out/${TAG}_bgzl python scripts/bgzf_probe.py $N > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_probe.log; tail -3 gpurun_out/${TAG}_ncu.log
