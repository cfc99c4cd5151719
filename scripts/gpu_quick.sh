#!/bin/bash
# Quick GPU check: parity tests, then a short bench (no CPU baseline / e2e).  usage: bash scripts/gpu_quick.sh <tag> [pytest args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu "$@" > gpurun_out/${TAG}_pytest.log 2>&1; tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
