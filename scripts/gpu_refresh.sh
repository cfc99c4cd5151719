#!/bin/bash
# Refresh the pieces of the evidence bundle that depend on bench.py / the tests only (kernels unchanged): bash scripts/gpu_refresh.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --config C1 > gpurun_out/${TAG}_c1.json 2> gpurun_out/${TAG}_c1.err; cut -c1-200 gpurun_out/${TAG}_c1.json; tail -3 gpurun_out/${TAG}_c1.err
python bench.py --config C1 --frags 16000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c1_16m.json 2> gpurun_out/${TAG}_c1_16m.err; tail -3 gpurun_out/${TAG}_c1_16m.err
