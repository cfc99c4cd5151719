#!/bin/bash
# quick GPU loop: parity tests + bench (+ optional full ncu capture of k_fused when $2 == ncu)
TAG=${1:-quick}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
if [ "$2" == "ncu" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/${TAG}_fused $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
