#!/bin/bash
# Run on the GPU box (via gpurun): tests, bench, then the ncu launch list and one full capture of k_fused.
# usage: bash scripts/gpu_bench_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.csv
nproc > gpurun_out/${TAG}_nproc.txt
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/${TAG}_fused $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1

# device BGZF kernel: plain run, then one full capture (second launch: warm)
python scripts/bgzf_probe.py > gpurun_out/${TAG}_bgzf_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bgzf_deflate -s 1 -c 1 -o gpurun_out/${TAG}_bgzf python scripts/bgzf_probe.py > gpurun_out/${TAG}_ncu_bgzf.log 2>&1
python tests/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
ls -la gpurun_out/ | tail -30
