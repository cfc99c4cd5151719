#!/bin/bash
# Run on the GPU box (via gpurun): tests, bench (+ the reference arm), then the ncu launch list and one full capture each of
# k_fused (C2), k_deam_stream (C1 at 16 M records) and k_bgzf_lz.  usage: bash scripts/gpu_bench_profile.sh <tag> [notests]
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.csv
nproc > gpurun_out/${TAG}_nproc.txt
if [ "$2" != "notests" ]; then
  python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
fi
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cut -c1-400 gpurun_out/${TAG}_bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/${TAG}_fused $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1

# stand-alone deamSim (BASELINE configs[0] shape at 16 M records): plain run with its JSON line, launch list, one full capture
C1="python bench.py --config C1 --frags 16000000 --steps 2 --warmup 3 --no-cpu-baseline"
$C1 > gpurun_out/${TAG}_c1_16m.json 2> gpurun_out/${TAG}_c1_16m.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_c1_launches.csv $C1 > gpurun_out/${TAG}_ncu_c1_launches.log 2>&1
$C1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_deam_stream -s 3 -c 1 -o gpurun_out/${TAG}_deam $C1 > gpurun_out/${TAG}_ncu_deam.log 2>&1

# device BGZF encoder (k_bgzf_lz<false>; launches alternate counting pass / encoder): plain run of both encoders, then one full capture (fourth launch: the encoder, warm)
GG_BGZF_MODE=huff python scripts/bgzf_probe.py > gpurun_out/${TAG}_bgzf_probe_huff.log 2>&1
python scripts/bgzf_probe.py > gpurun_out/${TAG}_bgzf_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bgzf_lz -s 3 -c 1 -o gpurun_out/${TAG}_bgzf python scripts/bgzf_probe.py > gpurun_out/${TAG}_ncu_bgzf.log 2>&1
ls -la gpurun_out/ | grep ${TAG}_
