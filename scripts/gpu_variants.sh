#!/bin/bash
# tuning sweep: launch-bounds / tile-size variants of k_fused (bench only, no e2e, no cpu baseline)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15; GG_FUSED_NOFAST=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -3; GG_LIB=$PWD/gargammel_b200/libgg_minb4.so GG_FUSED_CTAS=4 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { echo "== $1"; shift; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'])"; }
run "minb3 F=auto ctas3" A=1
run "minb4 F=auto ctas4" GG_LIB=$PWD/gargammel_b200/libgg_minb4.so GG_FUSED_CTAS=4
run "minb4 F=160 ctas4" GG_LIB=$PWD/gargammel_b200/libgg_minb4.so GG_FUSED_CTAS=4 GG_FUSED_F=160


export GG_LIB=$PWD/gargammel_b200/libgg_minb4.so GG_FUSED_CTAS=4 GG_FUSED_F=160
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/var_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/r01e_fused $CMD > gpurun_out/r01e_ncu_full.log 2>&1
