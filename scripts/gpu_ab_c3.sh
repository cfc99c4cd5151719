#!/bin/bash
# A/B of library builds on the Briggs configuration (C3', 2 x 500 Mb): bash scripts/gpu_ab_c3.sh <lib.so> ...
for lib in "$@"; do
  echo "== $lib"
  if [ "$lib" == "cur" ]; then python tests/bench_configs.py --skip c1,c5,c4 2>/dev/null; else GG_LIB=$PWD/$lib python tests/bench_configs.py --skip c1,c5,c4 2>/dev/null; fi | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][-32:], d['ms'], round(d['fragments_per_s']/1e9,3))"
done
