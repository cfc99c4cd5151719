#!/bin/bash
# The BASELINE configurations at their named sizes through bench.py --config (one JSON line each): bash scripts/gpu_configs.sh <tag> [configs...]
TAG=${1:-r02}; shift
CFGS=${@:-"C1 C3 C4 C5 C5f C5s"}
mkdir -p gpurun_out; : > gpurun_out/${TAG}_configs.jsonl
for c in $CFGS; do
  case $c in
    C5f) A="--config C5 --c5-emit frag --steps 1 --no-cpu-baseline" ;;
    C5s) A="--config C5 --c5-sizes sizefreq --steps 1" ;;
    C5)  A="--config C5 --steps 1" ;;
    C3)  A="--config C3 --steps 2" ;;
    C4)  A="--config C4 --steps 2" ;;
    *)   A="--config $c" ;;
  esac
  echo "== $c: bench.py $A $EXTRA" >&2
  timeout 1500 python bench.py $A $EXTRA >> gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_$c.err || tail -5 gpurun_out/${TAG}_$c.err >&2
done
python - <<PY
import json
for l in open("gpurun_out/${TAG}_configs.jsonl"):
    if not l.startswith("{"): continue
    d = json.loads(l); e = d.get("e2e", {})
    print("%-110s %8.3f G frag/s  frac %.3f  e2e %7.1f M  (dev bgzf %7.1f M)  cpu %s" % (d["config"]["workload"][:110], d["value"] / 1e9, d["roofline"]["frac"], e.get("value", 0) / 1e6,
          (e.get("with_device_bgzf", {}).get("value") or 0) / 1e6, (d.get("cpu_baseline") or {}).get("value")))
PY
