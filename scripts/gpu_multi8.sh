#!/bin/bash
# N-GPU runs of bench.py through torchrun (one process per GPU): bash scripts/gpu_multi8.sh <tag> <N> "<bench args>" ...
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/${TAG}_n${N}.jsonl
P=29500
for A in "$@"; do
  P=$((P+1))
  echo "== N=$N bench.py $A" >&2
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N $A >> gpurun_out/${TAG}_n${N}.jsonl 2> gpurun_out/${TAG}_n${N}.err || tail -5 gpurun_out/${TAG}_n${N}.err >&2
done
python - <<PY
import json
for l in open("gpurun_out/${TAG}_n${N}.jsonl"):
    if not l.startswith("{"): continue            # (NCCL prints its version on stdout)
    d = json.loads(l); e = d.get("e2e", {})
    print("%-100s N=%d %8.3f G frag/s %.3f ms  e2e %7.1f M (frac of ceiling %s, ceiling %s GB/s)  dev-bgzf %7.1f M" % (d["config"]["workload"][:100], d["n_gpus"], d["value"] / 1e9, d["ms_per_step"], e.get("value", 0) / 1e6,
          (e.get("roofline") or {}).get("frac"), (e.get("roofline") or {}).get("pcie_ceiling_GBps"), (e.get("with_device_bgzf", {}).get("value") or 0) / 1e6))
PY
