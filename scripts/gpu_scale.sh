#!/bin/bash
# 1/2/4/8-GPU weak-scaling sweep of bench.py on one box: gpurun --gpus 8 -- bash scripts/gpu_scale.sh
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; fi
  python - <<PY
import json
line=[l for l in open("gpurun_out/scale_$N.json").read().strip().splitlines() if l.startswith("{")][-1]
open("gpurun_out/r01_scale_n$N.json","w").write(line+"\n")
d=json.loads(line)
print("N=$N value %.3f G frag/s  ms/step %.3f  e2e %.1f M frag/s  host-bgzf %s  device-bgzf %s  clocks %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"].get("with_bgzf",{}).get("value"), d["e2e"].get("with_device_bgzf",{}).get("value"), d["clocks"]))
PY
done
