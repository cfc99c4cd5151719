#!/bin/bash
# e2e of the device-BGZF pipelines against the chunk size: bash scripts/gpu_chunks.sh <tag> <chunk fragments>...
TAG=${1:-ck}; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --bgzf-chunk $c > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_$c.json")); e = d["e2e"]
print($c, "plain %.1f M" % (e["value"] / 1e6), "lz %.1f M" % (e["with_device_bgzf"]["value"] / 1e6), "huff %.1f M" % (e["with_device_bgzf_huffman"]["value"] / 1e6),
      "host zlib-1 %.1f M, %.3f bits/B" % (e["with_bgzf"]["value"] / 1e6, e["with_bgzf"].get("bits_per_byte", 0)), "lz bits/B %.3f" % e["with_device_bgzf"]["bits_per_byte"])
PY
done
