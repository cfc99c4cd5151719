#!/bin/bash
# Device BGZF with matches: its parity tests, then the probe (both encoders) on the headline stream.  usage: bash scripts/gpu_lz.sh <tag>
TAG=${1:-lz}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "bgzf" > gpurun_out/${TAG}_pytest.log 2>&1; tail -25 gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/bgzf_probe.py 4000000 > gpurun_out/${TAG}_probe.log 2>&1; tail -4 gpurun_out/${TAG}_probe.log
GG_BGZF_MODE=huff timeout 300 python scripts/bgzf_probe.py 4000000 > gpurun_out/${TAG}_probe_huff.log 2>&1; tail -2 gpurun_out/${TAG}_probe_huff.log
