#!/bin/bash
# run with: gpurun --gpus 2 -- bash scripts/gpu_two.sh   (the --gpus 2 byte-identity test of the fused driver, skipped on one-GPU boxes)
mkdir -p gpurun_out
nvidia-smi -L
timeout 200 python -m pytest tests/test_cli.py -x -q -m gpu -k "two_gpus or art_hand_off" -rs > gpurun_out/r02_two_gpus_pytest.log 2>&1; tail -5 gpurun_out/r02_two_gpus_pytest.log
