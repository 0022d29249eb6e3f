#!/bin/bash
# 1 GPU: quick parity on the ordered work lists, then A/B of the ordering at 1 M / 8 M particles
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -3
bash tests/tools/ab_env.sh "" "SVB_WORK_ORDER=0"
