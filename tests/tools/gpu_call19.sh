#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
SVB_PARTS=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
bash tests/tools/ab_env1.sh "" "SVB_INVERT_ROWS=1" "SVB_PARTS=2" "SVB_PARTS=4" "SVB_PARTS=2 SVB_SPLIT_LAST=300"
