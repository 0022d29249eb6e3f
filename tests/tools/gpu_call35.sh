#!/bin/bash
# 1 GPU: G2P stores the two used words of quad 8 only (cur) against the build before (base)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
bash tests/tools/ab1.sh cur base 2>&1 | tee gpurun_out/r2I_ab.txt
