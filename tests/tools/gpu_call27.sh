#!/bin/bash
# 1 GPU: smaller CTAs — G2P with 64 threads at 14 CTAs per SM (gt64), P2G with 2 warps at 14 CTAs per SM (pw2), both (both2)
mkdir -p gpurun_out
SVB200_LIB=$PWD/squishy_volumes_b200/lib/variants/both2.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
bash tests/tools/ab1.sh cur gt64 pw2 both2 2>&1 | tee gpurun_out/r2D_ab.txt
