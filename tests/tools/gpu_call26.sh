#!/bin/bash
# 1 GPU: L2 prefetch of the next iteration's / chunk's rows in G2P (pfg), P2G (pfp), both (pfb) against the current build
mkdir -p gpurun_out
SVB200_LIB=$PWD/squishy_volumes_b200/lib/variants/pfb.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
bash tests/tools/ab1.sh cur pfg pfp pfb 2>&1 | tee gpurun_out/r2C_ab.txt
