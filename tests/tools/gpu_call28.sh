#!/bin/bash
# 1 GPU: CTA shapes — P2G with 8 warps at 3 / 4 CTAs per SM, 6 warps at 4; G2P 64 threads at 16 CTAs per SM (64 registers)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
SVB200_LIB=$PWD/squishy_volumes_b200/lib/variants/pw8c3.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
bash tests/tools/ab1.sh cur pw8c3 pw8c4 pw6c4 g64c16 2>&1 | tee gpurun_out/r2E_ab.txt
