#!/bin/bash
# 1 GPU: pair-packed P2G walk (SVB_P2G_WALK=3): parity, A/B (old = b05afe6, w2 = table walk, cur = pair walk, w3c6 = pair walk at 6 CTAs/SM,
# w3al = pair walk + line-aligned G2P warps), layout micro-benchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_cases.py tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1
tail -4 gpurun_out/r2y_pytest.log | cut -c1-300
SVB200_LIB=$PWD/squishy_volumes_b200/lib/variants/w3al.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
bash tests/tools/ab1.sh old w2 cur w3c6 w3al 2>&1 | tee gpurun_out/r2y_ab.txt
timeout 120 profiles/bin/gather_bench 2>&1 | tee gpurun_out/r2y_gather_bench.txt
