#!/bin/bash
# 1 GPU: work items one ahead (P2G / G2P) + table-driven P2G walk: parity, A/B against the previous build, racecheck
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_cases.py tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1
tail -4 gpurun_out/r2x_pytest.log | cut -c1-300
bash tests/tools/ab1.sh old walk1 cur
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 3 python tests/tools/sanitize_target.py > gpurun_out/r2x_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2x_sanitizer_racecheck.log | cut -c1-200
