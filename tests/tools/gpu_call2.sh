#!/bin/bash
# round 2, second GPU call: binning-ahead build — parity, A/B against the previous build, adaptive bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s > gpurun_out/r2b_pytest.log 2>&1
tail -5 gpurun_out/r2b_pytest.log
bash tests/tools/ab.sh r2a cur > gpurun_out/r2b_ab.txt 2>&1
cat gpurun_out/r2b_ab.txt
python bench.py --no-cpu --adaptive > gpurun_out/r2b_bench_jelly1M_adaptive.json 2> gpurun_out/r2b_bench_adaptive.err
tail -c 300 gpurun_out/r2b_bench_adaptive.err
python tests/tools/parity_report.py > gpurun_out/r2b_parity_percentiles.txt 2> gpurun_out/r2b_parity_report.err
