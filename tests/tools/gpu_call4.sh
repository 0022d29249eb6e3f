#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2d_pytest.log 2>&1
tail -8 gpurun_out/r2d_pytest.log | cut -c1-200
bash tests/tools/ab.sh r2c cur > gpurun_out/r2d_ab.txt 2>&1
cat gpurun_out/r2d_ab.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2d_launches_adaptive.csv python bench.py --no-cpu --no-e2e --adaptive --steps 4 --warmup 2 > /dev/null 2> gpurun_out/r2d_ncu_adaptive.err
python profiles/summarize.py launches gpurun_out/r2d_launches_adaptive.csv | head -20
