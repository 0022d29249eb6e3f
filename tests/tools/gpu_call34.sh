#!/bin/bash
# 1 GPU: upload and download queue all copies back to back: parity subset + the short bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_replay.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2J_bench_steps20.json 2> gpurun_out/r2J_bench_steps20.err
python -c "
import json;d=json.load(open('gpurun_out/r2J_bench_steps20.json'));print('steps20', d['steps'], d['warmup'], 'ms/step', round(d['ms_per_step'],4), round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), d['e2e']['seconds_per_pass'])" || tail -5 gpurun_out/r2J_bench_steps20.err
