#!/bin/bash
# 1 GPU: what the driver runs at round end — build(), smoke(), the bench with a short step count
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2G_bench_steps20.json 2> gpurun_out/r2G_bench_steps20.err
python -c "
import json;d=json.load(open('gpurun_out/r2G_bench_steps20.json'));print('steps20', d['steps'], d['warmup'], 'ms/step', round(d['ms_per_step'],4), round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), d['clocks'], d['gpu_launches'])" || tail -5 gpurun_out/r2G_bench_steps20.err
