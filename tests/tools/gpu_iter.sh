#!/bin/bash
# one GPU iteration: parity tests, bench line, optional ncu capture of kernels matching $1
mkdir -p gpurun_out
TAG=${2:-iter}
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6)
timeout 300 python bench.py --no-cpu > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print('ms/step', round(d['ms_per_step'],4), 'G/s', round(d['value']/1e9,3), {k:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v}, 'e2e', round(d['e2e']['value']/1e9,3))"
tail -3 gpurun_out/${TAG}_bench.err
if [ -n "$1" ]; then
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 3 -c 2 -o gpurun_out/${TAG}_ncu python tests/tools/ncu_target.py jelly_collision 6 > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
fi
