#!/bin/bash
# like ab_env.sh, 1 M particles only, 3 repetitions interleaved
mkdir -p gpurun_out
for rep in 1 2 3; do
for cfg in "$@"; do
    tag=$(echo "$cfg" | tr ' =' '__'); [ -z "$tag" ] && tag=default
    env $cfg timeout 300 python bench.py --no-cpu --no-e2e --steps 200 > gpurun_out/abe_${tag}_1.json 2>gpurun_out/abe_${tag}_1.err
    python -c "
import json;d=json.load(open('gpurun_out/abe_${tag}_1.json'));print('[$cfg]', d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" || tail -3 gpurun_out/abe_${tag}_1.err
done
done
