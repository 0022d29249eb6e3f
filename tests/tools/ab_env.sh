#!/bin/bash
# A/B of environment switches of ONE library build on one box: each argument is a quoted list of VAR=value settings ("" = defaults)
# usage: tests/tools/ab_env.sh "" "SVB_PDL=0" "SVB_PARTS=1" ...
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "$@"; do
  for sc in 1 8; do
    st=100; [ $sc = 8 ] && st=40
    tag=$(echo "$cfg" | tr ' =' '__'); [ -z "$tag" ] && tag=default
    env $cfg timeout 300 python bench.py --no-cpu --no-e2e --scale $sc --steps $st > gpurun_out/abe_${tag}_${sc}.json 2>gpurun_out/abe_${tag}_${sc}.err
    python -c "
import json;d=json.load(open('gpurun_out/abe_${tag}_${sc}.json'));print('[$cfg]', d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" || tail -3 gpurun_out/abe_${tag}_${sc}.err
  done
done
done
