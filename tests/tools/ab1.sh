#!/bin/bash
# A/B of library variants, 1 M particles, 3 interleaved repetitions (+ one 8 M run each).  usage: tests/tools/ab1.sh cur name1 ...
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in "$@"; do
  if [ "$v" = cur ]; then unset SVB200_LIB; else export SVB200_LIB=$PWD/squishy_volumes_b200/lib/variants/$v.so; fi
  for sc in 1 8; do
    [ $sc = 8 ] && [ $rep != 1 ] && continue
    st=200; [ $sc = 8 ] && st=40
    timeout 300 python bench.py --no-cpu --no-e2e --scale $sc --steps $st > gpurun_out/ab_${v}_${sc}.json 2>gpurun_out/ab_${v}_${sc}.err
    python -c "
import json;d=json.load(open('gpurun_out/ab_${v}_${sc}.json'));print('$v', d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" || tail -3 gpurun_out/ab_${v}_${sc}.err
  done
done
done
