#!/bin/bash
# A/B on one box: bench each library variant at 1M and 8M.  usage: tests/tools/ab.sh name1 name2 ...   (names under squishy_volumes_b200/lib/variants/, "cur" = the in-tree build)
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = cur ]; then unset SVB200_LIB; else export SVB200_LIB=$PWD/squishy_volumes_b200/lib/variants/$v.so; fi
  for sc in 1 8; do
    st=100; [ $sc = 8 ] && st=40
    timeout 300 python bench.py --no-cpu --no-e2e --scale $sc --steps $st > gpurun_out/ab_${v}_${sc}.json 2>gpurun_out/ab_${v}_${sc}.err
    python -c "
import json;d=json.load(open('gpurun_out/ab_${v}_${sc}.json'));print('$v', d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" || tail -3 gpurun_out/ab_${v}_${sc}.err
  done
done
done
