#!/bin/bash
# 1 GPU: particle words in 16-byte quads (P2G / G2P move a particle with 9 / 6 + 9 accesses per lane): the whole GPU suite, A/B
# against b05afe6 (old) and the word layout with the table walk (w2), collider scenes old vs cur
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1
tail -5 gpurun_out/r2z_pytest.log | cut -c1-300
bash tests/tools/ab1.sh old w2 cur 2>&1 | tee gpurun_out/r2z_ab.txt
for v in old cur; do
  if [ "$v" = cur ]; then unset SVB200_LIB; else export SVB200_LIB=$PWD/squishy_volumes_b200/lib/variants/$v.so; fi
  for sc in sand_torus dam_break; do
    timeout 300 python bench.py --scene $sc --scale 0.125 --no-cpu --no-e2e --steps 60 > gpurun_out/r2z_${v}_${sc}.json 2>gpurun_out/r2z_${v}_${sc}.err
    python -c "
import json;d=json.load(open('gpurun_out/r2z_${v}_${sc}.json'));print('$v $sc', d['config']['particles_total'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" | tee -a gpurun_out/r2z_ab.txt
  done
  timeout 300 python bench.py --adaptive --no-cpu --no-e2e --steps 100 > gpurun_out/r2z_${v}_adaptive.json 2>gpurun_out/r2z_${v}_adaptive.err
  python -c "
import json;d=json.load(open('gpurun_out/r2z_${v}_adaptive.json'));print('$v adaptive', 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" | tee -a gpurun_out/r2z_ab.txt
done
