#!/bin/bash
# 1 GPU: collide / bin / advance kernels on quads, k_advance at 3 blocks per SM; G2P at 6 / 7 / 8 CTAs per SM (80 / 72 / 64 registers), P2G at 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_cases.py tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/r2A_pytest.log 2>&1
tail -3 gpurun_out/r2A_pytest.log | cut -c1-300
bash tests/tools/ab1.sh cur g6 g7 g8 p8 2>&1 | tee gpurun_out/r2A_ab.txt
for v in cur g7; do
  if [ "$v" = cur ]; then unset SVB200_LIB; else export SVB200_LIB=$PWD/squishy_volumes_b200/lib/variants/$v.so; fi
  for sc in sand_torus dam_break; do
    timeout 300 python bench.py --scene $sc --scale 0.125 --no-cpu --no-e2e --steps 60 > gpurun_out/r2A_${v}_${sc}.json 2>gpurun_out/r2A_${v}_${sc}.err
    python -c "
import json;d=json.load(open('gpurun_out/r2A_${v}_${sc}.json'));print('$v $sc', d['config']['particles_total'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" | tee -a gpurun_out/r2A_ab.txt
  done
  timeout 300 python bench.py --adaptive --no-cpu --no-e2e --steps 100 > gpurun_out/r2A_${v}_adaptive.json 2>gpurun_out/r2A_${v}_adaptive.err
  python -c "
import json;d=json.load(open('gpurun_out/r2A_${v}_adaptive.json'));print('$v adaptive', 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" | tee -a gpurun_out/r2A_ab.txt
done
