#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1
tail -8 gpurun_out/r2i_pytest.log | cut -c1-200
bash tests/tools/ab_env.sh "SVB_PARTS=1 SVB_PDL=0" "SVB_PARTS=1" "SVB_PARTS=2" "SVB_PARTS=4" "" > gpurun_out/r2i_ab.txt 2>&1
cat gpurun_out/r2i_ab.txt
for sc in sand_torus dam_break mixed; do
  python bench.py --scene $sc --scale 0.125 --no-cpu --no-e2e --steps 60 > gpurun_out/r2i_bench_${sc}_0125.json 2>> gpurun_out/r2i_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/r2i_bench_${sc}_0125.json'));print('$sc', d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})"
done
