#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1
tail -4 gpurun_out/r2l_pytest.log | cut -c1-200
for sc in sand_torus dam_break mixed; do
  python bench.py --scene $sc --scale 0.125 --no-cpu --no-e2e --steps 60 > gpurun_out/r2l_bench_${sc}_0125.json 2>> gpurun_out/r2l_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/r2l_bench_${sc}_0125.json'));print('$sc', d['config']['particles_total'], 'ms/step', round(d['ms_per_step'],4), round(d['value']/1e9,3), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})"
done
for v in adv2 cur; do
  if [ "$v" = cur ]; then unset SVB200_LIB; else export SVB200_LIB=$PWD/squishy_volumes_b200/lib/variants/$v.so; fi
  python bench.py --adaptive --no-cpu --no-e2e > gpurun_out/r2l_adaptive_$v.json 2>> gpurun_out/r2l_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/r2l_adaptive_$v.json'));print('adaptive $v', d['steps'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})"
done
unset SVB200_LIB
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2l_launches_sand1M.csv python bench.py --scene sand_torus --scale 0.125 --no-cpu --no-e2e --steps 8 --warmup 2 > /dev/null 2> gpurun_out/r2l_ncu3.err
python profiles/summarize.py launches gpurun_out/r2l_launches_sand1M.csv | head -8
