#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1
tail -15 gpurun_out/r2c_pytest.log | cut -c1-200
bash tests/tools/ab.sh r2a r2b cur > gpurun_out/r2c_ab.txt 2>&1
cat gpurun_out/r2c_ab.txt
for sc in sand_torus dam_break mixed; do
  python bench.py --scene $sc --scale 0.125 --no-cpu --no-e2e --steps 60 > gpurun_out/r2c_bench_${sc}_0125.json 2>> gpurun_out/r2c_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/r2c_bench_${sc}_0125.json'));print('$sc', d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})"
done
python bench.py --no-cpu --adaptive > gpurun_out/r2c_bench_jelly1M_adaptive.json 2> gpurun_out/r2c_bench_adaptive.err
tail -c 400 gpurun_out/r2c_bench_adaptive.err
python -c "
import json;d=json.load(open('gpurun_out/r2c_bench_jelly1M_adaptive.json'));print('adaptive', d['steps'], 'ms/step', round(d['ms_per_step'],4), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v}, d['e2e']['value']/1e9)"
