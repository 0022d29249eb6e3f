#!/bin/bash
# collider benches with the two-list collide + compute-sanitizer logs (1 GPU)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_reference_cases.py -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1
tail -3 gpurun_out/r2n_pytest.log | cut -c1-200
for sc in sand_torus dam_break mixed; do
  python bench.py --scene $sc --scale 0.125 --no-cpu --no-e2e --steps 60 > gpurun_out/r2n_bench_${sc}_0125.json 2>> gpurun_out/r2n_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/r2n_bench_${sc}_0125.json'));print('$sc', d['config']['particles_total'], 'ms/step', round(d['ms_per_step'],4), round(d['value']/1e9,3), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})"
done
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tests/tools/sanitize_target.py > gpurun_out/r2n_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r2n_sanitizer_$tool.log | cut -c1-200
done
