#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2k_gpus.txt
python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1
tail -6 gpurun_out/r2k_pytest.log | cut -c1-200
python bench.py > gpurun_out/r2k_bench_jelly1M.json 2> gpurun_out/r2k_bench.err
python bench.py --adaptive --no-cpu > gpurun_out/r2k_bench_jelly1M_adaptive.json 2>> gpurun_out/r2k_bench.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r2k_reference_arm_jelly1M.json 2>> gpurun_out/r2k_bench.err
for f in r2k_bench_jelly1M r2k_bench_jelly1M_adaptive; do python -c "
import json;d=json.load(open('gpurun_out/$f.json'));print('$f', d['steps'], 'ms/step', round(d['ms_per_step'],4), round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['whole_substep']['frac'],3), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2k_launches_jelly1M.csv python bench.py --no-cpu --no-e2e --steps 8 --warmup 2 > /dev/null 2> gpurun_out/r2k_ncu1.err
ncu --set full --clock-control none --import-source on -k regex:'k_p2g|k_g2p|k_offsets|k_invert_zero' -s 16 -c 8 -f -o gpurun_out/r2k_jelly1M python tests/tools/ncu_target.py jelly_collision 8 1.0 > gpurun_out/r2k_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2k_launches_sand1M.csv python bench.py --scene sand_torus --scale 0.125 --no-cpu --no-e2e --steps 8 --warmup 2 > /dev/null 2> gpurun_out/r2k_ncu3.err
ncu --set full --clock-control none --import-source on -k regex:'k_collide_cand|k_collide_query|k_bin|k_meld' -s 12 -c 8 -f -o gpurun_out/r2k_sand1M python tests/tools/ncu_target.py sand_torus 6 0.125 > gpurun_out/r2k_ncu4.log 2>&1
for sc in sand_torus dam_break mixed; do
  python bench.py --scene $sc --scale 0.125 --no-cpu --no-e2e --steps 60 > gpurun_out/r2k_bench_${sc}_0125.json 2>> gpurun_out/r2k_bench.err
done
python bench.py --scene sand_torus --scale 1 --no-cpu --steps 30 > gpurun_out/r2k_bench_sand8M.json 2>> gpurun_out/r2k_bench.err
python bench.py --scene dam_break --scale 1 --no-cpu --steps 30 > gpurun_out/r2k_bench_dam16M.json 2>> gpurun_out/r2k_bench.err
python bench.py --scene mixed --scale 1 --no-cpu --no-e2e --steps 20 --warmup 3 > gpurun_out/r2k_bench_mixed64M.json 2>> gpurun_out/r2k_bench.err
for f in r2k_bench_sand_torus_0125 r2k_bench_dam_break_0125 r2k_bench_mixed_0125 r2k_bench_sand8M r2k_bench_dam16M r2k_bench_mixed64M; do python -c "
import json;d=json.load(open('gpurun_out/$f.json'));print('$f', d['config']['particles_total'], 'ms/step', round(d['ms_per_step'],4), round(d['value']/1e9,3), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['whole_substep']['frac'],3), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" || tail -2 gpurun_out/r2k_bench.err; done
python tests/tools/parity_report.py > gpurun_out/r2k_parity_percentiles.txt 2> gpurun_out/r2k_parity_report.err
