#!/bin/bash
# 1 GPU, final build of round 2 (r2G): whole GPU suite, default bench line, adaptive line, ncu launch list + --set full of the four
# substep kernels, full-size collider scenes, parity percentiles
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2G_gpus.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2G_pytest.log 2>&1
tail -4 gpurun_out/r2G_pytest.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2G_bench_jelly1M.json 2> gpurun_out/r2G_bench.err
timeout 300 python bench.py --adaptive --no-cpu > gpurun_out/r2G_bench_jelly1M_adaptive.json 2>> gpurun_out/r2G_bench.err
timeout 300 python bench.py --scale 8 --no-cpu --steps 60 > gpurun_out/r2G_bench_jelly8M.json 2>> gpurun_out/r2G_bench.err
for f in r2G_bench_jelly1M r2G_bench_jelly1M_adaptive r2G_bench_jelly8M; do python -c "
import json;d=json.load(open('gpurun_out/$f.json'));print('$f', d['steps'], 'ms/step', round(d['ms_per_step'],4), round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['whole_substep']['frac'],3), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2G_launches_jelly1M.csv python bench.py --no-cpu --no-e2e --steps 8 --warmup 2 > /dev/null 2> gpurun_out/r2G_ncu1.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_p2g|k_g2p|k_offsets|k_invert_zero' -s 16 -c 8 -f -o gpurun_out/r2G_jelly1M python tests/tools/ncu_target.py jelly_collision 8 1.0 > gpurun_out/r2G_ncu2.log 2>&1
tail -2 gpurun_out/r2G_ncu2.log
timeout 300 python bench.py --scene sand_torus --scale 1 --no-cpu --no-e2e --steps 30 > gpurun_out/r2G_bench_sand8M.json 2>> gpurun_out/r2G_bench.err
timeout 300 python bench.py --scene dam_break --scale 1 --no-cpu --no-e2e --steps 30 > gpurun_out/r2G_bench_dam16M.json 2>> gpurun_out/r2G_bench.err
for f in r2G_bench_sand8M r2G_bench_dam16M; do python -c "
import json;d=json.load(open('gpurun_out/$f.json'));print('$f', d['config']['particles_total'], 'ms/step', round(d['ms_per_step'],4), round(d['value']/1e9,3), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['whole_substep']['frac'],3), {k[:8]:round(v,4) for k,v in d['roofline']['stage_ms_per_substep'].items() if v})" || tail -2 gpurun_out/r2G_bench.err; done
timeout 300 python tests/tools/parity_report.py > gpurun_out/r2G_parity_percentiles.txt 2> gpurun_out/r2G_parity_report.err
tail -3 gpurun_out/r2G_parity_percentiles.txt
