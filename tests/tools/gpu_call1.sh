#!/bin/bash
# round 2, first GPU call: parity at the BASELINE sizes, percentile report, bench line, collider-scene profile
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
python tests/tools/parity_report.py > gpurun_out/r2a_parity_percentiles.txt 2> gpurun_out/r2a_parity_report.err
python -m pytest tests -m gpu -q -s > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench_jelly1M.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench_jelly1M.json
for sc in sand_torus dam_break; do
  python bench.py --scene $sc --scale 0.125 --no-cpu --no-e2e --steps 60 > gpurun_out/r2a_bench_${sc}_0125.json 2>> gpurun_out/r2a_bench.err
done
ncu --set full --clock-control none --import-source on -k regex:'k_collide_big|k_collide_query|k_g2p|k_bin|k_meld|k_p2g' -s 40 -c 12 -o gpurun_out/r2a_sand1M -f python tests/tools/ncu_target.py sand_torus 8 0.125 > gpurun_out/r2a_ncu_sand.log 2>&1
tail -2 gpurun_out/r2a_ncu_sand.log
