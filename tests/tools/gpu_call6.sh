#!/bin/bash
# 2 GPUs, strict time limits: slab parity scene by scene, the single-process multi-device handle, the weak-scaling bench line
mkdir -p gpurun_out
: > gpurun_out/r2f_slabs.log
for sc in "jelly 30" "jelly_shear 30" "split_layers 25" "sand 40" "jelly_adaptive 40" "energy_error 6" "jelly_rebalance 30"; do
  echo "=== $sc" >> gpurun_out/r2f_slabs.log
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/slab_worker.py $sc 2>&1 | grep -v "^W1\|^E1\|torch/\|frozen\|^\s*\^\|OMP_NUM\|^\*\*\*" | tail -25 >> gpurun_out/r2f_slabs.log
  echo "rc=${PIPESTATUS[0]}" >> gpurun_out/r2f_slabs.log
done
grep -c "within tolerance" gpurun_out/r2f_slabs.log
grep -n "===\|rc=\|FAILED\|Error" gpurun_out/r2f_slabs.log | cut -c1-200 | head -40
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s > gpurun_out/r2f_pytest_multi.log 2>&1
tail -6 gpurun_out/r2f_pytest_multi.log | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu --steps 100 > gpurun_out/r2f_bench_jelly2M_2gpu.json 2> gpurun_out/r2f_bench2.err
python -c "
import json;d=json.load(open('gpurun_out/r2f_bench_jelly2M_2gpu.json'));print('2gpu', d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],4), d['value']/1e9, d['e2e']['value']/1e9, d['slab_parity']); print(d['roofline']['stage_ms_per_substep_by_rank'])" || grep -n "Error" gpurun_out/r2f_bench2.err | head -5
