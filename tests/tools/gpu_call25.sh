#!/bin/bash
# 2 GPUs, strict time limits: slab parity scene by scene (full logs per scene), multi-device handle, bench 2 GPUs
mkdir -p gpurun_out
: > gpurun_out/r2B_slabs.log
for sc in "energy_error 6" "jelly 30" "jelly_shear 30" "split_layers 25" "sand 40" "jelly_adaptive 40" "jelly_rebalance 30"; do
  echo "=== $sc" >> gpurun_out/r2B_slabs.log
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/slab_worker.py $sc > gpurun_out/r2B_one.log 2>&1
  echo "rc=$?" >> gpurun_out/r2B_slabs.log
  grep -v "^W1\|^E1\|torch/\|frozen\|^\s*\^\|OMP_NUM\|^\*\*\*\|elastic\|^  \(time\|host\|rank\|exitcode\|error_file\|traceback\)" gpurun_out/r2B_one.log | tail -40 >> gpurun_out/r2B_slabs.log
done
grep -c "within tolerance" gpurun_out/r2B_slabs.log
grep -n "===\|rc=\|FAILED\|Error\|assert" gpurun_out/r2B_slabs.log | cut -c1-250 | head -40
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2B_pytest_multi.log 2>&1
tail -3 gpurun_out/r2B_pytest_multi.log | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu --steps 200 > gpurun_out/r2B_bench_jelly2M_2gpu.json 2> gpurun_out/r2B_bench2.err
python -c "
import json;d=json.load(open('gpurun_out/r2B_bench_jelly2M_2gpu.json'));print('2gpu', d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],4), d['value']/1e9, d['e2e']['value']/1e9, d['slab_parity']['within_tolerance']); print(d['roofline']['stage_ms_per_substep_by_rank'])" || grep -n "Error" gpurun_out/r2B_bench2.err | head -5
