#!/bin/bash
# 2 GPUs, final build (r2G): slab parity scene by scene, multi-device handle, weak-scaling jelly line, 16 M dam break on 2 GPUs
mkdir -p gpurun_out
: > gpurun_out/r2G_slabs.log
for sc in "energy_error 6" "jelly 30" "jelly_shear 30" "split_layers 25" "sand 40" "jelly_adaptive 40" "jelly_rebalance 30"; do
  echo "=== $sc" >> gpurun_out/r2G_slabs.log
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/slab_worker.py $sc > gpurun_out/r2G_one.log 2>&1
  echo "rc=$?" >> gpurun_out/r2G_slabs.log
  grep -v "^W1\|^E1\|torch/\|frozen\|^\s*\^\|OMP_NUM\|^\*\*\*\|elastic\|^  \(time\|host\|rank\|exitcode\|error_file\|traceback\)" gpurun_out/r2G_one.log | tail -40 >> gpurun_out/r2G_slabs.log
done
grep -c "within tolerance" gpurun_out/r2G_slabs.log
grep -n "rc=\|FAILED\|Error\|assert" gpurun_out/r2G_slabs.log | cut -c1-250 | head -20
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2G_pytest_multi.log 2>&1
tail -2 gpurun_out/r2G_pytest_multi.log | cut -c1-300
bash tests/tools/gpu_multi_bench.sh 2 r2G_bench_jelly2M_2gpu "--no-cpu --steps 200" r2G_bench_dam16M_2gpu "--scene dam_break --scale 1 --no-cpu --no-e2e --steps 30"
