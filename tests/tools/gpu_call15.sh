#!/bin/bash
# 2 GPUs: migration sender beside G2P (kernels preloaded) — parity scenes, then A/B of the weak-scaling bench line
mkdir -p gpurun_out
for sc in "jelly_shear 30" "jelly 30" "sand 40" "jelly_rebalance 30"; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/slab_worker.py $sc > gpurun_out/r2s_one.log 2>&1
  echo "[$sc] rc=$? ok=$(grep -c 'within tolerance' gpurun_out/r2s_one.log) $(grep -o 'FatalError: .\{0,200\}' gpurun_out/r2s_one.log | head -1)"
done
for ab in 1 0; do
  SVB_MIGRATE_BESIDE_G2P=$ab timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu --steps 200 > gpurun_out/r2s_bench_2gpu_beside$ab.json 2> gpurun_out/r2s_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/r2s_bench_2gpu_beside$ab.json'));print('beside=$ab ms/step', round(d['ms_per_step'],4), d['value']/1e9, d['e2e']['value']/1e9, d['slab_parity']['within_tolerance'])" || grep -n "Error" gpurun_out/r2s_bench.err | head -5
done
