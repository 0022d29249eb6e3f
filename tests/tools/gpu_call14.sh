#!/bin/bash
mkdir -p gpurun_out
export SVB_MIGRATE_BESIDE_G2P=1
for cfg in "SVB_PDL=0" "SVB_SENDER_PRIORITY_NORMAL=1" "CUDA_DEVICE_MAX_CONNECTIONS=32"; do
  env $cfg timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/slab_worker.py jelly_shear 30 > gpurun_out/r2r_exp.log 2>&1
  echo "[$cfg] rc=$? $(grep -c 'within tolerance' gpurun_out/r2r_exp.log) $(grep -o 'FatalError: .\{0,80\}' gpurun_out/r2r_exp.log | head -1)"
done
