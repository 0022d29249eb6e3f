#!/bin/bash
mkdir -p gpurun_out
export SVB_MIGRATE_BESIDE_G2P=1
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/slab_worker.py jelly 30 > gpurun_out/r2q_ms_beside.log 2>&1
echo rc=$?
grep -n "FatalError\|within tolerance\|FAILED" gpurun_out/r2q_ms_beside.log | cut -c1-400 | head
