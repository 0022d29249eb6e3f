#!/bin/bash
# 2 GPUs: slab parity (incl. adaptive steps and a failing particle) + the weak-scaling bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_slabs.py -m gpu -q -s > gpurun_out/r2e_pytest_slabs.log 2>&1
tail -12 gpurun_out/r2e_pytest_slabs.log | cut -c1-250
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu > gpurun_out/r2e_bench_jelly2M_2gpu.json 2> gpurun_out/r2e_bench2.err
python -c "
import json;d=json.load(open('gpurun_out/r2e_bench_jelly2M_2gpu.json'));print('2gpu', d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],4), d['value']/1e9, d['e2e']['value']/1e9, d['slab_parity']); print(d['roofline']['stage_ms_per_substep_by_rank'])" || tail -5 gpurun_out/r2e_bench2.err
