#!/bin/bash
# N GPUs (default 4): weak-scaling bench line with exchange wait accounting
N=${1:-4}
mkdir -p gpurun_out
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu --steps 200 > gpurun_out/r2u_bench_jelly_${N}gpu.json 2> gpurun_out/r2u_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2u_bench_jelly_${N}gpu.json'));print('N=$N ms/step', round(d['ms_per_step'],4), d['value']/1e9, d['e2e']['value']/1e9, d['slab_parity']['within_tolerance'])
for a,b in zip(d['roofline']['stage_ms_per_substep_by_rank'], d['roofline']['exchange_wait_ms_per_substep_by_rank']): print({k:round(v*1e3) for k,v in a.items()}, {k[:14]:(round(v*1e3) if v<10 else v) for k,v in b.items()})" || tail -5 gpurun_out/r2u_bench.err
