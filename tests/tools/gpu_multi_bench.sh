#!/bin/bash
# usage: gpu_multi_bench.sh N tag "bench args" [tag "bench args" ...]   — bench.py under torchrun on N GPUs, one JSON per tag
N=$1; shift
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  tag=$1; args=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $args > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
  python -c "
import json;d=json.load(open('gpurun_out/${tag}.json'));print('$tag', 'N=', d['n_gpus'], d['config']['particles_total'], 'particles, ms/step', round(d['ms_per_step'],4), 'G/s', round(d['value']/1e9,3), 'e2e', round(d.get('e2e',{}).get('value',0)/1e9,3), d['scaling'], 'parity', (d.get('slab_parity') or {}).get('within_tolerance'), (d.get('slab_parity') or {}).get('integer_fields_exact')); print('   stages by rank', d['roofline'].get('stage_ms_per_substep_by_rank'))" || (grep -n "Error\|error" gpurun_out/${tag}.err | head -5; tail -3 gpurun_out/${tag}.err)
done
