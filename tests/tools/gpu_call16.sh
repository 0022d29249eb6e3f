#!/bin/bash
# 1 GPU: the weak-scaling scene of 2 / 4 GPUs on ONE GPU (what does a slab rank's share cost without slabs?)
mkdir -p gpurun_out
for L in 1 2 4; do
  timeout 250 python bench.py --no-cpu --no-e2e --length $L --steps 200 > gpurun_out/r2t_len$L.json 2> gpurun_out/r2t.err
  python -c "
import json;d=json.load(open('gpurun_out/r2t_len$L.json'));r=d['roofline'];print('len=$L n', d['config']['particles_total'], 'ms/step', round(d['ms_per_step'],4), 'stages', r.get('stage_ms_per_substep'))" || tail -5 gpurun_out/r2t.err
done
