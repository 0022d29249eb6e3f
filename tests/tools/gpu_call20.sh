#!/bin/bash
# 1 GPU: P2G / G2P of the jelly collision before contact (substep 10) and deep in the collision (substep 215), ncu --set full
mkdir -p gpurun_out
for at in 10 215; do
  timeout 280 ncu --set full --clock-control none --import-source on -k regex:'k_p2g|k_g2p' --launch-skip $((2*at)) --launch-count 2 -f -o gpurun_out/r2v_jelly1M_at$at python tests/tools/ncu_target.py jelly_collision 220 1.0 > gpurun_out/r2v_ncu_$at.log 2>&1
  tail -2 gpurun_out/r2v_ncu_$at.log
done
