#!/bin/bash
# 1 GPU, final build: compute-sanitizer memcheck and racecheck over every code path on small scenes
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 110 compute-sanitizer --tool $tool --error-exitcode 3 python tests/tools/sanitize_target.py > gpurun_out/r2G_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r2G_sanitizer_$tool.log | cut -c1-200
done
