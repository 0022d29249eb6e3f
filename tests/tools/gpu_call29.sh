#!/bin/bash
# 1 GPU: streaming (evict-first) loads for rows that are dead after the read — G2P's source row (gcs), P2G's v / C quads (pcs), both (bcs)
mkdir -p gpurun_out
bash tests/tools/ab1.sh cur gcs pcs bcs 2>&1 | tee gpurun_out/r2F_ab.txt
