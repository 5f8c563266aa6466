#!/bin/bash
# Accuracy / throughput of the projection as a function of the TMEM accumulation segment length (C2 shape)
mkdir -p gpurun_out
for t in normal rademacher; do
  for seg in ${SEGS:-100000000 1024 512 256 128}; do
    GADM_PROJ_SEG_KB=$seg timeout 300 python tools/bench_projection.py --type $t --iters 3 --check 4 | tee -a gpurun_out/seg_sweep.jsonl | cut -c1-700
  done
done
