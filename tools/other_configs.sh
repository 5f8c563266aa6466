#!/bin/bash
# Projection at the other BASELINE shapes (parity-test configs, for the record; not bench lines), with the fp64 check
# C1: D=35.7M k=2048; C3: D=274 056 163 k=8192 (rows limited by HBM); C4: D=51 019 776 k=32768
for spec in "35746307 2048 0 normal" "35746307 2048 0 rademacher" "51019776 32768 0 normal" "51019776 32768 0 rademacher" "274056163 8192 256 normal" "274056163 8192 256 rademacher"; do
  set -- $spec
  echo "D=$1 k=$2 M=$3 type=$4"
  timeout 300 python tools/bench_projection.py --D $1 --k $2 --M $3 --type $4 --iters 2 --check 2 | tee -a gpurun_out/other_configs.jsonl | cut -c1-520
done
