#!/bin/bash
# round-2 GPU batch: full test suite per file, potrf phase profile, scorer / aggregation timings, ncu captures
mkdir -p gpurun_out
for f in tests/test_projection_gpu.py tests/test_composite_abi_gpu.py tests/test_scorer_gpu.py tests/test_aggregation_gpu.py tests/test_edge_cases_gpu.py tests/test_example_gpu.py tests/test_ridge_gpu.py; do
  echo "=== $f"; python -m pytest $f -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -12
done > gpurun_out/r2_tests.log 2>&1
tail -60 gpurun_out/r2_tests.log | grep -E "===|passed|failed|Error|error" 
echo "=== potrf"; python tools/bench_potrf.py
GADM_LIBRARY=$PWD/group-attribution-for-diffusion-models_b200/csrc/libgadm_potrfprof.so python tools/bench_potrf.py 2>&1 | grep "potrf cycles" | tail -3
echo "=== scorer"; python tools/bench_scorer.py 2>&1 | tail -3
echo "=== aggregation"; python tools/bench_aggregation.py | cut -c1-900
K=200000 bash tools/ncu_aggregation.sh 2>&1 | grep -E "mask_xty|ridge_gcv|lds_spearman|dgemm_dk" | tail -8
CMD="python tools/bench_aggregation.py --K 20000"
$CMD > gpurun_out/plain_agg20k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mask_xty_kernel -s 1 -c 1 -o gpurun_out/prof_xty $CMD > gpurun_out/ncu_xty.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ridge_gcv_score_kernel -c 1 -o gpurun_out/prof_gcv $CMD > gpurun_out/ncu_gcv.log 2>&1
tail -2 gpurun_out/ncu_xty.log gpurun_out/ncu_gcv.log
