#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_scorer_gpu.py tests/test_ridge_gpu.py tests/test_aggregation_gpu.py tests/test_composite_abi_gpu.py; do
  echo "=== $f"; python -m pytest $f -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -8
done
echo "=== potrf"; python tools/bench_potrf.py
GADM_LIBRARY=$PWD/group-attribution-for-diffusion-models_b200/csrc/libgadm_potrfprof.so python tools/bench_potrf.py 2>&1 | grep "potrf cycles" | tail -2
echo "=== scorer"; python tools/bench_scorer.py 2>&1 | tail -1
python tools/bench_scorer.py --n 5000 --k 32768 --t 50 --totals-only 2>&1 | tail -1
echo "=== aggregation"; python tools/bench_aggregation.py | cut -c1-900
K=200000 bash tools/ncu_aggregation.sh 2>&1 | grep -E "mask_xty|ridge_gcv|lds_spearman|dgemm_dk" | tail -6
