#!/bin/bash
# round-2 GPU batch 3: full test suite per file, smoke, the default bench (both staging formats) and the staging microbenchmark
mkdir -p gpurun_out
for f in tests/test_projection_gpu.py tests/test_composite_abi_gpu.py tests/test_scorer_gpu.py tests/test_aggregation_gpu.py tests/test_edge_cases_gpu.py tests/test_example_gpu.py tests/test_ridge_gpu.py; do
  echo "=== $f"; python -m pytest $f -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -n 12
done > gpurun_out/r2_tests3.log 2>&1
grep -E "===|passed|failed|Error|error" gpurun_out/r2_tests3.log | tail -n 30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -n 2 gpurun_out/r02_bench_n1.err; cut -c1-1500 gpurun_out/r02_bench_n1.json
GADM_STAGE_DTYPE=bf16 python bench.py --no-producer --no-cpu-baseline > gpurun_out/r02_bench_n1_bf16.json 2>/dev/null; cut -c1-600 gpurun_out/r02_bench_n1_bf16.json
python bench.py --steps 20 --no-producer --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r02_bench_n1_steps20.json 2>/dev/null; cut -c1-600 gpurun_out/r02_bench_n1_steps20.json
python tools/bench_staging.py | tail -n 1
python tools/bench_staging.py --type rademacher | tail -n 1
python tools/bench_aggregation.py | cut -c1-1200
