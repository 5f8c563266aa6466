#!/bin/bash
timeout 900 python -m pytest tests/test_scorer_gpu.py tests/test_aggregation_gpu.py tests/test_edge_cases_gpu.py tests/test_composite_abi_gpu.py -x -q 2>&1 | tail -n 3
python tools/bench_scorer.py | tail -n 1
python tools/bench_scorer.py | tail -n 1
GADM_CHOL_LOOKAHEAD=0 python tools/bench_scorer.py | tail -n 1
python tools/bench_scorer.py --n 5000 --k 32768 --t 50 | tail -n 1
python tools/bench_aggregation.py --K 200000 | tail -n 1 | cut -c1-700
