#!/bin/bash
# round-2 experiment 8: TMA bulk-copy stream in the wide staging kernel
L=$PWD/group-attribution-for-diffusion-models_b200/csrc
timeout 600 python -m pytest tests/test_projection_gpu.py -x -q -k "staging or accumulation or overlap" 2>&1 | tail -n 3
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_projection_gpu.py -x -q -k "staging_is_bit_identical" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head -n 5
python tools/bench_staging.py | tail -n 1
python tools/bench_staging.py --aligned | tail -n 1
GADM_LIBRARY=$L/libgadm_pround.so python tools/bench_staging.py | tail -n 1
timeout 900 python -m pytest tests/test_projection_gpu.py tests/test_composite_abi_gpu.py tests/test_example_gpu.py -x -q 2>&1 | tail -n 2
