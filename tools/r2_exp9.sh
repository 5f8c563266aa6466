#!/bin/bash
L=$PWD/group-attribution-for-diffusion-models_b200/csrc
timeout 600 python -m pytest tests/test_projection_gpu.py -x -q -k "staging or accumulation or overlap" 2>&1 | tail -n 2
python tools/bench_staging.py | tail -n 1
python tools/bench_staging.py --aligned | tail -n 1
