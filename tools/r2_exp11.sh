#!/bin/bash
timeout 600 python -m pytest tests/test_projection_gpu.py -x -q -k "staging or accumulation or overlap" 2>&1 | tail -n 2
python tools/bench_staging.py | tail -n 1
python tools/bench_staging.py --blocks 270 | tail -n 1
python tools/bench_staging.py --blocks 270 --coresident 1 | tail -n 1
python tools/bench_staging.py --type rademacher --blocks 270 | tail -n 1
