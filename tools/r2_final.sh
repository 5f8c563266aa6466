#!/bin/bash
# round-2 final evidence: all GPU tests, smoke, default bench line, ncu of the CTA-pair Gram
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -q -m gpu -p no:cacheprovider 2>&1 | tail -n 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -n 2 gpurun_out/r02_bench_n1.err; cut -c1-400 gpurun_out/r02_bench_n1.json
python tools/bench_scorer.py | tail -n 1 > gpurun_out/r02_bench_scorer.json; cut -c1-700 gpurun_out/r02_bench_scorer.json
python tools/check_gemm_2cta.py > gpurun_out/plain_gemm_2cta.log 2>&1 && \
GADM_WATCHDOG_SEC=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ts2_kernel -s 9 -c 1 -f -o gpurun_out/r02_prof_gemm_pair python tools/check_gemm_2cta.py > gpurun_out/ncu_ts2.log 2>&1
tail -n 2 gpurun_out/ncu_ts2.log | cut -c1-200
bash tools/ncu_scorer_launches.sh 2>&1 | tail -n 16
