#!/bin/bash
# Round evidence (run under gpurun): ncu --set full captures of the dominant kernels + the launch list of bench.py.
# GADM_WATCHDOG_SEC=0: instrumented replays stretch in-kernel barrier waits beyond the watchdog.
set -u
mkdir -p gpurun_out
export GADM_WATCHDOG_SEC=0
export GADM_PROJ_COOPERATIVE=0   # ncu kernel replay does not support cooperative launches
for t in rademacher normal; do
  CMD="python tools/bench_projection.py --type $t --k 4096 --D 4468288 --iters 1"
  timeout 200 $CMD > gpurun_out/plain_$t.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^project_(quad_)?kernel" -s 1 -c 1 \
      -o gpurun_out/prof_proj_$t $CMD > gpurun_out/ncu_$t.log 2>&1
  tail -n 1 gpurun_out/plain_$t.log | cut -c1-200; tail -n 2 gpurun_out/ncu_$t.log
done
CMD="python tools/bench_scorer.py --n 20000 --k 2048 --t 256"
timeout 200 $CMD > gpurun_out/plain_scorer.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_3xtf32 -s 0 -c 1 \
    -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
tail -n 2 gpurun_out/plain_scorer.log | cut -c1-300; tail -n 2 gpurun_out/ncu_gemm.log
CMD="python bench.py --steps 2 --warmup 3"
timeout 400 $CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
tail -n 3 gpurun_out/launches_bench.csv | cut -c1-250
