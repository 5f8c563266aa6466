#!/bin/bash
# ncu evidence for the projection kernel (run under gpurun).  Smaller D than the bench shape so that the ~40
# replay passes of `--set full` stay short; the steady-state pipeline behaviour is the same (units are
# 256-column x D-split tiles either way).  GADM_WATCHDOG_SEC=0: instrumented replays stretch barrier waits.
set -u
mkdir -p gpurun_out
export GADM_WATCHDOG_SEC=0
export GADM_PROJ_COOPERATIVE=0   # ncu kernel replay does not support cooperative launches
for t in rademacher normal; do
  CMD="python tools/bench_projection.py --type $t --k 4096 --D 4468288 --iters 1"
  timeout 200 $CMD > gpurun_out/plain_$t.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:project_kernel -s 1 -c 1 \
      -o gpurun_out/prof_proj_$t $CMD > gpurun_out/ncu_$t.log 2>&1
  tail -n 2 gpurun_out/plain_$t.log; tail -n 4 gpurun_out/ncu_$t.log
done
CMD="python tools/bench_projection.py --type normal --k 4096 --D 4468288 --iters 2"
timeout 200 $CMD > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_projection.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -n 12 gpurun_out/launches_projection.csv
