#!/bin/bash
# DRAM traffic + tensor-pipe activity of the projection kernel at the FULL bench shape (few metrics -> few passes).
set -u
mkdir -p gpurun_out
export GADM_WATCHDOG_SEC=0
export GADM_PROJ_COOPERATIVE=0   # ncu kernel replay does not support cooperative launches
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second"
for t in normal rademacher; do
  CMD="python tools/bench_projection.py --type $t --k 4096 --iters 1"
  timeout 200 $CMD > gpurun_out/plain_full_$t.log 2>&1 && \
  timeout 600 ncu --metrics $M --clock-control none -k "regex:^project_(quad_)?kernel" -s 1 -c 1 --csv \
      --log-file gpurun_out/traffic_full_$t.csv $CMD > gpurun_out/ncu_full_$t.log 2>&1
  tail -n 8 gpurun_out/traffic_full_$t.csv | cut -d, -f5,13- | cut -c1-200
done
