#!/bin/bash
# Round-2 evidence (run under gpurun): plain bench lines, ncu launch list of the bench command, DRAM traffic of the
# projection kernels at full size, ncu --set full of the projection (small D) and of the staging kernel.
# GADM_WATCHDOG_SEC=0: instrumented replays stretch in-kernel barrier waits beyond the watchdog.
# GADM_PROJ_COOPERATIVE=0 + GADM_PROJ_UNSAFE_LOCKSTEP=1: ncu cannot replay cooperative launches; under ncu every kernel
# runs alone, so all clusters are resident and the lockstep stays on.
set -u
mkdir -p gpurun_out
export GADM_WATCHDOG_SEC=0 GADM_PROJ_COOPERATIVE=0 GADM_PROJ_UNSAFE_LOCKSTEP=1
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline"
timeout 400 $CMD > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
tail -n 2 gpurun_out/r02_launches_bench.csv | cut -c1-250
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second"
for t in normal rademacher; do
  CMD="python tools/bench_projection.py --type $t --k 4096 --iters 1"
  timeout 200 $CMD > gpurun_out/plain_full_$t.log 2>&1 && \
  timeout 600 ncu --metrics $M --clock-control none -k "regex:^project_(quad_)?kernel" -s 1 -c 1 --csv \
      --log-file gpurun_out/r02_traffic_full_$t.csv $CMD > gpurun_out/ncu_full_$t.log 2>&1
  tail -n 8 gpurun_out/r02_traffic_full_$t.csv | cut -d, -f5,13- | cut -c1-200
done
CMD="python tools/bench_projection.py --type normal --k 4096 --D 4468288 --iters 1"
timeout 200 $CMD > gpurun_out/plain_normal.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^project_(quad_)?kernel" -s 1 -c 1 \
    -o gpurun_out/r02_prof_proj_normal $CMD > gpurun_out/ncu_normal.log 2>&1
tail -n 2 gpurun_out/ncu_normal.log
CMD="python tools/bench_staging.py --D 4468288"
timeout 200 $CMD > gpurun_out/plain_staging.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:stage_groups_wide_kernel" -s 4 -c 1 \
    -o gpurun_out/r02_prof_stage $CMD > gpurun_out/ncu_stage.log 2>&1
tail -n 2 gpurun_out/ncu_stage.log
MS="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct"
CMD="python tools/bench_staging.py"
timeout 300 ncu --metrics $MS --clock-control none -k "regex:stage_groups_wide_kernel" -s 40 -c 4 --csv \
    --log-file gpurun_out/r02_traffic_stage.csv $CMD > gpurun_out/ncu_stage2.log 2>&1
tail -n 6 gpurun_out/r02_traffic_stage.csv | cut -d, -f5,13- | cut -c1-200
# Rademacher path: narrow (co-resident) staging kernel
timeout 300 ncu --metrics $MS --clock-control none -k "regex:stage_groups_kernel" -s 40 -c 4 --csv \
    --log-file gpurun_out/r02_traffic_stage_narrow.csv python tools/bench_staging.py --type rademacher > gpurun_out/ncu_stage3.log 2>&1
tail -n 6 gpurun_out/r02_traffic_stage_narrow.csv | cut -d, -f5,13- | cut -c1-200
