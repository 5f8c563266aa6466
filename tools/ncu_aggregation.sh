#!/bin/bash
# ncu evidence for the aggregation kernels (run under gpurun): HBM GB/s and fp64-pipe use of the streaming
# contractions at a bandwidth-meaningful K (config 5 itself is 12 MB and latency-bound), plus the ridge GCV kernel.
set -u
mkdir -p gpurun_out
K=${K:-200000}
CMD="python tools/bench_aggregation.py --K $K"
timeout 300 $CMD > gpurun_out/plain_aggregation.log 2>&1 || { tail -5 gpurun_out/plain_aggregation.log; exit 1; }
tail -n 1 gpurun_out/plain_aggregation.log | cut -c1-900
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 900 ncu --metrics $M --clock-control none -k "regex:mask_xty_kernel|dgemm_dk_kernel|lds_spearman|ridge_gcv_score_kernel|mask_gram_kernel|sym_pinv_kernel" \
    -c 40 --csv --log-file gpurun_out/ncu_aggregation.csv $CMD > gpurun_out/ncu_aggregation.log 2>&1
tail -n 3 gpurun_out/ncu_aggregation.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/ncu_aggregation.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[ii], r[ki].split("(")[0]), {})[r[mi]] = r[vi]
for (i, k), m in per.items():
    t = float(m.get("gpu__time_duration.sum", "0").replace(",", "")) / 1e3
    rd = float(m.get("dram__bytes_read.sum", "0").replace(",", "")); wr = float(m.get("dram__bytes_write.sum", "0").replace(",", ""))
    print(f"{i:>3} {k:40s} {t:10.1f} us  dram {(rd + wr) / 1e6:9.1f} MB  {(rd + wr) / max(t, 1e-9) / 1e3:8.1f} GB/s  "
          f"fp64 pipe {m.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', '?')} %  issue {m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', '?')} %")
PY
