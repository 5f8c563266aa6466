#!/bin/bash
# Launch list (kernel durations) of one TRAK scoring pass at the given shape (run under gpurun).
set -u
mkdir -p gpurun_out
CMD="python tools/bench_scorer.py --totals-only ${ARGS:---n 50000 --k 4096 --t 1000}"
timeout 300 $CMD > gpurun_out/plain_scorer_totals.log 2>&1 || { tail -3 gpurun_out/plain_scorer_totals.log; exit 1; }
tail -n 1 gpurun_out/plain_scorer_totals.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_scorer.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_scorer.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki].split('(')[0][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[vi].replace(',',''))/1e6
for k,(n,t) in agg.items(): print(f"{k:62s} n={n:4d} total={t:10.3f} ms  avg={t/n*1e3:9.1f} us")
PY
