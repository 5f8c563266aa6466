#!/bin/bash
# perf sweep of the projection kernel with SM clock / power sampled during each run
mkdir -p gpurun_out
for t in rademacher normal; do for gw in 2 4; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 100 > gpurun_out/clk_${t}_${gw}.csv &
  SMI=$!
  GADM_PROJ_GEN_WARPS=$gw timeout 120 python tools/bench_projection.py --type $t --k 4096 --iters 5 | cut -c1-250
  kill $SMI
  python - <<PY
import statistics
rows=[l.split(',') for l in open('gpurun_out/clk_${t}_${gw}.csv') if ',' in l]
clk=[float(r[0]) for r in rows]; pw=[float(r[1]) for r in rows]
busy=[(c,p) for c,p in zip(clk,pw) if p>600]
print('  $t gw=$gw samples',len(rows),'busy',len(busy),'clk median under load',statistics.median([c for c,_ in busy]) if busy else None,'power max',max(pw))
PY
done; done
