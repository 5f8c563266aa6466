#!/bin/bash
# round-2 experiment 3: fewer quad clusters (more SMs for the wide staging CTAs)
for q in 33 32 31 30; do
  echo "== quad clusters $q f16"; GADM_QUAD_CLUSTERS=$q python tools/bench_staging.py --coresident 0 2>&1 | tail -n 1
done
for q in 33 32; do
  echo "== quad clusters $q bf16"; GADM_QUAD_CLUSTERS=$q python tools/bench_staging.py --coresident 0 --stage-dtype bf16 2>&1 | tail -n 1
done
B="python bench.py --steps 6 --warmup 3 --no-extra --no-e2e --no-cpu-baseline"
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({"value":round(d["value"],1),"ms_per_step":round(d["ms_per_step"],1),"kernel_ms_in_situ":round(d["roofline"]["kernel_ms"],1),"kernel_alone_ms":round(d["extra"]["kernel_only"]["ms_per_pass"],1),"clk":d["clocks"]["sm_mhz"]}))'
for q in 33 32 31; do echo "== bench quad clusters $q"; GADM_QUAD_CLUSTERS=$q $B 2>/dev/null | python -c "$pick"; done
echo "== bench quad clusters 32 bf16"; GADM_STAGE_DTYPE=bf16 GADM_QUAD_CLUSTERS=32 $B 2>/dev/null | python -c "$pick"
