#!/bin/bash
# round-2 experiment 7: staging grid order; P rounded to 8 significant bits (power)
L=$PWD/group-attribution-for-diffusion-models_b200/csrc
echo "== rowfast 0"; python tools/bench_staging.py | tail -n 1
echo "== rowfast 1"; GADM_STAGE_ROWFAST=1 python tools/bench_staging.py | tail -n 1
for v in "" _pround _pmask; do
  echo "== lib$v"; GADM_LIBRARY=$L/libgadm$v.so python tools/bench_staging.py | tail -n 1
done
B="python bench.py --steps 8 --warmup 3 --no-extra --no-e2e --no-cpu-baseline --no-producer"
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({"value":round(d["value"],1),"ms_per_step":round(d["ms_per_step"],1),"kernel_ms_in_situ":round(d["roofline"]["kernel_ms"],1),"kernel_alone_ms":round(d["extra"]["kernel_only"]["ms_per_pass"],1),"clk":d["clocks"]["sm_mhz"]}))'
for v in "" _pround; do echo "== bench lib$v"; GADM_LIBRARY=$L/libgadm$v.so $B 2>/dev/null | python -c "$pick"; done
echo "== bench rowfast"; GADM_STAGE_ROWFAST=1 $B 2>/dev/null | python -c "$pick"
