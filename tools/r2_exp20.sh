#!/bin/bash
B="python bench.py --steps 10 --warmup 3 --no-extra --no-e2e --no-cpu-baseline --no-producer --proj-type rademacher"
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({"value":round(d["value"],1),"ms_per_step":round(d["ms_per_step"],1),"kernel_ms_in_situ":round(d["roofline"]["kernel_ms"],1),"kernel_alone_ms":round(d["extra"]["kernel_only"]["ms_per_pass"],1),"clk":d["clocks"]["sm_mhz"]}))'
echo "== rademacher wide (default)"; $B 2>/dev/null | python -c "$pick"
echo "== rademacher narrow co-resident"; GADM_STAGE_NARROW=1 $B 2>/dev/null | python -c "$pick"
echo "== rademacher serial"; $B --no-overlap 2>/dev/null | python -c "$pick"
echo "== rademacher wide bf16"; GADM_STAGE_DTYPE=bf16 $B 2>/dev/null | python -c "$pick"
