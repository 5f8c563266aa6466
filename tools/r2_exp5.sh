#!/bin/bash
timeout 900 python -m pytest tests/test_projection_gpu.py -x -q 2>&1 | tail -n 2
for w in 0 1; do echo "== wait_resident $w"; python tools/bench_staging.py --wait-resident $w 2>&1 | tail -n 1; done
echo "== rademacher wait 1"; python tools/bench_staging.py --type rademacher --wait-resident 1 2>&1 | tail -n 1
B="python bench.py --steps 6 --warmup 3 --no-extra --no-e2e --no-cpu-baseline"
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({"value":round(d["value"],1),"ms_per_step":round(d["ms_per_step"],1),"kernel_ms_in_situ":round(d["roofline"]["kernel_ms"],1),"kernel_alone_ms":round(d["extra"]["kernel_only"]["ms_per_pass"],1),"rad":d["extra"]["projection_other_type"].get("value"),"clk":d["clocks"]["sm_mhz"]}))'
echo "== bench f16"; $B 2>/dev/null | python -c "$pick"
echo "== bench bf16"; GADM_STAGE_DTYPE=bf16 $B 2>/dev/null | python -c "$pick"
