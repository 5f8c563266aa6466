#!/bin/bash
L=$PWD/group-attribution-for-diffusion-models_b200/csrc
for g in 0 1; do echo "== gate $g"; python tools/bench_staging.py --gate $g | tail -n 1; done
echo "== rademacher gate 1"; python tools/bench_staging.py --type rademacher --gate 1 | tail -n 1
B="python bench.py --steps 8 --warmup 3 --no-extra --no-e2e --no-cpu-baseline --no-producer"
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({"value":round(d["value"],1),"ms_per_step":round(d["ms_per_step"],1),"kernel_ms_in_situ":round(d["roofline"]["kernel_ms"],1),"kernel_alone_ms":round(d["extra"]["kernel_only"]["ms_per_pass"],1),"clk":d["clocks"]["sm_mhz"]}))'
echo "== bench gate 1"; $B 2>/dev/null | python -c "$pick"
echo "== bench gate 0"; GADM_PIPE_GATE=0 $B 2>/dev/null | python -c "$pick"
echo "== bench gate 1 pround"; GADM_LIBRARY=$L/libgadm_pround.so $B 2>/dev/null | python -c "$pick"
echo "== bench rademacher gate 1"; $B --proj-type rademacher 2>/dev/null | python -c "$pick"
echo "== bench rademacher gate 0"; GADM_PIPE_GATE=0 $B --proj-type rademacher 2>/dev/null | python -c "$pick"
