#!/bin/bash
# round-2 experiment 1: staging overlap and staging-format A/B (one GPU)
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-extra --no-e2e --no-cpu-baseline"
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({"value":round(d["value"],1),"ms_per_step":round(d["ms_per_step"],1),"kernel_ms_in_situ":round(d["roofline"]["kernel_ms"],1),"kernel_alone_ms":round(d["extra"]["kernel_only"]["ms_per_pass"],1),"rad_alone_ms":round(d["extra"]["projection_other_type"]["ms_per_step"],1),"clk":d["clocks"]["sm_mhz"],"staging":d["pipeline"]["staging"][:40]}))'
echo "== default"; $B 2>gpurun_out/e1_a.err | python -c "$pick"
echo "== bf16 staging"; GADM_STAGE_DTYPE=bf16 $B 2>gpurun_out/e1_b.err | python -c "$pick"
echo "== 64-register projection kernels"; GADM_LIBRARY=$PWD/group-attribution-for-diffusion-models_b200/csrc/libgadm_r64.so $B 2>gpurun_out/e1_c.err | python -c "$pick"
echo "== serial"; $B --no-overlap 2>gpurun_out/e1_d.err | python -c "$pick"
tail -n 3 gpurun_out/e1_*.err
