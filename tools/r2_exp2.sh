#!/bin/bash
# round-2 experiment 2: wide staging CTAs confined to the SMs the quad grid strands; P low-mantissa-bit power test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_projection_gpu.py tests/test_composite_abi_gpu.py tests/test_example_gpu.py -x -q 2>&1 | tail -n 5
B="python bench.py --steps 3 --warmup 3 --no-extra --no-e2e --no-cpu-baseline"
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({"value":round(d["value"],1),"ms_per_step":round(d["ms_per_step"],1),"kernel_ms_in_situ":round(d["roofline"]["kernel_ms"],1),"kernel_alone_ms":round(d["extra"]["kernel_only"]["ms_per_pass"],1),"rad_api":d["extra"]["projection_other_type"],"clk":d["clocks"]["sm_mhz"],"staging":d["pipeline"]["staging"][:40]}))'
echo "== default"; $B 2>gpurun_out/e2_a.err | python -c "$pick"
echo "== bf16 staging"; GADM_STAGE_DTYPE=bf16 $B 2>gpurun_out/e2_b.err | python -c "$pick"
echo "== P mask"; GADM_LIBRARY=$PWD/group-attribution-for-diffusion-models_b200/csrc/libgadm_pmask.so $B 2>gpurun_out/e2_c.err | python -c "$pick"
echo "== serial"; $B --no-overlap 2>gpurun_out/e2_d.err | python -c "$pick"
for c in 0 1; do python tools/bench_staging.py --coresident $c 2>&1 | tail -n 4; done
python tools/bench_staging.py --type rademacher --coresident 1 2>&1 | tail -n 4
tail -n 3 gpurun_out/e2_*.err
