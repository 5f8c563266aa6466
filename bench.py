"""Benchmark of the attribution hot path (driver contract: one JSON line on rank 0).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU projector (trak BasicProjector restatement)

Workload (BASELINE.json configs[1], "CIFAR-10 DDPM TRAK ... proj_dim 4096"): one *step* = one pass of the JL
projection over one staged batch of 1024 (normal type; 512 for Rademacher) per-example gradients of the
35 746 307-parameter DDPM-CIFAR U-Net (synthetic bf16 rows resident in HBM) -> 1024 x 4096 features.  Metric =
projected gradients per second, whole job (sum over ranks, weak scaling: every rank projects its own batch per step, no collective on
the projection path).  `roofline` = 2*M*D*k flops per launch / CUDA-event time against the measured dense
bf16 peak; `e2e` = the same metric through the public `CudaProjector.deferred()` API with fp32 gradients
coming from pinned host memory and the features read back to the host inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRAD_DIM = 35_746_307  # DDPM-CIFAR U-Net (src/ddpm_config.py:48-82), SURVEY.md section 8
PROJ_DIM = 4096
STAGE_ROWS = 512  # rows per pass of the pair kernel (Rademacher); the normal type stages 1024 (quad kernel)
N_TRAIN, N_GEN = 50_000, 1_000
METRIC = "projected_grads_per_sec"
UNIT = "grads/s"


def bench_config(proj_type: str, rows: int, world: int):
    """The workload both arms are measured on (the reference arm prints the same dict)."""
    return {"workload": "CIFAR-10 DDPM TRAK featurisation (BASELINE configs[1]): JL projection of per-example U-Net "
                        f"gradients, fp32 [32, D] batches -> [{rows}, 4096] features per step per GPU",
            "grad_dim": GRAD_DIM, "proj_dim": PROJ_DIM, "proj_type": proj_type, "rows_per_step_per_gpu": rows,
            "sharding": f"examples x{world}",
            "l2": f"staged input ({rows * GRAD_DIM * 2 / 1e9:.1f} GB) and split-K partials exceed the 126 MB L2"}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p.get("bf16_tflops", 1590.0), "bf16_sustained": p.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": p.get("hbm_gbs", 6650.0), "source": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def _cpu_projector_rate(seconds_target: float = 8.0, proj_type: str = "normal"):
    """grads/s of trak BasicProjector (oracle restatement) on the host cores, on a bounded sample."""
    import torch

    from oracle.projector import BasicProjectorOracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = 8  # the reference's batch (d_trak_grad.py:327)
    # calibrate on 1/64 of D, then pick the slice of D that costs ~seconds_target for ONE 100-column block
    d_small = GRAD_DIM // 64
    g = torch.randn(batch, d_small)
    p = BasicProjectorOracle(d_small, PROJ_DIM, 42, proj_type)
    t0 = time.perf_counter(); p.project(g, 0, blocks=1); t_small = time.perf_counter() - t0
    frac = min(1.0, max(1.0 / 64, seconds_target / (t_small * 64)))
    d_s = int(GRAD_DIM * frac)
    g = torch.randn(batch, d_s)
    p = BasicProjectorOracle(d_s, PROJ_DIM, 42, proj_type)
    t0 = time.perf_counter(); p.project(g, 0, blocks=1); t_blk = time.perf_counter() - t0
    blocks = -(-PROJ_DIM // 100)
    t_full_batch = t_blk * (GRAD_DIM / d_s) * blocks  # linear in D and in the number of 100-column blocks
    sample = (f"BasicProjector({proj_type}) 1 of {blocks} column blocks x [8, {d_s}] fp32 (D/{GRAD_DIM / d_s:.1f}) "
              f"= {t_blk:.2f} s, extrapolated linearly to D={GRAD_DIM}, k={PROJ_DIM}")
    return batch / t_full_batch, cores, sample, t_blk


def run_reference(args):
    """--impl reference: the reference's CPU projection path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    sample = ""
    cores = os.cpu_count() or 1
    t_begin = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, cores, sample, t_blk = _cpu_projector_rate(seconds_target=3.0, proj_type=args.proj_type)
        if i >= args.warmup:
            vals.append(v)
        if time.perf_counter() - t_begin > 150 and vals:
            break
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": 1e3 * 1024 / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.proj_type, 1024 if args.proj_type == "normal" else STAGE_ROWS, max(1, args.gpus)),
        "reference_batch": 8,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gadm", choices=["gadm", "reference"])
    ap.add_argument("--proj-type", default="normal", choices=["normal", "rademacher"],
                    help="normal is what the reference instantiates (d_trak_grad.py:508)")
    ap.add_argument("--no-extra", action="store_true", help="skip the scoring / aggregation side measurements")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="stage and project serially on one stream")
    ap.add_argument("--no-producer", action="store_true", help="skip the U-Net vmap(grad) end-to-end leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import gadm_b200
    from gadm_b200 import CudaProjector, ProjectionType

    gadm_b200.load_library()  # fails loudly if the CUDA extension is missing
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = None
    if world > 1:  # several ranks stream host->device at once in the e2e leg: keep each rank's pinned buffers NUMA-local
        from gadm_b200.distributed import bind_to_gpu_numa

        numa_cpus = bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = _peaks()
    ptype = ProjectionType(args.proj_type)
    rows = 1024 if args.proj_type == "normal" else STAGE_ROWS  # staged examples per step per GPU
    chunk = 32  # examples per add(): the gradients of one producer batch, [32, D] fp32 resident in HBM
    proj = CudaProjector(GRAD_DIM, PROJ_DIM, 42, ptype, dev, 32, stage_rows=rows)
    handle = proj._handle
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    src = torch.randn(chunk, GRAD_DIM, device=dev, generator=gen) * 1e-3  # synthetic per-example gradients
    overlap = not args.no_overlap
    if overlap:
        free, _ = torch.cuda.mem_get_info(dev)
        overlap = free > 2 * rows * proj.d_pad * 2 + (12 << 30)  # two staging buffers + workspace + the e2e leg's chunks

    # ---------------- value: device-resident fp32 gradients through the PUBLIC API (staging inside the timed region)
    sink = proj.deferred(model_id=0, overlap=overlap, record_events=True)

    def api_step():
        for _ in range(rows // chunk):
            sink.add(src)  # one staging launch per batch; the rows-th example triggers the projection pass

    for _ in range(args.warmup):
        api_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = handle.launch_count()
    n_warm_passes = len(sink.pass_events)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        api_step()
    feats = sink.result()  # joins the side stream: every pass of the timed steps has finished
    t1.record()
    barrier()
    launches = handle.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(t0.elapsed_time(t1))
    ms_per_step = total_ms / args.steps
    value = rows * world / (ms_per_step * 1e-3)
    pass_ms = [e0.elapsed_time(e1) for (e0, e1, _r) in sink.pass_events[n_warm_passes:]]
    assert len(pass_ms) == args.steps and feats.shape[0] == rows * (args.steps + args.warmup)
    flops_per_launch = 2.0 * rows * GRAD_DIM * PROJ_DIM
    kernel_ms = sum(pass_ms) / len(pass_ms)  # project kernel + its (<0.1 %) split-K reduce, in situ (staging overlapped)
    achieved = flops_per_launch / (kernel_ms * 1e-3) / 1e12
    wd = handle.watchdog_code()
    assert wd == 0, f"kernel watchdog fired: {wd:#x}"
    stage_launches_per_step = rows // chunk
    del feats, sink

    # ---------------- kernel alone on the rows just staged (no staging in flight): the round-1 `value`, for continuity
    stage0 = proj._stage(rows, 0)
    out = torch.empty(rows, PROJ_DIM, device=dev)
    for _ in range(2):
        proj._project_rows(stage0, rows, 0, out)
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(3):
        proj._project_rows(stage0, rows, 0, out)
    k1.record()
    barrier()
    alone_ms = max_over_ranks(k0.elapsed_time(k1)) / 3
    kernel_only = {"ms_per_pass": alone_ms, "value": rows * world / (alone_ms * 1e-3), "unit": UNIT,
                   "tflops_per_gpu": flops_per_launch / (alone_ms * 1e-3) / 1e12,
                   "what": "gadm_project_staged on pre-staged rows, nothing else on the GPU (round-1 headline definition)"}

    # the other projection type on the same staged rows (side number; the headline is --proj-type)
    other = "rademacher" if args.proj_type == "normal" else "normal"
    rows_o = min(rows, 1024 if other == "normal" else STAGE_ROWS)
    proj_o = CudaProjector(GRAD_DIM, PROJ_DIM, 42, ProjectionType(other), dev, 32, stage_rows=rows_o)
    proj_o._stages = proj._stages
    for _ in range(2):
        proj_o._project_rows(stage0, rows_o, 0, out)
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for _ in range(3):
        proj_o._project_rows(stage0, rows_o, 0, out)
    o1.record()
    barrier()
    other_ms = max_over_ranks(o0.elapsed_time(o1)) / 3
    flops_o = 2.0 * rows_o * GRAD_DIM * PROJ_DIM
    other_line = {"proj_type": other, "rows_per_step_per_gpu": rows_o, "ms_per_step": other_ms,
                  "value": rows_o * world / (other_ms * 1e-3), "unit": UNIT,
                  "tflops_per_gpu": flops_o / (other_ms * 1e-3) / 1e12,
                  "frac_of_burst_peak": flops_o / (other_ms * 1e-3) / 1e12 / peaks["bf16_burst"],
                  "frac_of_sustained_peak": flops_o / (other_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                  "what": "kernel alone on pre-staged rows"}
    proj_o._stages = []
    proj_o._ws = None
    del proj_o

    # ---------------- the unmodified reference script: project() per batch of 8 / 16 (d_trak_grad.py:327,776;
    # grad_text_to_image_lora.py:561-568) -- every call regenerates all of P for a handful of rows
    unmodified = {}
    if rank == 0:
        for bsz in (8, 16):
            proj.project(src[:bsz], 0)
            torch.cuda.synchronize()
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record()
            for _ in range(3):
                proj.project(src[:bsz], 0)
            u1.record()
            torch.cuda.synchronize()
            ms = u0.elapsed_time(u1) / 3
            unmodified[f"batch_{bsz}"] = {"ms_per_call": ms, "value": bsz / (ms * 1e-3), "unit": UNIT}
        unmodified["what"] = ("CudaProjector.project([B, D] fp32 on the device) per call, as the reference scripts call it "
                              "through shims/trak without deferred(): one GPU")
    barrier()

    # ---------------- end to end: pinned host fp32 gradients -> public API -> features back on the host
    e2e = None
    if not args.no_e2e:
        host = torch.empty(chunk, GRAD_DIM, dtype=torch.float32).pin_memory()
        host.normal_(0, 1e-3)
        host_out = torch.empty(rows, PROJ_DIM, dtype=torch.float32).pin_memory()
        bufs = [src, torch.empty(chunk, GRAD_DIM, dtype=torch.float32, device=dev)]
        copy_stream = torch.cuda.Stream(device=dev)
        free_ev = [torch.cuda.Event() for _ in range(2)]
        full_ev = [torch.cuda.Event() for _ in range(2)]
        main_stream = torch.cuda.current_stream(dev)

        def e2e_step():
            with proj.deferred(model_id=0, overlap=overlap) as sk:
                for c in range(rows // chunk):
                    b = c % 2
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(free_ev[b])
                        bufs[b].copy_(host, non_blocking=True)  # H2D of this chunk's gradients
                        full_ev[b].record(copy_stream)
                    main_stream.wait_event(full_ev[b])
                    sk.add(bufs[b])  # staged in one launch (projects when the last row of the pass is staged)
                    free_ev[b].record(main_stream)
            host_out.copy_(sk.result(), non_blocking=True)  # D2H of the step's result
            main_stream.synchronize()

        for e in free_ev:
            e.record(main_stream)
        n_e2e_warm = 1
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e_warm):
            e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_e2e):
            e2e_step()
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / n_e2e
        e2e = {"value": rows * world / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": rows * GRAD_DIM * 4, "d2h_bytes_per_step": rows * PROJ_DIM * 4,
               "ms_per_step": e2e_ms, "steps": n_e2e,
               "api": "CudaProjector.deferred().add(fp32 grads from pinned host) -> result() -> host",
               "numa_bound_cpus": len(numa_cpus) if numa_cpus else None}
        del bufs, host
    proj.free_memory()
    del stage0, src
    torch.cuda.empty_cache()

    # ---------------- side measurements: full TRAK scoring at C2 dims, aggregation at config 5, the real producer
    extra = {}
    if not args.no_extra:
        try:
            extra = side_measurements(dev, rank, world)
        except Exception as e:  # never lose the headline line
            extra = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_producer:
            try:
                extra["e2e_producer"] = producer_leg(dev, args.proj_type)
            except Exception as e:
                extra["e2e_producer"] = {"error": f"{type(e).__name__}: {e}"}
    extra["projection_other_type"] = other_line
    extra["kernel_only"] = kernel_only
    extra["unmodified_script"] = unmodified

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = _cpu_projector_rate(seconds_target=8.0, proj_type=args.proj_type)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        traffic = None
        summ = os.path.join(ROOT, "profiles", "projection_traffic.json")
        if os.path.exists(summ):
            with open(summ) as f:
                traffic = json.load(f).get(args.proj_type, {}).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 (group-scaled gradients, 11 bits) x f16 (P, 8 significant bits), fp32 accumulate" if proj.stage_dtype == "f16" else "bf16",
            "data": "synthetic",
            "config": bench_config(args.proj_type, rows, world),
            "pipeline": {"api": "CudaProjector.deferred().add(fp32 [32, D] on the device) -> result()",
                         "staging": f"{proj.stage_dtype} staging inside the timed region"
                                    + (", overlapped with the previous pass (second staging buffer, side stream)" if overlap else ", serial")},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_sustained"], "traffic": traffic,
                         "peak_kind": f"bf16_tflops_sustained of measured ({peaks['source']}); timed inside a multi-step loop",
                         "frac_of_burst_peak": achieved / peaks["bf16_burst"], "flops_per_launch": flops_per_launch,
                         "kernel_ms": kernel_ms,
                         "timing": "CUDA events around every pass on the stream it is launched on, staging of the next "
                                   "pass running concurrently" if overlap else "CUDA events around every pass",
                         "kernel": "gadm::proj::project_quad_kernel<4, 4>" if args.proj_type == "normal"
                         else "gadm::proj::project_kernel<2,2>"},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "gpu_launches_per_step": {"stage_groups_kernel": stage_launches_per_step, "project": 1, "project_reduce": 1},
            "clocks": clocks, "tflops": achieved * world, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def producer_leg(dev, proj_type):
    """SURVEY 8(f)-1, the real end to end: the reference's DDPM-CIFAR U-Net (35 746 307 parameters, random init)
    under vmap(grad) for 10 timesteps per example -> DeferredProjection.accumulate -> projection -> host.
    Reference loop: d_trak_grad.py:718-792.  Reports grads/s and how the wall time splits."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from ddpm_unet import DDPMCifarUNet, DDPMScheduler, count_parameters
    from featurize_and_score import featurize

    from gadm_b200 import CudaProjector, ProjectionType

    torch.manual_seed(0)
    model = DDPMCifarUNet().to(dev).eval()
    n_params = count_parameters(model)
    assert n_params == GRAD_DIM
    sched = DDPMScheduler(device=dev)
    K, batch = 10, 16
    g = torch.Generator(device=dev).manual_seed(1)
    images = torch.rand(1024, 3, 32, 32, device=dev, generator=g) * 2 - 1
    # producer alone on two batches: sizes the sample so that the leg stays near 20 s
    featurize(model, images[:batch], None, sched, K, 42, "loss", batch, project=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    featurize(model, images[:2 * batch], None, sched, K, 42, "loss", batch, project=False)
    torch.cuda.synchronize()
    prod_s_per_example = (time.perf_counter() - t0) / (2 * batch)
    n = int(min(1024, max(64, (18.0 / prod_s_per_example) // batch * batch)))
    proj = CudaProjector(n_params, PROJ_DIM, 42, ProjectionType(proj_type), dev, batch, stage_rows=min(1024, n))
    host_out = torch.empty(n, PROJ_DIM, dtype=torch.float32).pin_memory()
    sink = proj.deferred(0, record_events=True)
    featurize(model, images[:batch], proj, sched, K, 42, "loss", batch, sink=sink)  # warm-up: allocations, autotune
    sink.result()
    torch.cuda.synchronize()
    sink = proj.deferred(0, record_events=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = proj._handle.launch_count()
    e0.record()
    featurize(model, images[:n], proj, sched, K, 42, "loss", batch, sink=sink)
    host_out.copy_(sink.result(), non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    pass_ms = sum(a.elapsed_time(b) for (a, b, _r) in sink.pass_events)
    proj.free_memory()
    return {"value": n / (total_ms * 1e-3), "unit": UNIT, "examples": n, "timesteps": K, "batch": batch,
            "total_ms": total_ms, "producer_alone_ms": prod_s_per_example * n * 1e3, "projection_pass_ms": pass_ms,
            "projection_share_of_wall": pass_ms / total_ms,
            "staging_and_timestep_sum_share_of_wall": max(0.0, 1.0 - pass_ms / total_ms - prod_s_per_example * n * 1e3 / total_ms),
            "gadm_launches": int(proj._handle.launch_count() - l0), "finite": bool(torch.isfinite(host_out).all()),
            "d2h_bytes": n * PROJ_DIM * 4,
            "what": "DDPMCifarUNet vmap(grad) x 10 timesteps -> DeferredProjection.accumulate(dict of per-parameter "
                    "grads, 1/K) -> staged on the last timestep -> projection passes -> features on the host"}


def side_measurements(dev, rank, world):
    """Full TRAK score time (C2 dims, sharded by example) and Shapley/Banzhaf/LDS aggregation (config 5)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import gadm_b200 as G

    out = {}
    n_local = N_TRAIN // world
    g = torch.Generator(device=dev).manual_seed(rank)
    train = torch.randn(n_local, PROJ_DIM, device=dev, generator=g)
    g2 = torch.Generator(device=dev).manual_seed(10_000)
    gen = torch.randn(N_GEN, PROJ_DIM, device=dev, generator=g2)
    handle = G._lib.get_handle(dev) if hasattr(G, "_lib") else None
    for it in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = G.trak_scores(train, gen, lam=0.5, variants=("trak",), gather=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    gram_flops = 2.0 * N_TRAIN * PROJ_DIM * PROJ_DIM
    out["trak_score"] = {"ms": ms, "n_train": N_TRAIN, "n_gen": N_GEN, "proj_dim": PROJ_DIM, "lam": 0.5,
                         "what": "transpose + wave-balanced Gram (+NCCL all-reduce) -> Cholesky -> single-row substitution for "
                                 "mean_t(gen_t) K^-1 (one cooperative launch, no explicit inverse) -> matvec over the training "
                                 "features (+all-gather), incl. the factorisation check (one D2H)",
                         "gram_flops": gram_flops, "finite": bool(torch.isfinite(res["trak"]).all())}
    # parity of the sharded path (all-reduce of the Gram, all-gather of the score slices) against the unsharded
    # computation on rank 0, and of both against an fp64 product for a sample of training examples
    if world > 1:
        parts = [torch.empty_like(train) for _ in range(world)]
        dist.all_gather(parts, train)
        train_all = torch.cat(parts, dim=0)
        del parts
    else:
        train_all = train
    if rank == 0:
        single = G.trak_scores(train_all, gen, lam=0.5, variants=("trak",), group=G.LOCAL)["trak"]
        scale = float(single.abs().max())
        k_top = 100
        top_s = set(torch.argsort(single, descending=True, stable=True)[:k_top].tolist())
        top_d = set(torch.argsort(res["trak"], descending=True, stable=True)[:k_top].tolist())
        # fp64 spot check: s_n = mean(gen) K^-1 phi_n for 64 examples, K^-1 applied by fp64 Cholesky (torch, checker only)
        tp = train_all.double()
        K = tp.T @ tp
        K.diagonal().add_(0.5)
        zbar = torch.cholesky_solve(gen.double().mean(dim=0)[:, None], torch.linalg.cholesky(K))[:, 0]
        idx = torch.arange(0, N_TRAIN, N_TRAIN // 64, device=dev)
        want = tp[idx] @ zbar
        out["trak_score"].update({
            "rel_err_vs_single_gpu": float((res["trak"] - single).abs().max()) / scale,
            "topk_equal": top_s == top_d, "topk": k_top, "topk_overlap": len(top_s & top_d),
            "rel_err_vs_fp64_sample": float((res["trak"][idx].double() - want).abs().max()) / float(want.abs().max()),
            "single_gpu_rel_err_vs_fp64_sample": float((single[idx].double() - want).abs().max()) / float(want.abs().max())})
        del tp, K, single
    del train, gen, res, train_all
    torch.cuda.empty_cache()
    if rank == 0:
        n, d, K, m = 1000, 100, 1000, 100
        rng = np.random.RandomState(0)
        Xs = G.masks_from_seeds(d, list(range(n)), "shapley").astype(np.float64)
        w = rng.normal(size=(d, K))
        Ys = Xs @ w + 0.1 * rng.normal(size=(n, K))
        tests = []
        for t in range(3):
            Xt = G.masks_from_seeds(d, list(range(5000 + 100 * t, 5000 + 100 * t + m)), "datamodel").astype(np.float64)
            tests.append((Xt, Xt @ w + 0.5 * rng.normal(size=(m, K))))
        agg_ms = float("inf")
        for it in range(4):  # best of 4: the first passes pay cudaMalloc after the empty_cache() above
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            phi = G.data_shapley_batched(Xs, Ys, w.sum(axis=0), np.zeros(K))
            phb = G.data_banzhaf_batched(Xs, Ys)
            lds = G.evaluate_lds(phi, tests, K)
            torch.cuda.synchronize()
            agg_ms = min(agg_ms, (time.perf_counter() - t0) * 1e3)
        ridge_ms = float("inf")
        for it in range(3):  # datamodel estimator of lds.py:411-421: RidgeCV over 100 alphas for every behaviour
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = G.ridge_cv_batched(Xs, Ys)
            torch.cuda.synchronize()
            ridge_ms = min(ridge_ms, (time.perf_counter() - t0) * 1e3)
        out["ridge_cv"] = {"ms_host_to_host": ridge_ms, "n_masks": n, "contributors": d, "behaviors": K, "alphas": 100,
                           "what": "RidgeCV(alphas=linspace(0.01, 10, 100)) leave-one-out fit of every behaviour, numpy in / numpy out"}
        bytes_alg = n * d + 8 * n * K + 8 * d * K + 3 * (m * d + 8 * m * K) + 8 * d * K + 8 * 3 * K
        out["aggregation"] = {"ms_host_to_host": agg_ms, "n_masks": n, "contributors": d, "behaviors": K,
                              "algorithmic_bytes": bytes_alg, "lds": list(map(float, lds)),
                              "what": "data_shapley x K + data_banzhaf x K + evaluate_lds (3 x 100 subsets), numpy in / numpy out"}
    return out


if __name__ == "__main__":
    main()
