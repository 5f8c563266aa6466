"""Staging throughput (gadm_stage_rows) alone and beside a running projection pass (CUDA events, one GPU).

    python tools/bench_staging.py [--aligned]   # --aligned: D rounded up to a multiple of 4 (16-byte aligned rows)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gadm_b200 import CudaProjector, ProjectionType
from gadm_b200.projectors import _as_blocks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=35_746_307)
    ap.add_argument("--aligned", action="store_true")
    ap.add_argument("--type", default="normal")
    ap.add_argument("--stage-dtype", default=None)
    ap.add_argument("--blocks", type=int, default=0, help="split the source into this many per-parameter blocks (U-Net like sizes)")
    ap.add_argument("--coresident", type=int, default=None, help="0 wide CTAs, 1 narrow (default: what the API picks)")
    a = ap.parse_args()
    D = (a.D + 3) // 4 * 4 if a.aligned else a.D
    dev = torch.device("cuda:0")
    rows = 1024 if a.type == "normal" else 512
    p = CudaProjector(D, 4096, 42, ProjectionType(a.type), dev, 32, stage_rows=rows, stage_dtype=a.stage_dtype)
    src = torch.randn(32, D, device=dev) * 1e-3
    if a.blocks:  # a mix of tiny (bias / norm) and large (conv / attention weight) tensors, separately allocated
        import numpy as np
        rng = np.random.RandomState(0)
        small = rng.choice([128, 256, 512], size=a.blocks // 2)
        big_n = a.blocks - len(small)
        big = rng.dirichlet(np.ones(big_n)) * (D - small.sum())
        big = np.maximum(1, big.astype(np.int64)); big[-1] += D - small.sum() - big.sum()
        sizes = np.empty(a.blocks, dtype=np.int64); sizes[0::2] = big[:len(sizes[0::2])]; sizes[1::2] = small[:len(sizes[1::2])]
        assert sizes.sum() == D and sizes.min() > 0
        cuts = np.concatenate([[0], np.cumsum(sizes)])
        parts = [src[:, int(lo):int(hi)].clone() for lo, hi in zip(cuts[:-1], cuts[1:])]
        del src
        blocks = _as_blocks(parts)
    else:
        blocks = _as_blocks(src)
    s0, s1 = p._stage(rows, 0), p._stage(rows, 1)
    out = torch.empty(rows, 4096, device=dev)
    bytes_per_add = 32 * D * 6
    co = bool(a.coresident) if a.coresident is not None else p._group_for(rows) != 4

    def stage_all(st):
        for r in range(0, rows, 32):
            p._pack(blocks, st, r, coresident=co)

    stage_all(s0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); stage_all(s0); e1.record(); torch.cuda.synchronize()
    alone_ms = e0.elapsed_time(e1)
    p._project_rows(s0, rows, 0, out)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(); p._project_rows(s0, rows, 0, out); k1.record(); torch.cuda.synchronize()
    proj_alone_ms = k0.elapsed_time(k1)
    side = torch.cuda.Stream(device=dev, priority=-1)
    res = {}
    for order in ("project_first", "stage_first", "queued"):
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        if order == "project_first":
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                a0.record(); p._project_rows(s0, rows, 0, out); a1.record()
            b0.record(); stage_all(s1); b1.record()
        elif order == "queued":  # as in the pipeline: everything is enqueued while an earlier staging still runs
            stage_all(s0)
            t0.record()
            ev = torch.cuda.Event(); ev.record()
            side.wait_event(ev)
            with torch.cuda.stream(side):
                a0.record(); p._project_rows(s0, rows, 0, out); a1.record()
            b0.record(); stage_all(s1); b1.record()
        else:
            b0.record(); stage_all(s1); b1.record()
            side.wait_event(b0)
            with torch.cuda.stream(side):
                a0.record(); p._project_rows(s0, rows, 0, out); a1.record()
        torch.cuda.current_stream().wait_stream(side)
        t1.record()
        torch.cuda.synchronize()
        res[order] = {"project_ms": a0.elapsed_time(a1), "stage_ms": b0.elapsed_time(b1), "both_ms": t0.elapsed_time(t1)}
    print(json.dumps({"D": D, "type": a.type, "stage_dtype": p.stage_dtype, "rows": rows, "coresident": co,
                      "stage_alone_ms": alone_ms, "stage_alone_gbs_algorithmic": rows / 32 * bytes_per_add / alone_ms / 1e6,
                      "project_alone_ms": proj_alone_ms, "concurrent": res, "watchdog": p._handle.watchdog_code()}))


if __name__ == "__main__":
    main()
