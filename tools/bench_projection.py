"""Kernel-only timing of the projection at a named shape (CUDA events, staged bf16 input resident)."""
import argparse
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gadm_b200 import CudaProjector, ProjectionType


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=35_746_307)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--M", type=int, default=0, help="staged rows (default: 1024 for normal, 512 for rademacher)")
    ap.add_argument("--type", default="rademacher")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--cta-group", type=int, default=0, help="0 = projector default (4 for normal, 2 for rademacher)")
    a = ap.parse_args()
    dev = "cuda:0"
    if a.M == 0:
        a.M = 1024 if (a.type == "normal" and a.cta_group in (0, 4)) else 512
    p = CudaProjector(a.D, a.k, 42, ProjectionType(a.type), dev, 32, stage_rows=a.M, cta_group=a.cta_group or None)
    a.cta_group = p._group_for(a.M)
    stage = p._stage_buffer(a.M)  # tile-major [nkb, M, 64]
    g = torch.Generator(device=dev).manual_seed(1234)
    nkb = stage.shape[0]
    sq = torch.zeros(a.M, device=dev, dtype=torch.float64)
    for k0 in range(0, nkb, 8192):  # fill the staged buffer in slabs (bf16 randn * 1e-3)
        k1 = min(nkb, k0 + 8192)
        blk = (torch.randn(k1 - k0, a.M, 64, device=dev, generator=g) * 1e-3).to(torch.bfloat16)
        if k1 == nkb and a.D % 64:
            blk[-1, :, a.D % 64:] = 0  # positions beyond the gradient length stay zero
        stage[k0:k1] = blk
        sq += blk.double().pow(2).sum(dim=(0, 2))
    gnorm = sq.sqrt().mean()
    out = torch.empty(a.M, a.k, device=dev)
    p._project_rows(stage, a.M, 0, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p._project_rows(stage, a.M, 0, out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    flops = 2.0 * a.M * a.D * a.k
    print(json.dumps({"D": a.D, "k": a.k, "M": a.M, "type": a.type, "cta_group": a.cta_group, "ms": ts,
                      "tflops": flops / ms / 1e9, "frac_of_1590": flops / ms / 1e9 / 1590.4,
                      "gen_elems_per_s": a.D * a.k / ms * 1e3, "watchdog": p._handle.watchdog_code(),
                      "norm_ratio": float(out.double().norm(dim=1).mean() / (gnorm * a.k ** 0.5))}))


if __name__ == "__main__":
    main()
