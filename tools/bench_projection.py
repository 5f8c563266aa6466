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
    ap.add_argument("--M", type=int, default=512)
    ap.add_argument("--type", default="rademacher")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--cta-group", type=int, default=2)
    a = ap.parse_args()
    dev = "cuda:0"
    p = CudaProjector(a.D, a.k, 42, ProjectionType(a.type), dev, 32, stage_rows=a.M, cta_group=a.cta_group)
    stage = p._stage_buffer(a.M)
    g = torch.Generator(device=dev).manual_seed(1234)
    for r in range(0, a.M, 64):  # fill the staged buffer in slices (bf16 randn * 1e-3)
        stage[r:r + 64, :a.D] = (torch.randn(min(64, a.M - r), a.D, device=dev, generator=g) * 1e-3).to(torch.bfloat16)
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
                      "norm_ratio": float(out.norm(dim=1).mean() / (stage[:, :a.D].float().norm(dim=1).mean() * a.k ** 0.5))}))


if __name__ == "__main__":
    main()
