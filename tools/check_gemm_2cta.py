"""Check the CTA-pair GEMM (256 x 256 tiles, default for contractions >= 2048; GADM_GEMM_2CTA=0 disables it) against
fp64 and time the config-2 Gram with it.

    python tools/check_gemm_2cta.py [--quick]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gadm_b200 as G

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
res = {"env": os.environ.get("GADM_GEMM_2CTA")}
worst = 0.0
for (m, n, k, lower, tri) in [(384, 256, 2048, False, None), (1000, 520, 2100, False, None), (130, 300, 2048, False, None),
                              (1024, 1024, 4096, True, None), (700, 2560, 2560, False, "lower"), (700, 2560, 2560, False, "upper"),
                              (256, 128, 64, False, None),  # short contraction: the single-CTA kernel
                              (2048, 2048, 50000 if "--quick" not in sys.argv else 5000, True, None)]:
    a = torch.randn(m, k, device=dev, generator=g)
    b = torch.randn(n, k, device=dev, generator=g)
    if tri == "lower":
        b = torch.tril(b)
    elif tri == "upper":
        b = torch.triu(b)
    out = G.gemm_tn(a, b, lower_only=lower, b_tri=tri, diag_add=0.5 if lower else 0.0)
    torch.cuda.synchronize()
    want = a.double() @ b.double().T
    if lower:
        want = want + 0.5 * torch.eye(m, n, device=dev, dtype=torch.float64)
        # block-lower: tiles with col_tile <= row_tile
        rt = torch.arange(m, device=dev)[:, None] // 128
        ct = torch.arange(n, device=dev)[None, :] // 128
        mask = ct <= rt
        err = float(((out.double() - want) * mask).abs().max() / want.abs().max())
    else:
        err = float((out.double() - want).abs().max() / want.abs().max())
    res[f"{m}x{n}x{k}{'L' if lower else ''}{tri or ''}"] = err
    worst = max(worst, err)
res["worst_rel_err"] = worst
res["watchdog"] = G._lib.get_handle(torch.device(dev)).watchdog_code()
# beta / accumulate
a = torch.randn(512, 2304, device=dev, generator=g); b = torch.randn(384, 2304, device=dev, generator=g)
c0 = torch.randn(512, 384, device=dev, generator=g)
out = G.gemm_tn(a, b, out=c0.clone(), alpha=-1.0, beta=1.0)
res["beta_err"] = float((out.double() - (c0.double() - a.double() @ b.double().T)).abs().max())
if "--quick" in sys.argv:
    print(json.dumps(res))
    sys.exit(0)
# timing: config-2 Gram
phi_t = torch.randn(4096, 50000, device=dev, generator=g)
for _ in range(2):
    G.gemm_tn(phi_t, phi_t, lower_only=True, diag_add=0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    G.gemm_tn(phi_t, phi_t, lower_only=True, diag_add=0.5)
e1.record(); torch.cuda.synchronize()
res["gram_c2_ms"] = e0.elapsed_time(e1) / 3
print(json.dumps(res))
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    G.gemm_tn(phi_t, phi_t, lower_only=True, diag_add=0.5)
    torch.cuda.synchronize()
print([(e.key[:60], round(e.device_time_total / 1e3, 3)) for e in prof.key_averages() if "gemm" in e.key])
