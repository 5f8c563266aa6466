import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200 as G
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(1024, 3000, device=dev, generator=g)
want = a.double() @ a.double().T
got = G.gemm_tn(a, a)
err = float((got.double() - want).abs().max() / want.abs().max())
phi_t = torch.randn(4096, 50000, device=dev, generator=g)
for _ in range(2):
    G.gemm_tn(phi_t, phi_t, lower_only=True, diag_add=0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    G.gemm_tn(phi_t, phi_t, lower_only=True, diag_add=0.5)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"n64": os.environ.get("GADM_GEMM_DEBUG_N64"), "rel_err": err, "gram_c2_ms": e0.elapsed_time(e1) / 3,
                  "watchdog": G._lib.get_handle(torch.device(dev)).watchdog_code()}))
