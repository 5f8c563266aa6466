"""Warm per-launch time of the diagonal-block kernel: gadm_cholesky on a 128 x 128 matrix is exactly one potrf launch."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gadm_b200 import _lib as L

dev = torch.device("cuda:0")
h = L.get_handle(dev)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(1000, 128, device=dev, generator=g)
spd = (x.T @ x + 0.5 * torch.eye(128, device=dev)).contiguous()
mats = [spd.clone() for _ in range(200)]
blocks = torch.empty(int(h.lib.gadm_cholesky_workspace_bytes(128)), dtype=torch.uint8, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
st = L.stream_ptr(dev)
def run(m):
    L.check(h.lib.gadm_cholesky(h.ptr, m.data_ptr(), 128, 128, blocks.data_ptr(), blocks.numel(), C.cast(info.data_ptr(), C.POINTER(C.c_int)), st))
for m in mats[:20]: run(m)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for m in mats[20:]: run(m)
e1.record(); torch.cuda.synchronize()
Ld = torch.tril(mats[-1].double())
print(json.dumps({"potrf_us_per_launch_incl_memset": e0.elapsed_time(e1) * 1e3 / 180,
                  "recon_err": float((Ld @ Ld.T - spd.double()).abs().max() / spd.abs().max())}))
