"""Aggregation kernels at BASELINE config 5 and at a bandwidth-meaningful scaled K (CUDA events, device-resident)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gadm_b200 as G
from gadm_b200 import aggregation as agg


def ev(fn, iters=3):
    best = 1e30
    for _ in range(iters):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--d", type=int, default=100)
    ap.add_argument("--K", type=int, default=1000)
    ap.add_argument("--m", type=int, default=100)
    ap.add_argument("--cpu-targets", type=int, default=0, help="time sklearn RidgeCV on this many behaviours (host)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(0)
    X = (rng.rand(a.n, a.d) > 0.5).astype(np.uint8)
    Xt = (rng.rand(a.m, a.d) > 0.5).astype(np.uint8)
    masks, tmasks = agg.PackedMasks(X, dev), agg.PackedMasks(Xt, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    Y = torch.randn(a.n, a.K, device=dev, dtype=torch.float64, generator=g)
    Yt = torch.randn(a.m, a.K, device=dev, dtype=torch.float64, generator=g)
    v0 = torch.zeros(a.K, device=dev, dtype=torch.float64)
    res = {"n": a.n, "d": a.d, "K": a.K, "m": a.m}
    ms, b = ev(lambda: masks.xty(Y, v0, 0.0, 1.0 / a.n))
    res["xty_ms"] = ms
    res["xty_GBps"] = (8.0 * a.n * a.K + 8.0 * a.d * a.K + a.n * a.d / 8) / ms / 1e6
    ms, A = ev(lambda: masks.gram(0)); res["gram_ms"] = ms
    ms, Ainv = ev(lambda: agg.sym_pinv(A, 1e-15)); res["pinv_ms"] = ms
    ms, phi = ev(lambda: agg._dgemm(Ainv, b, 1e-10)); res["dgemm_ms"] = ms
    res["dgemm_GBps"] = (16.0 * a.d * a.K) / ms / 1e6
    ms, pred = ev(lambda: tmasks.times(phi)); res["pred_ms"] = ms
    res["pred_GBps"] = (8.0 * a.d * a.K + 8.0 * a.m * a.K) / ms / 1e6
    ms, rho = ev(lambda: agg.spearman_matrix(tmasks, Yt, phi, as_numpy=False)); res["pred_plus_spearman_ms"] = ms
    res["spearman_GBps"] = (8.0 * a.d * a.K + 16.0 * a.m * a.K + 8.0 * a.K) / ms / 1e6
    ms, _ = ev(lambda: agg.data_shapley_batched(masks, Y, v0 + 1.0, v0, as_numpy=False)); res["shapley_total_ms"] = ms
    alg = a.n * a.d / 8 + 8.0 * a.n * a.K + 8.0 * a.d * a.K
    res["shapley_algorithmic_GBps"] = alg / ms / 1e6
    # datamodel estimator (lds.py:411-421): RidgeCV over 100 alphas for every behaviour
    from gadm_b200.datamodel import ridge_cv_batched
    Xf = torch.as_tensor(X.astype(np.float64)).to(dev)
    ms, rc = ev(lambda: ridge_cv_batched(Xf, Y, as_numpy=False)); res["ridgecv_total_ms"] = ms
    res["ridgecv_gcv_fp64_gflops"] = 2.0 * 100 * a.n * a.d * a.K / ms / 1e6
    if a.cpu_targets:
        import time
        from sklearn.linear_model import RidgeCV
        Yh = Y[:, :a.cpu_targets].cpu().numpy(); Xh = X.astype(np.float64)
        t0 = time.perf_counter()
        for i in range(a.cpu_targets):
            RidgeCV(alphas=np.linspace(0.01, 10, 100)).fit(Xh, Yh[:, i])
        res["ridgecv_sklearn_ms_per_target"] = (time.perf_counter() - t0) * 1e3 / a.cpu_targets
        res["ridgecv_sklearn_ms_extrapolated_K"] = res["ridgecv_sklearn_ms_per_target"] * a.K
    print(json.dumps(res))


if __name__ == "__main__":
    main()
