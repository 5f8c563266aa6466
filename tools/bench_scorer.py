"""Timing of the TRAK scorer stages (CUDA events) at a given (N, k, T)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200 as G


def timed(fn, iters=2):
    best = 1e30
    out = None
    for _ in range(iters):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--t", type=int, default=1000)
    ap.add_argument("--totals-only", action="store_true", help="skip the per-stage timings (large k)")
    ap.add_argument("--primal-too", action="store_true", help="also time the forced k x k path when N < k")
    a = ap.parse_args()
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    train = torch.randn(a.n, a.k, device=dev, generator=g)
    gen = torch.randn(a.t, a.k, device=dev, generator=g)
    res = {"n": a.n, "k": a.k, "t": a.t}
    if a.totals_only:
        ms, out = timed(lambda: G.trak_scores(train, gen, variants=("trak",), return_scorer=True)); res["trak_total_ms"] = ms
        res["dual"] = bool(out[1].dual)
        ms, _ = timed(lambda: G.trak_scores(train, gen)); res["all_variants_total_ms"] = ms
        if a.primal_too:
            ms, outp = timed(lambda: G.trak_scores(train, gen, variants=("trak",), dual=False), iters=1)
            res["trak_total_primal_ms"] = ms
            res["max_rel_diff_dual_vs_primal"] = float((out[0]["trak"] - outp["trak"]).abs().max() / outp["trak"].abs().max())
        res["watchdog"] = G._lib.get_handle(dev).watchdog_code()
        print(json.dumps(res))
        return
    ms, phi_t = timed(lambda: G.transpose(train)); res["transpose_ms"] = ms
    ms, gram = timed(lambda: G.gemm_tn(phi_t, phi_t, lower_only=True, diag_add=0.5)); res["gram_ms"] = ms
    from gadm_b200.scoring import gram_lower, _gram_plan
    ms, gram_b = timed(lambda: gram_lower(train, 0.5), iters=3); res["transpose_plus_balanced_gram_ms"] = ms
    res["gram_plan"] = _gram_plan(a.n, a.k, train.device)
    res["balanced_gram_max_abs_diff_lower"] = float((torch.tril(gram_b) - torch.tril(gram)).abs().max())
    res["gram_tflops_fp32_equiv_full"] = 2.0 * a.n * a.k * a.k / ms / 1e9
    res["gram_tflops_computed_lower"] = res["gram_tflops_fp32_equiv_full"] * (0.5 + 64.0 / a.k)
    sc = G.TrakScorer(0.5)
    g2 = [gram.clone() for _ in range(3)]
    it = iter(g2)
    ms, _ = timed(lambda: sc._cholesky_(next(it)), iters=3); res["cholesky_ms"] = ms
    ms, _ = timed(lambda: sc._tri_inverse_(), iters=3); res["tri_inverse_ms"] = ms
    ms, z = timed(lambda: sc.solve_rows(gen)); res["solve_gen_ms"] = ms
    ms, s = timed(lambda: G.gemm_tn(z, train)); res["score_gemm_ms"] = ms
    res["score_gemm_tflops"] = 2.0 * a.t * a.k * a.n / ms / 1e9
    ms, _ = timed(lambda: G.col_mean_scaled(s)); res["col_mean_ms"] = ms
    ms, _ = timed(lambda: G.trak_scores(train, gen, variants=("trak",))); res["trak_total_ms"] = ms
    ms, _ = timed(lambda: G.trak_scores(train, gen)); res["all_variants_total_ms"] = ms
    # reference-style torch path on the same GPU (cuBLAS fp32 + torch.inverse), for context
    def ref():
        k = train.T @ train; k += 0.5 * torch.eye(a.k, device=dev); ki = torch.inverse(k)
        return (gen @ (ki @ train.T)).mean(dim=0)
    torch.backends.cuda.matmul.allow_tf32 = False
    ms, r = timed(ref); res["torch_fp32_reference_path_ms"] = ms
    ours = G.trak_scores(train, gen, variants=("trak",))["trak"]
    res["max_rel_diff_vs_torch"] = float((ours - r).abs().max() / r.abs().max())
    res["watchdog"] = G._lib.get_handle(dev).watchdog_code()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
