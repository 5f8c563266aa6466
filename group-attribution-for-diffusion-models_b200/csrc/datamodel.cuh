// Bootstrapped datamodel scores: src/attributions/methods/datamodel.py:8-37
//     for _ in range(num_runs):
//         idx = np.random.choice(n, n, replace=True)
//         reg = RidgeCV(cv=5, alphas=[0.1, 1.0, 10.0]).fit(x_train[idx], y_train[idx]);  coeff.append(reg.coef_)
// i.e. per bootstrap resample a 5-fold grid search over alpha (sklearn GridSearchCV(Ridge(fit_intercept=True)),
// KFold without shuffling, score = R^2 on the held-out fold, mean over folds, first best alpha) and a refit on the
// whole resample.  x_train is [n retrained models, d training examples] with d >> n, so every ridge fit is done in
// "Gram space": with the uncentred Gram G0 = X X^T of the ORIGINAL rows (exact integers from mask popcounts) a
// resample is just an index list, the centred train kernel is
//     Kc[p, q] = G0[p, q] - m_p - m_q + mm,   m_p = mean_{q in train} G0[p, q],  mm = mean_{p in train} m_p
// the dual coefficients c = (Kc + alpha I)^-1 (y - ybar), held-out predictions (Kc_test,train c + ybar) and the final
// coefficients coef = X^T w with w_i = sum_{p: idx_p = i} c_p - count_i * sum(c) / n  (centring folded in).
// One CTA per (resample, fold or refit, alpha) system: build Kc + alpha I, fp64 Cholesky in a global workspace slot,
// two triangular solves, then either the fold's R^2 or the scattered dual weights.  All fp64, deterministic.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gadm {
namespace dm {

constexpr int kThreads = 1024;

struct System {
  int32_t run;     // bootstrap resample
  int32_t f0, f1;  // held-out positions [f0, f1) of the resample; f0 == f1: refit on everything
  int32_t out;     // slot of the result: score index (fold systems) or run (refit systems)
  double alpha;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kThreads / 32; ++w) s += red[w];  // fixed order: deterministic
  __syncthreads();
  return s;
}

// G0: [n, n] Gram of the original rows; y: [n]; idx: [runs, n] resample indices; work: per-CTA slots of
// (n*n + 4*n) doubles; scores: R^2 per fold system; wdual: [runs, n] scattered dual weights of the refits.
__global__ void __launch_bounds__(kThreads, 1)
ridge_fold_kernel(const double* __restrict__ G0, const double* __restrict__ y, const int32_t* __restrict__ idx, int n,
                  const System* __restrict__ systems, int n_systems, double* __restrict__ work,
                  double* __restrict__ scores, double* __restrict__ wdual) {
  __shared__ double red[kThreads / 32];
  __shared__ double s_piv;
  const int tid = threadIdx.x;
  double* A = work + static_cast<size_t>(blockIdx.x) * (static_cast<size_t>(n) * n + 4 * static_cast<size_t>(n));
  double* m = A + static_cast<size_t>(n) * n;  // m_p for every position of the resample
  double* c = m + n;                           // rhs -> dual coefficients (train positions, compacted)
  double* yv = c + n;                          // y of every position
  int32_t* pos = reinterpret_cast<int32_t*>(yv + n);  // compacted train position -> resample position

  for (int s = blockIdx.x; s < n_systems; s += gridDim.x) {
    const System sys = systems[s];
    const int32_t* id = idx + static_cast<size_t>(sys.run) * n;
    const int nte = sys.f1 - sys.f0, ntr = n - nte;
    for (int p = tid; p < ntr; p += kThreads) pos[p] = (p < sys.f0) ? p : p + nte;
    for (int p = tid; p < n; p += kThreads) yv[p] = y[id[p]];
    __syncthreads();
    // m_p = mean over train q of G0[id_p, id_q]  (one warp per p, lanes over q, fixed tree)
    for (int p = tid >> 5; p < n; p += kThreads / 32) {
      const double* row = G0 + static_cast<size_t>(id[p]) * n;
      double a = 0.0;
      for (int q = tid & 31; q < ntr; q += 32) a += row[id[pos[q]]];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if ((tid & 31) == 0) m[p] = a / ntr;
    }
    __syncthreads();
    double part = 0.0, ypart = 0.0;
    for (int q = tid; q < ntr; q += kThreads) { part += m[pos[q]]; ypart += yv[pos[q]]; }
    const double mm = block_sum(part, red) / ntr;
    const double ybar = block_sum(ypart, red) / ntr;
    // A = Kc + alpha I (lower triangle), rhs = y - ybar
    for (size_t e = tid; e < static_cast<size_t>(ntr) * ntr; e += kThreads) {
      const int p = static_cast<int>(e / ntr), q = static_cast<int>(e % ntr);
      if (q <= p) {
        const int pp = pos[p], qq = pos[q];
        A[static_cast<size_t>(p) * ntr + q] =
            G0[static_cast<size_t>(id[pp]) * n + id[qq]] - m[pp] - m[qq] + mm + ((p == q) ? sys.alpha : 0.0);
      }
    }
    for (int p = tid; p < ntr; p += kThreads) c[p] = yv[pos[p]] - ybar;
    __syncthreads();
    // ---- right-looking Cholesky, column by column
    for (int j = 0; j < ntr; ++j) {
      if (tid == 0) s_piv = sqrt(A[static_cast<size_t>(j) * ntr + j]);
      __syncthreads();
      const double piv = s_piv;
      if (tid == 0) A[static_cast<size_t>(j) * ntr + j] = piv;
      for (int i = j + 1 + tid; i < ntr; i += kThreads) A[static_cast<size_t>(i) * ntr + j] /= piv;
      __syncthreads();
      const int rem = ntr - 1 - j;
      // trailing lower triangle: row i = j+1+r, column t = j+1+cidx, cidx <= r; flattened over (r, cidx <= r)
      const size_t tri = static_cast<size_t>(rem) * (rem + 1) / 2;
      for (size_t e = tid; e < tri; e += kThreads) {
        // invert e = r(r+1)/2 + cidx
        int r = static_cast<int>((sqrt(8.0 * static_cast<double>(e) + 1.0) - 1.0) * 0.5);
        while (static_cast<size_t>(r) * (r + 1) / 2 > e) --r;
        while (static_cast<size_t>(r + 1) * (r + 2) / 2 <= e) ++r;
        const int cidx = static_cast<int>(e - static_cast<size_t>(r) * (r + 1) / 2);
        const int i = j + 1 + r, t = j + 1 + cidx;
        A[static_cast<size_t>(i) * ntr + t] -= A[static_cast<size_t>(i) * ntr + j] * A[static_cast<size_t>(t) * ntr + j];
      }
      __syncthreads();
    }
    // ---- forward L z = rhs, backward L^T c = z (warp 0; dot products across lanes in a fixed tree)
    if (tid < 32) {
      for (int i = 0; i < ntr; ++i) {
        double a = 0.0;
        for (int t = tid; t < i; t += 32) a += A[static_cast<size_t>(i) * ntr + t] * c[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (tid == 0) c[i] = (c[i] - a) / A[static_cast<size_t>(i) * ntr + i];
        __syncwarp();
      }
      for (int i = ntr - 1; i >= 0; --i) {
        double a = 0.0;
        for (int t = i + 1 + tid; t < ntr; t += 32) a += A[static_cast<size_t>(t) * ntr + i] * c[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (tid == 0) c[i] = (c[i] - a) / A[static_cast<size_t>(i) * ntr + i];
        __syncwarp();
      }
    }
    __syncthreads();
    if (nte > 0) {
      // ---- R^2 on the held-out positions (sklearn r2_score): 1 - sum (y - yhat)^2 / sum (y - mean(y_test))^2
      double yte = 0.0;
      for (int t = sys.f0 + tid; t < sys.f1; t += kThreads) yte += yv[t];
      const double ymean_te = block_sum(yte, red) / nte;
      double ss_res = 0.0, ss_tot = 0.0;
      for (int t = sys.f0 + (tid >> 5); t < sys.f1; t += kThreads / 32) {
        const double* row = G0 + static_cast<size_t>(id[t]) * n;
        double a = 0.0;
        for (int q = tid & 31; q < ntr; q += 32) {
          const int qq = pos[q];
          a += (row[id[qq]] - m[t] - m[qq] + mm) * c[q];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((tid & 31) == 0) {
          const double e = yv[t] - (a + ybar), dlt = yv[t] - ymean_te;
          ss_res += e * e;
          ss_tot += dlt * dlt;
        }
      }
      const double r = block_sum(ss_res, red), tt = block_sum(ss_tot, red);
      // sklearn.metrics.r2_score(force_finite=True): a constant held-out target scores 1 if predicted exactly, else 0
      if (tid == 0) scores[sys.out] = (tt != 0.0) ? 1.0 - r / tt : ((r != 0.0) ? 0.0 : 1.0);
    } else {
      // ---- refit: w_i = sum_{p: id_p = i} c_p - count_i * sum(c) / n  (one thread per original row, fixed order)
      double cs = 0.0;
      for (int p = tid; p < n; p += kThreads) cs += c[p];
      const double csum = block_sum(cs, red);
      double* w = wdual + static_cast<size_t>(sys.out) * n;
      for (int i = tid; i < n; i += kThreads) {
        double a = 0.0;
        int cnt = 0;
        for (int p = 0; p < n; ++p)
          if (id[p] == i) { a += c[p]; ++cnt; }
        w[i] = a - cnt * csum / n;
      }
    }
    __syncthreads();
  }
}

}  // namespace dm
}  // namespace gadm
