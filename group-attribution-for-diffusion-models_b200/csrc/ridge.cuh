// Datamodel estimator of lds.py:411-421:  RidgeCV(alphas=np.linspace(0.01, 10, 100)).fit(masks, targets[:, i])
// for every model behaviour i -- sklearn's efficient leave-one-out (GCV) ridge with an intercept -- batched over
// all K behaviours and all alphas.  (The reference refits, and re-decomposes the same mask matrix, once per
// behaviour.)
//
// sklearn 1.9 arithmetic restated (sklearn/linear_model/_ridge.py, _RidgeGCV, dense X, fit_intercept=True,
// gcv_mode "cov" for n > d / "gram" for n <= d; both are the same function of the centred data):
//   Xc = X - mean_rows(X), yc = y - mean(y);  C = Xc^T Xc = V L V^T
//   H^-1(a) = V diag(1 / (L + a)) V^T
//   alpha*c = yc - Xc H^-1 Xc^T yc
//   alpha*d = 1 - diag(Xc H^-1 Xc^T) - (1 - Xc H^-1 Xc^T 1) / n
//   looe = (alpha*c) / (alpha*d);  score(a) = -mean(looe^2);  alpha_ = first a with the largest score
//   coef = H^-1(alpha_) Xc^T yc;  intercept = mean(y) - mean_rows(X) . coef
// With Z = Xc V (n x d), T = Z^T Yc (d x K), q = Z^T 1 and w_j(a) = 1 / (L_j + a):
//   (Xc H^-1 Xc^T yc)_i = sum_j Z_ij w_j T_jk      diag_i = sum_j Z_ij^2 w_j      (Xc H^-1 Xc^T 1)_i = sum_j Z_ij w_j q_j
// so one decomposition serves every alpha and every behaviour; the work is the A x (n x d) x (d x K) contraction
// in ridge_gcv_score_kernel (fp64 FMA pipe), everything else is O(n d^2 + n d K).
// All fp64 SIMT with fixed summation orders (deterministic).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "aggregate.cuh"

namespace gadm {
namespace ridge {

// mean[j] = mean_i X[i, j] (sequential over i: deterministic); Xc = X - mean.  One thread per column.
__global__ void center_columns_kernel(const double* __restrict__ X, int64_t n, int64_t d, double* __restrict__ Xc,
                                      double* __restrict__ mean) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= d) return;
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += X[i * d + j];
  const double m = s / static_cast<double>(n);
  mean[j] = m;
  for (int64_t i = 0; i < n; ++i) Xc[i * d + j] = X[i * d + j] - m;
}

// C[i, k] = sum_j op(A)(i, j) * B[j, k];  op(A)(i, j) = A[i * lda + j] (kTransA = false) or A[j * lda + i] (true).
// B: [J, N] row pitch ldb, C: [M, N] row pitch ldc.  32 x 32 output tile, blockDim (32, 8), 4 rows per thread.
template <bool kTransA>
__global__ void dgemm_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                             int64_t M, int64_t J, int64_t N, double* __restrict__ Cout, int64_t ldc) {
  __shared__ double sA[32][33];
  __shared__ double sB[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * 32, k0 = static_cast<int64_t>(blockIdx.x) * 32;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t j0 = 0; j0 < J; j0 += 32) {
    for (int t = ty; t < 32; t += 8) {
      // sA[row i][col j]
      if (kTransA) {
        const int64_t j = j0 + t, i = i0 + tx;  // coalesced along i
        sA[tx][t] = (i < M && j < J) ? A[j * lda + i] : 0.0;
      } else {
        const int64_t i = i0 + t, j = j0 + tx;  // coalesced along j
        sA[t][tx] = (i < M && j < J) ? A[i * lda + j] : 0.0;
      }
      const int64_t jb = j0 + t, kb = k0 + tx;
      sB[t][tx] = (jb < J && kb < N) ? B[jb * ldb + kb] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const double b = sB[jj][tx];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[t] += sA[ty + 8 * t][jj] * b;
    }
    __syncthreads();
  }
  const int64_t k = k0 + tx;
  if (k < N) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int64_t i = i0 + ty + 8 * t;
      if (i < M) Cout[i * ldc + k] = acc[t];
    }
  }
}

// Symmetric eigendecomposition A = V diag(evals) V^T by the one-sided Jacobi of aggregate.cuh (one CTA, fp64).
// V: [d, d] row-major, COLUMN i = eigenvector i; evals unsorted.  work: 2 dp^2 doubles (smem when it fits).
__global__ void __launch_bounds__(agg::kPinvThreads, 1)
sym_eig_kernel(const double* __restrict__ A, int d, double* __restrict__ evals, double* __restrict__ V, double* gwork,
               int use_smem, int* __restrict__ info) {
  extern __shared__ double eig_smem[];
  const int dp = (d + 1) & ~1;
  double* G = use_smem ? eig_smem : gwork;
  double* Vt = G + static_cast<size_t>(dp) * dp;
  __shared__ int s_rot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = agg::kPinvThreads / 32;
  for (int idx = tid; idx < dp * dp; idx += agg::kPinvThreads) {
    const int i = idx / dp, j = idx % dp;
    G[idx] = (i < d && j < d) ? A[i * d + j] : 0.0;
    Vt[idx] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int sweeps = agg::jacobi_orthogonalise_rows(G, Vt, dp, &s_rot);
  __syncthreads();
  // lambda_i = g_i . v_i (signed); the padding row (d odd) carries lambda = 0 and the unit vector e_{dp-1}: dropped
  // by writing only the d x d block -- its coupling to the others is exactly zero because row/column dp-1 of A is zero.
  for (int i = warp; i < d; i += nwarps) {
    double a = 0.0;
    for (int j = lane; j < dp; j += 32) a += G[static_cast<size_t>(i) * dp + j] * Vt[static_cast<size_t>(i) * dp + j];
    a = agg::warp_sum(a);
    if (lane == 0) evals[i] = a;
  }
  for (int idx = tid; idx < d * d; idx += agg::kPinvThreads) {
    const int r = idx / d, c = idx % d;
    V[idx] = Vt[static_cast<size_t>(c) * dp + r];
  }
  if (tid == 0 && info) info[0] = sweeps;
}

// q[j] = sum_i Z[i, j]  (column sums, sequential: deterministic)
__global__ void column_sums_kernel(const double* __restrict__ Z, int64_t n, int64_t d, double* __restrict__ q) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= d) return;
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += Z[i * d + j];
  q[j] = s;
}

// den[a, i] = alpha*d of the header comment.  One warp per (a, i).
__global__ void ridge_denominator_kernel(const double* __restrict__ Z, const double* __restrict__ evals,
                                         const double* __restrict__ q, const double* __restrict__ alphas, int64_t n,
                                         int64_t d, int64_t A, double* __restrict__ den) {
  const int lane = threadIdx.x & 31;
  const int64_t job = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (job >= A * n) return;
  const int64_t a = job / n, i = job % n;
  const double alpha = alphas[a];
  double s2 = 0.0, s1 = 0.0;
  for (int64_t j = lane; j < d; j += 32) {
    const double z = Z[i * d + j], w = 1.0 / (evals[j] + alpha);
    s2 += z * z * w;
    s1 += z * w * q[j];
  }
  s2 = agg::warp_sum(s2);
  s1 = agg::warp_sum(s1);
  if (lane == 0) den[job] = 1.0 - s2 - (1.0 - s1) / static_cast<double>(n);
}

// score[a, k] = -(1/n) sum_i ( (Yc[i,k] - sum_j Z[i,j] w_j(a) T[j,k]) / den[a,i] )^2
//
// 2 A n d K flops of fp64 over 8 n K bytes of Yc: ~2500 flop / byte at (A, d) = (100, 100), i.e. bound by the fp64
// FMA pipe.  Round 1 ran one CTA per (alpha, 64 behaviours): Yc was re-read once per alpha (176.8 GB of DRAM traffic
// for ~2 GB at K = 2 * 10^5) and the 8-accumulator thread tile kept the pipe at 31 %.  Now one CTA owns a tile of
// 128 rows x 64 behaviours for ALL alphas: the Yc tile stays in shared memory across the alpha loop,
// Z^T (contiguous along rows) and T stream through shared memory by cp.async in chunks of 32 contraction indices, one
// chunk ahead of the arithmetic, w_j(alpha) is applied on the fly, and the inner loop is 32 DFMA per 6 LDS.128.  Row tiles write per-tile partial sums; ridge_gcv_reduce_kernel adds them
// in tile order (deterministic).  No limit on d any more.
constexpr int kGcvRows = 128;
constexpr int kGcvCols = 64;
constexpr int kGcvJ = 32;
constexpr int kGcvThreads = 512;  // 16 (column groups of 4) x 32 (row groups of 4): 16 warps per SM
// double-buffered chunks + reduce scratch + the Yc tile + two mbarriers (+ d doubles: w table)
constexpr int kGcvSmemBytes =
    2 * kGcvJ * (kGcvRows + kGcvCols) * 8 + (kGcvThreads / 16) * kGcvCols * 8 + kGcvRows * kGcvCols * 8 + 64;

// Zt[j, i] = Z[i, j]  (so that a row tile of Z^T is contiguous)
__global__ void transpose_f64_kernel(const double* __restrict__ in, int64_t rows, int64_t cols, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 32, r0 = static_cast<int64_t>(blockIdx.y) * 32;
  for (int t = threadIdx.y; t < 32; t += blockDim.y) {
    const int64_t r = r0 + t, c = c0 + threadIdx.x;
    tile[t][threadIdx.x] = (r < rows && c < cols) ? in[r * cols + c] : 0.0;
  }
  __syncthreads();
  for (int t = threadIdx.y; t < 32; t += blockDim.y) {
    const int64_t c = c0 + t, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[c * rows + r] = tile[threadIdx.x][t];
  }
}

__global__ void __launch_bounds__(kGcvThreads)
ridge_gcv_score_kernel(const double* __restrict__ Zt, const double* __restrict__ T, const double* __restrict__ Yc,
                       const double* __restrict__ evals, const double* __restrict__ den,
                       const double* __restrict__ alphas, int64_t n, int64_t d, int64_t K, int64_t A,
                       double* __restrict__ partial) {
  extern __shared__ __align__(16) double gcv_smem[];
  constexpr int kGroups = kGcvThreads / 16;               // row groups of 4
  double* zt = gcv_smem;                                  // [2][kGcvJ][kGcvRows]  Z^T chunk
  double* tt = gcv_smem + 2 * kGcvJ * kGcvRows;           // [2][kGcvJ][kGcvCols]  T chunk (unscaled)
  double* red = tt + 2 * kGcvJ * kGcvCols;                // [kGroups][kGcvCols]
  double* ycs = red + kGroups * kGcvCols;                 // [kGcvRows][kGcvCols]  Yc tile (read once per alpha)
  double* bars = ycs + kGcvRows * kGcvCols;               // two mbarriers
  double* wtab = bars + 8;                                // [d]  w_j(alpha) of the current alpha
  const uint32_t bar0 = smem_u32(bars);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t k0 = static_cast<int64_t>(blockIdx.x) * kGcvCols;
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * kGcvRows;
  const int64_t tiles = gridDim.y;
  const int vr = static_cast<int>((n - i0) < kGcvRows ? (n - i0) : kGcvRows);  // valid rows / columns of this tile
  const int vc = static_cast<int>((K - k0) < kGcvCols ? (K - k0) : kGcvCols);
  // one TMA bulk copy per tile row needs 16-byte aligned, even-length rows; otherwise per-thread 8-byte cp.async
  const bool bulk = (n % 2 == 0) && (K % 2 == 0) && ((reinterpret_cast<uintptr_t>(Zt) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(T) & 15) == 0);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  for (int idx = tid; idx < kGcvRows * kGcvCols; idx += kGcvThreads) {
    const int r = idx / kGcvCols, c = idx % kGcvCols;
    ycs[idx] = (r < vr && c < vc) ? Yc[(i0 + r) * K + k0 + c] : 0.0;
  }
  // rows / columns beyond the matrix are never copied: zero them once (both buffers)
  for (int idx = tid; idx < 2 * kGcvJ * kGcvRows; idx += kGcvThreads)
    if (idx % kGcvRows >= vr) zt[idx] = 0.0;
  for (int idx = tid; idx < 2 * kGcvJ * kGcvCols; idx += kGcvThreads)
    if (idx % kGcvCols >= vc) tt[idx] = 0.0;
  __syncthreads();
  const int nchunks = static_cast<int>((d + kGcvJ - 1) / kGcvJ);
  auto issue_chunk = [&](int buf, int c) {
    const int64_t j0 = static_cast<int64_t>(c) * kGcvJ;
    const int jn = static_cast<int>((d - j0) < kGcvJ ? (d - j0) : kGcvJ);
    if (bulk) {
      if (tid == 0) {
        mbar_arrive_expect_tx(bar0 + 8 * buf, static_cast<uint32_t>(jn * (vr + vc) * 8));
        for (int jj = 0; jj < jn; ++jj) {
          bulk_copy_global_to_smem(smem_u32(zt + (buf * kGcvJ + jj) * kGcvRows), Zt + (j0 + jj) * n + i0,
                                   static_cast<uint32_t>(vr * 8), bar0 + 8 * buf);
          bulk_copy_global_to_smem(smem_u32(tt + (buf * kGcvJ + jj) * kGcvCols), T + (j0 + jj) * K + k0,
                                   static_cast<uint32_t>(vc * 8), bar0 + 8 * buf);
        }
      }
    } else {
      for (int idx = tid; idx < jn * vr; idx += kGcvThreads)
        agg::cp_async_f64(zt + (buf * kGcvJ + idx / vr) * kGcvRows + idx % vr, Zt + (j0 + idx / vr) * n + i0 + idx % vr, true);
      for (int idx = tid; idx < jn * vc; idx += kGcvThreads)
        agg::cp_async_f64(tt + (buf * kGcvJ + idx / vc) * kGcvCols + idx % vc, T + (j0 + idx / vc) * K + k0 + idx % vc, true);
    }
  };
  auto wait_chunk = [&](int buf, uint32_t phase) {
    if (bulk) mbar_wait(bar0 + 8 * buf, phase, 0x910 + buf);
    else agg::cp_async_commit_wait_all();
  };

  uint32_t fills[2] = {0u, 0u};  // how often each buffer has been filled (mbarrier phase parity)
  for (int64_t a = 0; a < A; ++a) {
    const double alpha = alphas[a];
    double acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[p][c] = 0.0;
    issue_chunk(0, 0);
    for (int64_t j = tid; j < d; j += kGcvThreads) wtab[j] = 1.0 / (evals[j] + alpha);
    wait_chunk(0, fills[0] & 1u);
    ++fills[0];
    __syncthreads();
    int buf = 0;
    for (int c = 0; c < nchunks; ++c) {
      if (c + 1 < nchunks) issue_chunk(buf ^ 1, c + 1);
      const int64_t j0 = static_cast<int64_t>(c) * kGcvJ;
      const int jn = static_cast<int>((d - j0) < kGcvJ ? (d - j0) : kGcvJ);
      const double* zb = zt + buf * kGcvJ * kGcvRows + ty * 4;
      const double* tb = tt + buf * kGcvJ * kGcvCols + tx * 2;
#pragma unroll 4
      for (int jj = 0; jj < jn; ++jj) {
        const double w = wtab[j0 + jj];
        double x[4], y[4];
        const double2* xp = reinterpret_cast<const double2*>(zb + jj * kGcvRows);
        const double* yrow = tb + jj * kGcvCols;
#pragma unroll
        for (int q = 0; q < 2; ++q) { const double2 t2 = xp[q]; x[2 * q] = t2.x; x[2 * q + 1] = t2.y; }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const double2 t2 = *reinterpret_cast<const double2*>(yrow + q * 32);  // columns agg::tile_col(tx, .)
          y[2 * q] = t2.x * w; y[2 * q + 1] = t2.y * w;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) acc[p][cc] = fma(x[p], y[cc], acc[p][cc]);
      }
      if (c + 1 < nchunks) {
        wait_chunk(buf ^ 1, fills[buf ^ 1] & 1u);
        ++fills[buf ^ 1];
      }
      __syncthreads();
      buf ^= 1;
    }
    // squared leave-one-out errors of this thread's 4 rows, summed in row order; then over the row groups in order
    double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = ty * 4 + p;
      if (r < vr) {
        const double dd = den[a * n + i0 + r];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double e = (ycs[r * kGcvCols + agg::tile_col(tx, c)] - acc[p][c]) / dd;
          s[c] += e * e;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) red[ty * kGcvCols + agg::tile_col(tx, c)] = s[c];
    __syncthreads();
    if (tid < kGcvCols) {
      double t = 0.0;
#pragma unroll 8
      for (int g = 0; g < kGroups; ++g) t += red[g * kGcvCols + tid];
      if (tid < vc) partial[(a * tiles + blockIdx.y) * K + k0 + tid] = t;
    }
    __syncthreads();
  }
}

// score[a, k] = -(sum over row tiles, in order) / n
__global__ void ridge_gcv_reduce_kernel(const double* __restrict__ partial, int64_t A, int64_t tiles, int64_t K, int64_t n,
                                        double* __restrict__ score) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= A * K) return;
  const int64_t a = idx / K, k = idx % K;
  double t = 0.0;
  for (int64_t r = 0; r < tiles; ++r) t += partial[(a * tiles + r) * K + k];
  score[idx] = -t / static_cast<double>(n);
}

// Model selection (_RidgeGCV.fit loop): best[k] = first alpha index with the strictly largest score
// (per behaviour, or -- per_target == 0 -- of the mean over behaviours, shared by all k), then
// Ts[j, k] = T[j, k] / (evals[j] + alphas[best[k]]) so that coef = V Ts.
__global__ void ridge_select_kernel(const double* __restrict__ score, int64_t A, int64_t K, int per_target,
                                    const double* __restrict__ alphas, const double* __restrict__ evals,
                                    const double* __restrict__ T, int64_t d, int32_t* __restrict__ best,
                                    double* __restrict__ best_score, double* __restrict__ Ts) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int32_t bi = 0;
  double bs = 0.0;
  for (int64_t a = 0; a < A; ++a) {
    double sc;
    if (per_target) sc = score[a * K + k];
    else {  // np.mean(-squared_errors) over all n*K entries = mean_k score[a, k] (fixed order; every thread the same)
      double s = 0.0;
      for (int64_t kk = 0; kk < K; ++kk) s += score[a * K + kk];
      sc = s / static_cast<double>(K);
    }
    if (a == 0 || sc > bs) { bs = sc; bi = static_cast<int32_t>(a); }
  }
  best[k] = bi;
  best_score[k] = bs;
  const double alpha = alphas[bi];
  for (int64_t j = 0; j < d; ++j) Ts[j * K + k] = T[j * K + k] / (evals[j] + alpha);
}

// intercept[k] = ymean[k] - sum_j xmean[j] * coef[j, k]   (LinearModel._set_intercept)
__global__ void ridge_intercept_kernel(const double* __restrict__ coef, const double* __restrict__ xmean,
                                       const double* __restrict__ ymean, int64_t d, int64_t K,
                                       double* __restrict__ intercept) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double s = 0.0;
  for (int64_t j = 0; j < d; ++j) s += xmean[j] * coef[j * K + k];
  intercept[k] = ymean[k] - s;
}

}  // namespace ridge
}  // namespace gadm
