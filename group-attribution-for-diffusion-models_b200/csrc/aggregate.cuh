// Subset-mask regressions (Shapley / Banzhaf), LDS rank correlations, group reductions, stable ranks.
// All fp64 / integer SIMT kernels: the problems are HBM/L2- and latency-bound (12.9 MB at
// BASELINE config 5), so the rules that matter are coalesced loads, bit-packed masks and fixed
// summation orders -- no tensor cores.
//
// Reference arithmetic restated here:
//   data_shapley   src/attributions/methods/datashapley.py:8-48
//   data_banzhaf   src/attributions/methods/databanzhaf.py:5-26
//   evaluate_lds   text_to_image/shapley_lds.py:138-150, lds.py:158-170 (scipy.stats.spearmanr)
//   bootstrap stat lds.py:458-476
//   group sums     text_to_image/traks.py:188-204 ; ranks traks.py:216-218, shapley_lds.py:294
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gadm_ptx.cuh"

namespace gadm {
namespace agg {

// ------------------------------------------------------------------ mask packing
// X: [n, d] uint8 (0/1).  rowbits: [n, wd] (bit i%32 of word i/32 = X[r, i]); colbits: [d, wn].
__global__ void pack_masks_kernel(const uint8_t* __restrict__ X, int64_t n, int64_t d, uint32_t* __restrict__ rowbits,
                                  int64_t wd, uint32_t* __restrict__ colbits, int64_t wn) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // first n*wd threads build row words, the next d*wn threads build column words
  if (tid < n * wd) {
    const int64_t r = tid / wd, w = tid % wd;
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
      const int64_t i = w * 32 + b;
      if (i < d && X[r * d + i]) bits |= (1u << b);
    }
    rowbits[tid] = bits;
  } else if (tid < n * wd + d * wn) {
    const int64_t t = tid - n * wd;
    const int64_t w = t / d, i = t % d;  // consecutive threads -> consecutive i (coalesced bytes)
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
      const int64_t r = w * 32 + b;
      if (r < n && X[r * d + i]) bits |= (1u << b);
    }
    colbits[i * wn + w] = bits;
  }
}

// ------------------------------------------------------------------ normal equations
// mode 0 (Shapley):  A[i,j] = N11(i,j) / n                        (datashapley.py:29)
// mode 1 (Banzhaf):  A[i,j] = N11 - (c_i + c_j)/2 + n/4           (databanzhaf.py:20-22, (X-1/2)^T (X-1/2))
// N11 = co-occurrence count (exact integer), c_i = column count.
__global__ void mask_gram_kernel(const uint32_t* __restrict__ colbits, int64_t d, int64_t wn, int64_t n, int mode,
                                 double* __restrict__ A) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= d * d) return;
  const int64_t i = tid / d, j = tid % d;
  int64_t n11 = 0, ci = 0, cj = 0;
  for (int64_t w = 0; w < wn; ++w) {
    const uint32_t a = colbits[i * wn + w], b = colbits[j * wn + w];
    n11 += __popc(a & b);
    ci += __popc(a);
    cj += __popc(b);
  }
  if (mode == 0) A[tid] = static_cast<double>(n11) / static_cast<double>(n);
  else if (mode == 2) A[tid] = static_cast<double>(n11);  // raw co-occurrence count (X X^T when given the row bit planes)
  else A[tid] = static_cast<double>(n11) - 0.5 * static_cast<double>(ci + cj) + 0.25 * static_cast<double>(n);
}

// out[i, k] = ( sum_r X[r,i] * (Y[r,k] - shift[k]) - half * sum_r (Y[r,k] - shift[k]) ) * scale
//   Shapley: shift = v0, half = 0, scale = 1/n   (datashapley.py:30)
//   Banzhaf: shift = 0 (null), half = 0.5, scale = 1   (databanzhaf.py:23)
//
// This is the dense fp64 contraction X^T Y with a 0/1 left operand: 2 d n K flops over 8 n K bytes of Y, i.e. d / 4
// flop per byte (25 at d = 100) -- far to the right of the fp64 ridge (~6 flop / B on B200), so the bound is the
// fp64 FMA pipe, not HBM.  Round 1 did it with predicated adds (one useful flop and ~3 issue slots per (player, row,
// behaviour): 28 % of the fp64 pipe, 0.4 TB/s of Y).  Now a register-tiled DFMA GEMM: the mask bits of a 16-row chunk
// are expanded once per CTA into 0.0 / 1.0 doubles in shared memory, Y arrives as coalesced 512-byte row segments,
// and every thread owns 8 players x 4 behaviours (32 independent accumulators; 32 DFMA per 6 LDS.128).  fma(x, y, acc)
// with x in {0, 1} is exactly "acc += y or nothing" and every (player, behaviour) sum still visits the rows in order,
// so the results are bit-identical to the predicated form and deterministic.
// CTA = 256 threads = 16 (behaviour groups of 4) x 16 (player groups of 8) -> tile 128 players x 64 behaviours.
constexpr int kXtyCols = 64;      // behaviours per CTA
constexpr int kXtyPlayers = 128;  // outputs (players, or test rows for X_test @ attrs) per CTA
constexpr int kXtyRows = 32;      // summed rows per shared-memory chunk
constexpr int kXtyThreads = 256;

// 8-byte asynchronous global -> shared copy; src_bytes = 0 writes zeros instead (bounds handling without a branch)
__device__ __forceinline__ void cp_async_f64(double* smem_dst, const double* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Column ownership inside a 64-column tile row: thread tx owns columns {2 tx, 2 tx + 1, 32 + 2 tx, 33 + 2 tx}, so
// that the 16 threads of one LDS.128 read 256 consecutive bytes of the natural row layout (columns 4 tx .. 4 tx + 3
// would be 16 segments 32 bytes apart: 4-way bank conflicts).
__device__ __forceinline__ int tile_col(int tx, int c) { return (c >> 1) * 32 + 2 * tx + (c & 1); }

// The next chunk's Y rows travel global -> shared asynchronously while the current chunk is being multiplied (a plain
// load + store parks every warp on the store until the load returns: the in-order issue exposed the full DRAM latency
// once per chunk and held the fp64 pipe at 39 %).  One thread issues one TMA bulk copy per 512-byte row (mbarrier
// completion); per-thread 8-byte cp.async -- the fallback for odd K / unaligned Y -- worked too but 1536 small
// requests per chunk kept the LSU queue full (ncu: stall_lg 26 % in the sister kernel of ridge.cuh).  The chunk's mask
// words wait in a register meanwhile and are expanded to 0.0 / 1.0 after the multiply.
constexpr int kXtySmemBytes = 2 * kXtyRows * (kXtyPlayers + kXtyCols) * 8 + 64;
template <bool kShift>
__global__ void __launch_bounds__(kXtyThreads, 2)
mask_xty_kernel(const uint32_t* __restrict__ rowbits, int64_t wd, const double* __restrict__ Y, int64_t n, int64_t d,
                int64_t K, const double* __restrict__ shift, double half, double scale, double* __restrict__ out) {
  extern __shared__ __align__(16) double xty_smem[];
  double (*xs)[kXtyRows][kXtyPlayers] = reinterpret_cast<double (*)[kXtyRows][kXtyPlayers]>(xty_smem);  // 0.0 / 1.0
  double (*ys)[kXtyRows][kXtyCols] =
      reinterpret_cast<double (*)[kXtyRows][kXtyCols]>(xty_smem + 2 * kXtyRows * kXtyPlayers);          // raw Y
  const uint32_t bar0 = smem_u32(xty_smem + 2 * kXtyRows * (kXtyPlayers + kXtyCols));
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t k0 = static_cast<int64_t>(blockIdx.x) * kXtyCols;
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * kXtyPlayers;
  const int64_t word0 = i0 >> 5;
  const int vc = static_cast<int>((K - k0) < kXtyCols ? (K - k0) : kXtyCols);  // valid columns of this tile
  const bool bulk = (K % 2 == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  double sh[4] = {0.0, 0.0, 0.0, 0.0};  // shift of this thread's columns
  if (kShift) {
#pragma unroll
    for (int c = 0; c < 4; ++c) sh[c] = (tile_col(tx, c) < vc) ? shift[k0 + tile_col(tx, c)] : 0.0;
  }
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  // columns beyond K are never copied: zero them once (both buffers)
  for (int idx = tid; idx < 2 * kXtyRows * kXtyCols; idx += kXtyThreads)
    if (idx % kXtyCols >= vc) (&ys[0][0][0])[idx] = 0.0;
  __syncthreads();

  auto issue_y = [&](int buf, int64_t r0) {
    const int rows = static_cast<int>((n - r0) < kXtyRows ? (n - r0) : kXtyRows);
    if (bulk) {
      if (tid == 0) {
        mbar_arrive_expect_tx(bar0 + 8 * buf, static_cast<uint32_t>(rows * vc * 8));
        for (int r = 0; r < rows; ++r)
          bulk_copy_global_to_smem(smem_u32(&ys[buf][r][0]), Y + (r0 + r) * K + k0, static_cast<uint32_t>(vc * 8),
                                   bar0 + 8 * buf);
      }
    } else {
      for (int idx = tid; idx < rows * vc; idx += kXtyThreads) {
        const int r = idx / vc, c = idx % vc;
        cp_async_f64(&ys[buf][r][c], Y + (r0 + r) * K + k0 + c, true);
      }
    }
  };
  auto wait_y = [&](int buf, uint32_t phase) {
    if (bulk) mbar_wait(bar0 + 8 * buf, phase, 0x900 + buf);
    else cp_async_commit_wait_all();
  };
  auto load_bits = [&](int64_t r0) -> uint32_t {
    if (tid >= kXtyRows * 4) return 0u;
    const int64_t rw = r0 + (tid >> 2);
    const int w = tid & 3;
    return (rw < n && word0 + w < wd) ? rowbits[rw * wd + word0 + w] : 0u;
  };
  auto expand_bits = [&](int buf, uint32_t bits) {
    if (tid < kXtyRows * 4) {
      double* dst = &xs[buf][tid >> 2][(tid & 3) * 32];
#pragma unroll
      for (int t = 0; t < 32; ++t) dst[t] = ((bits >> t) & 1u) ? 1.0 : 0.0;
    }
  };

  double acc[8][4];
  double tot[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = 0.0;

  issue_y(0, 0);
  expand_bits(0, load_bits(0));
  wait_y(0, 0);
  __syncthreads();
  int buf = 0;
  uint32_t it = 0;
  for (int64_t r0 = 0; r0 < n; r0 += kXtyRows, ++it) {
    const bool more = r0 + kXtyRows < n;
    uint32_t next_bits = 0u;
    if (more) {
      issue_y(buf ^ 1, r0 + kXtyRows);  // the other buffer: last read before the previous barrier
      next_bits = load_bits(r0 + kXtyRows);
    }
    const int rows = static_cast<int>((n - r0) < kXtyRows ? (n - r0) : kXtyRows);
#pragma unroll 4
    for (int rr = 0; rr < rows; ++rr) {
      double x[8], y[4];
      const double2* xp = reinterpret_cast<const double2*>(&xs[buf][rr][ty * 8]);
#pragma unroll
      for (int q = 0; q < 4; ++q) { const double2 t2 = xp[q]; x[2 * q] = t2.x; x[2 * q + 1] = t2.y; }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double2 t2 = *reinterpret_cast<const double2*>(&ys[buf][rr][q * 32 + tx * 2]);
        y[2 * q] = t2.x; y[2 * q + 1] = t2.y;
      }
      if (kShift) {
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] -= sh[c];
      }
      if (ty == 0) {  // the column totals are the same for every player group: one group computes them (rows in order)
#pragma unroll
        for (int c = 0; c < 4; ++c) tot[c] += y[c];
      }
#pragma unroll
      for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[p][c] = fma(x[p], y[c], acc[p][c]);
    }
    if (more) {
      expand_bits(buf ^ 1, next_bits);
      wait_y(buf ^ 1, ((it + 1) >> 1) & 1u);  // buffer b is filled for the (j + 1)-th time at iteration 2 j + b
    }
    __syncthreads();
    buf ^= 1;
  }
  // hand the column totals to the other player groups through the (now idle) Y buffer
  double* tot_s = &ys[0][0][0];
  if (ty == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) tot_s[tile_col(tx, c)] = tot[c];
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 4; ++c) tot[c] = tot_s[tile_col(tx, c)];
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int64_t i = i0 + ty * 8 + p;
    if (i >= d) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = tile_col(tx, c);
      if (col < vc) out[i * K + k0 + col] = (acc[p][c] - half * tot[c]) * scale;
    }
  }
}

// X_test @ attrs (shapley_lds.py:145): out[r, k] = sum_i X[r,i] * M[i,k] is the same kernel with the roles of
// rows and players swapped -- pass the *column* bit planes (bits over rows r for each player i) as `rowbits`,
// n := d (the summed index), d := m (the outputs): every M[i, :] row is then streamed exactly once.

// ------------------------------------------------------------------ symmetric pseudo-inverse
// One-sided (Hestenes) Jacobi SVD of a symmetric matrix, fp64, one CTA.  Rows of G start as the rows
// (= columns) of A and are rotated until mutually orthogonal; Vt accumulates the rotations.  Then
//   pinv(A) = sum_{i: sigma_i > rcond * sigma_max} v_i g_i^T / sigma_i^2 ,  sigma_i = ||g_i||
// which is numpy.linalg.pinv's SVD formula with its cutoff (datashapley.py:37 uses the default
// rcond = 1e-15; numpy.linalg.lstsq(rcond=None) in databanzhaf.py:20-25 uses eps * d).
// work: 2 * dp * dp doubles (dp = d rounded up to even) + dp doubles; lives in smem when it fits.
constexpr int kPinvThreads = 1024;
constexpr int kPinvMaxSweeps = 40;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One-sided (Hestenes) Jacobi: rotate the rows of G (dp x dp, dp even) until mutually orthogonal, accumulating the
// rotations in Vt.  Whole CTA (kPinvThreads); returns the number of sweeps.  For symmetric A = V L V^T the rows end
// as g_i = lambda_i v_i^T with v_i^T = row i of Vt.
__device__ __forceinline__ int jacobi_orthogonalise_rows(double* G, double* Vt, int dp, int* s_rot_p) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kPinvThreads / 32;
  int& s_rot = *s_rot_p;
  const int npairs = dp / 2;
  int sweep = 0;
  for (; sweep < kPinvMaxSweeps; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int round = 0; round < dp - 1; ++round) {
      // round-robin tournament: dp/2 disjoint pairs per round, every pair once per sweep
      for (int pr = warp; pr < npairs; pr += nwarps) {
        // circle method: (dp-1, round) and every {a, b} with a + b == 2*round (mod dp-1)
        const int a = (pr == 0) ? dp - 1 : (round + pr) % (dp - 1);
        const int b = (pr == 0) ? round : (round - pr + (dp - 1)) % (dp - 1);
        const int p = a < b ? a : b, q = a < b ? b : a;
        double* gp = G + static_cast<size_t>(p) * dp;
        double* gq = G + static_cast<size_t>(q) * dp;
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int j = lane; j < dp; j += 32) {
          const double x = gp[j], y = gq[j];
          alpha += x * x; beta += y * y; gamma += x * y;
        }
        alpha = warp_sum(alpha); beta = warp_sum(beta); gamma = warp_sum(gamma);
        if (fabs(gamma) > 1e-15 * sqrt(alpha * beta) && gamma != 0.0) {
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
          double* vp = Vt + static_cast<size_t>(p) * dp;
          double* vq = Vt + static_cast<size_t>(q) * dp;
          for (int j = lane; j < dp; j += 32) {
            const double x = gp[j], y = gq[j];
            gp[j] = c * x - s * y; gq[j] = s * x + c * y;
            const double vx = vp[j], vy = vq[j];
            vp[j] = c * vx - s * vy; vq[j] = s * vx + c * vy;
          }
          if (lane == 0) atomicAdd(&s_rot, 1);
        }
      }
      __syncthreads();
    }
    const int rot = s_rot;
    __syncthreads();
    if (rot == 0) break;
  }

  return sweep;
}

__global__ void __launch_bounds__(kPinvThreads, 1)
sym_pinv_kernel(const double* __restrict__ A, int d, double rcond, double* __restrict__ out, double* gwork,
                int use_smem, int* __restrict__ info) {
  extern __shared__ double pinv_smem[];
  const int dp = (d + 1) & ~1;
  double* G = use_smem ? pinv_smem : gwork;
  double* Vt = G + static_cast<size_t>(dp) * dp;
  double* sig2 = Vt + static_cast<size_t>(dp) * dp;
  __shared__ int s_rot;
  __shared__ double s_max;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kPinvThreads / 32;

  for (int idx = tid; idx < dp * dp; idx += kPinvThreads) {
    const int i = idx / dp, j = idx % dp;
    G[idx] = (i < d && j < d) ? A[i * d + j] : 0.0;
    Vt[idx] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();

  // ---- fast path: A symmetric positive definite and well conditioned (the usual n >= d case) -> A^-1 by Cholesky.
  // A pivot below 1e-8 * max diag means sigma_min / sigma_max could approach numpy's cut-off, so only then is the
  // Jacobi SVD (exact pinv semantics, rank-deficient input) run.  For a full-rank matrix inv == pinv up to kappa*eps.
  __shared__ int s_chol_ok;
  __shared__ double s_maxdiag;
  if (tid == 0) {
    double mx = 0.0;
    for (int i = 0; i < d; ++i) mx = fmax(mx, G[static_cast<size_t>(i) * dp + i]);
    s_maxdiag = mx;
    s_chol_ok = (mx > 0.0) ? 1 : 0;
  }
  __syncthreads();
  for (int j = 0; j < d && s_chol_ok; ++j) {
    const double piv = G[static_cast<size_t>(j) * dp + j];
    __syncthreads();
    if (!(piv > 1e-8 * s_maxdiag)) {
      if (tid == 0) s_chol_ok = 0;
      __syncthreads();
      break;
    }
    const double sd = sqrt(piv);
    if (tid == 0) G[static_cast<size_t>(j) * dp + j] = sd;
    for (int i = j + 1 + tid; i < d; i += kPinvThreads) G[static_cast<size_t>(i) * dp + j] /= sd;
    __syncthreads();
    const int m = d - 1 - j;  // trailing rows j+1..d-1, lower triangle incl. diagonal
    for (int e = tid; e < m * m; e += kPinvThreads) {
      const int i = j + 1 + e / m, c = j + 1 + e % m;
      if (c <= i) G[static_cast<size_t>(i) * dp + c] -= G[static_cast<size_t>(i) * dp + j] * G[static_cast<size_t>(c) * dp + j];
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_chol_ok) {
    // Vt <- L^-1 (lower triangular), column c handled by 4 lanes of one warp
    const int c = tid >> 2, part = tid & 3;
    for (int c0 = 0; c0 < d; c0 += kPinvThreads / 4) {
      const int col = c0 + c;
      const bool active = col < d;
      const int cc = active ? col : 0;
      if (active && part == 0)
        for (int r = 0; r < cc; ++r) Vt[static_cast<size_t>(r) * dp + cc] = 0.0;
      for (int r = 0; r < d; ++r) {
        double sum = 0.0;
        if (active && r >= cc)
          for (int t = cc + part; t < r; t += 4) sum += G[static_cast<size_t>(r) * dp + t] * Vt[static_cast<size_t>(t) * dp + cc];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (active && part == 0 && r >= cc)
          Vt[static_cast<size_t>(r) * dp + cc] = (((r == cc) ? 1.0 : 0.0) - sum) / G[static_cast<size_t>(r) * dp + r];
        __syncwarp();
      }
    }
    __syncthreads();
    // A^-1 = L^-T L^-1
    for (int idx = tid; idx < d * d; idx += kPinvThreads) {
      const int a = idx / d, b = idx % d;
      double acc = 0.0;
      for (int i = (a > b ? a : b); i < d; ++i) acc += Vt[static_cast<size_t>(i) * dp + a] * Vt[static_cast<size_t>(i) * dp + b];
      out[idx] = acc;
    }
    if (tid == 0 && info) { info[0] = 0; info[1] = d; }
    return;
  }
  // ---- general path: restore the working copies and run the one-sided Jacobi SVD
  for (int idx = tid; idx < dp * dp; idx += kPinvThreads) {
    const int i = idx / dp, j = idx % dp;
    G[idx] = (i < d && j < d) ? A[i * d + j] : 0.0;
    Vt[idx] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();

  const int sweep = jacobi_orthogonalise_rows(G, Vt, dp, &s_rot);

  // singular values
  if (tid == 0) s_max = 0.0;
  __syncthreads();
  for (int i = warp; i < dp; i += nwarps) {
    double a = 0.0;
    for (int j = lane; j < dp; j += 32) { const double x = G[static_cast<size_t>(i) * dp + j]; a += x * x; }
    a = warp_sum(a);
    if (lane == 0) sig2[i] = a;
  }
  __syncthreads();
  if (tid == 0) {
    double mx = 0.0;
    for (int i = 0; i < dp; ++i) mx = fmax(mx, sig2[i]);
    s_max = mx;
    if (info) { info[0] = sweep; }
  }
  __syncthreads();
  const double cut = rcond * sqrt(s_max);
  int rank = 0;
  for (int idx = tid; idx < d * d; idx += kPinvThreads) {
    const int a = idx / d, b = idx % d;
    double acc = 0.0;
    for (int i = 0; i < dp; ++i) {
      const double s2 = sig2[i];
      if (sqrt(s2) > cut) acc += Vt[static_cast<size_t>(i) * dp + a] * G[static_cast<size_t>(i) * dp + b] / s2;
    }
    out[idx] = acc;
  }
  if (tid == 0 && info) {
    for (int i = 0; i < dp; ++i) rank += (sqrt(sig2[i]) > cut) ? 1 : 0;
    info[1] = rank;
  }
}

// ------------------------------------------------------------------ small fp64 GEMM with epilogues
// C[i, k] = sum_j A[i, j] * B[j, k]   (A: [d, d] row-major, B/C: [d, K]);  |C| < zero_below -> 0
// (datashapley.py:45 `coef[np.abs(coef) < 1e-10] = 0`).  Fixed j order: deterministic.
constexpr int kDgemmTile = 32;
__global__ void dgemm_dk_kernel(const double* __restrict__ A, const double* __restrict__ B, int64_t d, int64_t K,
                                double zero_below, double* __restrict__ Cout) {
  __shared__ double sA[kDgemmTile][kDgemmTile + 1];
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // blockDim = (32, 8)
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * kDgemmTile;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};  // rows i0 + threadIdx.y + {0, 8, 16, 24}
  for (int64_t j0 = 0; j0 < d; j0 += kDgemmTile) {
    for (int t = threadIdx.y; t < kDgemmTile; t += 8) {
      const int64_t i = i0 + t, j = j0 + threadIdx.x;
      sA[t][threadIdx.x] = (i < d && j < d) ? A[i * d + j] : 0.0;
    }
    __syncthreads();
    const int64_t jmax = (d - j0 < kDgemmTile) ? (d - j0) : kDgemmTile;
    if (k < K) {
      for (int64_t jj = 0; jj < jmax; ++jj) {
        const double b = B[(j0 + jj) * K + k];
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t] += sA[threadIdx.y + 8 * t][jj] * b;
      }
    }
    __syncthreads();
  }
  if (k < K) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int64_t i = i0 + threadIdx.y + 8 * t;
      if (i < d) {
        double v = acc[t];
        if (fabs(v) < zero_below) v = 0.0;
        Cout[i * K + k] = v;
      }
    }
  }
}

// Shapley constraint step (datashapley.py:38-43):
//   colsum = 1^T Ainv ; dd = 1^T Ainv 1 ; c_k = colsum . b[:,k] - v1_k + v0_k ; rhs[:,k] = b[:,k] - c_k / dd
__global__ void shapley_colsum_kernel(const double* __restrict__ Ainv, int64_t d, double* __restrict__ colsum) {
  // one block; colsum[d] holds dd
  for (int64_t j = threadIdx.x; j < d; j += blockDim.x) {
    double a = 0.0;
    for (int64_t i = 0; i < d; ++i) a += Ainv[i * d + j];
    colsum[j] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int64_t j = 0; j < d; ++j) t += colsum[j];
    colsum[d] = t;
  }
}
__global__ void shapley_rhs_kernel(const double* __restrict__ colsum, const double* __restrict__ b, int64_t d, int64_t K,
                                   const double* __restrict__ v1, const double* __restrict__ v0,
                                   double* __restrict__ rhs) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double c = 0.0;
  for (int64_t i = 0; i < d; ++i) c += colsum[i] * b[i * K + k];
  c = c - v1[k] + v0[k];
  const double corr = c / colsum[d];
  for (int64_t i = 0; i < d; ++i) rhs[i * K + k] = b[i * K + k] - corr;
}

// ------------------------------------------------------------------ Spearman / LDS
// rho[e, k] = Spearman( pred[idx[e, :], k], y[idx[e, :], k] )  (average ranks for ties; NaN when a
// side is constant or holds a NaN -- scipy.stats.spearmanr semantics).  idx == nullptr: identity over the m rows.
// One warp per (e, k); m_r <= kLdsMaxRows rows (round 1: 1024).
//
// Ranks by sorting (round 1 counted "how many are smaller" for every element: O(m^2) comparisons per job): the warp
// runs a bitonic network over (value, original index) keys in shared memory -- log2(mp) (log2(mp) + 1) / 2 stages of
// mp / 2 compare-exchanges, mp = m rounded up to a power of two, padded with +inf keys that sort last -- then every
// sorted position finds the ends of its run of equal values by two binary searches and scatters twice its average
// rank to the element's original slot.  Twice-ranks are integers <= 2 m, and the three sums of products below are
// sums of multiples of 1/4 far below 2^53, hence exact in any order: the result is bit-identical to the counting
// kernel's.
constexpr int kLdsMaxRows = 8192;  // sorting form: 14 bytes of shared memory per (padded) row and warp -> 112 KiB at the cap
__host__ __device__ inline int64_t lds_pow2(int64_t m) { int64_t p = 32; while (p < m) p <<= 1; return p; }
__host__ __device__ inline size_t lds_warp_smem_bytes(int64_t mr) {  // keys [mp] f64, order [mp] u16, twice-ranks 2 x [mp] u16
  return static_cast<size_t>(lds_pow2(mr)) * (8 + 2 + 2 + 2);
}

// sorts key[0..mp) ascending by (key, ord); all 32 lanes participate
__device__ __forceinline__ void warp_bitonic_sort(double* key, uint16_t* ord, int mp, int lane) {
  for (int k = 2; k <= mp; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (mp >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const double a = key[i], b = key[l];
        const uint16_t oa = ord[i], ob = ord[l];
        const bool gt = (a > b) || (a == b && oa > ob);
        if (gt == ((i & k) == 0)) {  // ascending block: swap when out of order; descending block: the opposite
          key[i] = b; key[l] = a;
          ord[i] = ob; ord[l] = oa;
        }
      }
      __syncwarp();
    }
  }
}

// twice the average 1-based rank of every element, written to rank2[original index]; key / ord sorted, m real entries
__device__ __forceinline__ void warp_average_ranks(const double* key, const uint16_t* ord, int m, uint16_t* rank2, int lane) {
  for (int p = lane; p < m; p += 32) {
    const double v = key[p];
    int lo = 0, hi = p;  // first position holding v
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (key[mid] < v) lo = mid + 1; else hi = mid; }
    const int first = lo;
    lo = p; hi = m - 1;  // last position holding v
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (key[mid] > v) hi = mid - 1; else lo = mid; }
    rank2[ord[p]] = static_cast<uint16_t>(first + lo + 2);  // 2 * ((first + 1) + (last + 1)) / 2
  }
  __syncwarp();
}

__global__ void lds_spearman_kernel(const double* __restrict__ pred, const double* __restrict__ y, int64_t m, int64_t K,
                                    const int32_t* __restrict__ idx, int64_t R, int64_t mr, double* __restrict__ rho) {
  extern __shared__ __align__(16) unsigned char lds_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t job = static_cast<int64_t>(blockIdx.x) * nw + warp;
  if (job >= R * K) return;
  const int64_t e = job / K, k = job % K;
  const int mp = static_cast<int>(lds_pow2(mr)), mi = static_cast<int>(mr);
  unsigned char* base = lds_smem + static_cast<size_t>(warp) * lds_warp_smem_bytes(mr);
  double* key = reinterpret_cast<double*>(base);
  uint16_t* ord = reinterpret_cast<uint16_t*>(base + static_cast<size_t>(mp) * 8);
  uint16_t* ra2 = ord + mp;
  uint16_t* rb2 = ra2 + mp;
  bool has_nan = false;
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const double* src = side == 0 ? pred : y;
    for (int r = lane; r < mp; r += 32) {
      double v = __longlong_as_double(0x7ff0000000000000ll);  // +inf padding, original index >= m: sorts last
      if (r < mi) {
        const int64_t row = idx ? idx[e * mr + r] : r;
        v = src[row * K + k];
        has_nan |= (v != v);
      }
      key[r] = v;
      ord[r] = static_cast<uint16_t>(r);
    }
    __syncwarp();
    warp_bitonic_sort(key, ord, mp, lane);
    warp_average_ranks(key, ord, mi, side == 0 ? ra2 : rb2, lane);
  }
  const double mean2 = static_cast<double>(mr) + 1.0;  // twice the mean rank
  double sab = 0.0, saa = 0.0, sbb = 0.0;
  for (int r = lane; r < mi; r += 32) {
    const double ra = 0.5 * (static_cast<double>(ra2[r]) - mean2);  // average rank minus the mean rank (exact)
    const double rb = 0.5 * (static_cast<double>(rb2[r]) - mean2);
    sab += ra * rb; saa += ra * ra; sbb += rb * rb;
  }
  sab = warp_sum(sab); saa = warp_sum(saa); sbb = warp_sum(sbb);
  has_nan = __any_sync(0xffffffffu, has_nan);
  if (lane == 0)  // corrcoef's two-step normalisation; 0/0 -> NaN
    rho[job] = has_nan ? __longlong_as_double(0x7ff8000000000000ll) : (sab / sqrt(saa)) / sqrt(sbb);
}

// Small evaluation sets (m <= kLdsCountRows): ranks by counting "how many are smaller / equal" -- m^2 / 32 steps of
// broadcast reads per lane, which beats the sorting network's shared-memory round trips and warp barriers up to
// m ~ 128 (measured at m = 100, K = 2e5: 1.39 ms counting vs 1.68 ms sorting; at m = 1024 the sort is ~7x ahead).
// Same exact sums, same result bit for bit.
constexpr int kLdsCountRows = 128;
__global__ void lds_spearman_count_kernel(const double* __restrict__ pred, const double* __restrict__ y, int64_t m, int64_t K,
                                    const int32_t* __restrict__ idx, int64_t R, int64_t mr, double* __restrict__ rho) {
  extern __shared__ __align__(16) unsigned char lds_smem_raw[];
  double* lds_smem = reinterpret_cast<double*>(lds_smem_raw);  // [warps][2][mr]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t job = static_cast<int64_t>(blockIdx.x) * nw + warp;
  if (job >= R * K) return;
  const int64_t e = job / K, k = job % K;
  double* sa = lds_smem + static_cast<size_t>(warp) * 2 * mr;
  double* sb = sa + mr;
  for (int64_t r = lane; r < mr; r += 32) {
    const int64_t row = idx ? idx[e * mr + r] : r;
    sa[r] = pred[row * K + k];
    sb[r] = y[row * K + k];
  }
  __syncwarp();
  bool has_nan = false;
  for (int64_t r = lane; r < mr; r += 32) has_nan |= (sa[r] != sa[r]) || (sb[r] != sb[r]);
  has_nan = __any_sync(0xffffffffu, has_nan);
  const double mean = 0.5 * (static_cast<double>(mr) + 1.0);
  double sab = 0.0, saa = 0.0, sbb = 0.0;
  for (int64_t r = lane; r < mr; r += 32) {
    const double a = sa[r], b = sb[r];
    int la = 0, ea = 0, lb = 0, eb = 0;
    for (int64_t t = 0; t < mr; ++t) {
      const double at = sa[t], bt = sb[t];
      la += (at < a); ea += (at == a);
      lb += (bt < b); eb += (bt == b);
    }
    const double ra = la + 0.5 * (ea + 1) - mean;  // average rank (1-based) minus the mean rank
    const double rb = lb + 0.5 * (eb + 1) - mean;
    sab += ra * rb; saa += ra * ra; sbb += rb * rb;
  }
  sab = warp_sum(sab); saa = warp_sum(saa); sbb = warp_sum(sbb);
  if (lane == 0)  // corrcoef's two-step normalisation; 0/0 -> NaN
    rho[job] = has_nan ? __longlong_as_double(0x7ff8000000000000ll) : (sab / sqrt(saa)) / sqrt(sbb);
}

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h, pairwise_sum_@TYPE@), restated over a stream of
// values consumed in order: fewer than 8 -> running sum from -0; up to 128 -> eight strided accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus the remainder in order; more -> split at n/2 rounded down to a multiple of
// 8 and recurse.  np.sum / np.mean over a contiguous axis are 0 + this (pinned against numpy for every n <= 300 and
// a set of larger sizes by tests/test_oracle_golden.py).  The reference ranks contributors by
// np.argsort(-x.mean(-1)) and sums group attributions with .sum() (traks.py:199-218, shapley_lds.py:294): the last
// ulp of these sums decides near-ties, so the order is reproduced exactly instead of "some fixed order".
template <typename T, typename Next>
__device__ T numpy_pairwise_sum(Next& next, int64_t n) {
  if (n < 8) {
    T res = static_cast<T>(-0.0);
    for (int64_t i = 0; i < n; ++i) res = res + next();
    return res;
  }
  if (n <= 128) {
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = next();
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = r[j] + next();
    }
    T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res = res + next();
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  const T left = numpy_pairwise_sum<T>(next, n2);
  return left + numpy_pairwise_sum<T>(next, n - n2);
}

// lds[e] = mean_k(rho[e, k] * 100) = np.mean of the per-behaviour list (shapley_lds.py:147, lds.py:167)
__global__ void lds_mean_kernel(const double* __restrict__ rho, int64_t R, int64_t K, double* __restrict__ out) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= R) return;
  const double* p = rho + e * K;
  auto next = [&]() { return __dmul_rn(*p++, 100.0); };  // rounded product, then summed (no FMA contraction)
  out[e] = (0.0 + numpy_pairwise_sum<double>(next, K)) / static_cast<double>(K);
}

// ------------------------------------------------------------------ group reductions and ranks
// out[g] = sum / mean / max over {values[i] : group[i] == g} exactly as the reference computes them
// (traks.py:188-204 `attrs[group_indices].sum()` / `.mean()` / `.max()`; attribution_utils.py:15-48): members in index
// order, numpy's pairwise summation IN THE VALUES' OWN PRECISION (the reference's attrs are float32 arrays, so its
// group sums are float32 sums widened afterwards), mean = sum / count in that precision.  One thread per group:
// count the members, then stream them through numpy_pairwise_sum (N * G reads from L2; N <= 5 * 10^4, G <= 10^3).
template <typename T>
__global__ void group_reduce_kernel(const T* __restrict__ values, const int32_t* __restrict__ group, int64_t N,
                                    int64_t G, int mode, double* __restrict__ out) {
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= G) return;
  int64_t cnt = 0;
  T mx = static_cast<T>(-INFINITY);
  bool any_nan = false;
  for (int64_t i = 0; i < N; ++i) {
    if (group[i] == g) {
      const T v = values[i];
      ++cnt;
      any_nan |= (v != v);
      mx = v > mx ? v : mx;
    }
  }
  if (mode == 2) {
    out[g] = any_nan ? static_cast<double>(NAN) : static_cast<double>(mx);  // np.max propagates NaN
    return;
  }
  int64_t cursor = 0;
  auto next = [&]() {
    while (group[cursor] != g) ++cursor;
    return values[cursor++];
  };
  const T sum = static_cast<T>(0) + numpy_pairwise_sum<T>(next, cnt);
  out[g] = static_cast<double>(mode == 0 ? sum : sum / static_cast<T>(cnt));  // empty group: 0 / (0/0 = NaN) like numpy
}

// rank[pos] = i where pos = #{j : x_j > x_i} + #{j < i : x_j == x_i}  == np.argsort(-x, kind="stable")
// (NaNs sort last, in index order, as numpy does).
__global__ void stable_rank_desc_kernel(const double* __restrict__ x, int64_t n, int64_t* __restrict__ rank) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double xi = x[i];
  const bool nan_i = isnan(xi);
  int64_t pos = 0;
  for (int64_t j = 0; j < n; ++j) {
    const double xj = x[j];
    const bool nan_j = isnan(xj);
    bool before;
    if (nan_i) before = !nan_j || j < i;
    else before = !nan_j && (xj > xi || (xj == xi && j < i));
    pos += before ? 1 : 0;
  }
  rank[pos] = i;
}

// row means of a [n, K] fp64 matrix (attrs_all.mean(axis=-1) before ranking, shapley_lds.py:294)
__global__ void row_mean_kernel(const double* __restrict__ x, int64_t n, int64_t K, double* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* p = x + i * K;
  auto next = [&]() { return *p++; };
  out[i] = (0.0 + numpy_pairwise_sum<double>(next, K)) / static_cast<double>(K);  // == np.mean(x, axis=-1)
}

}  // namespace agg
}  // namespace gadm
