// One right-hand side through the Cholesky factor: x = (L L^T)^-1 b.
//
// The mean-first TRAK score (traks.py:152-157 with the mean over generated images taken first) needs K^-1 applied to
// ONE row.  Building the explicit triangular inverse for that (gadm_tri_inverse: 14 batched GEMM launches, 0.7 ms at
// k = 4096) is 20x the arithmetic of two substitutions; but a substitution is 2 * k / 128 strictly dependent steps,
// and as separate launches each step costs a launch latency.  Here the whole solve is one cooperative launch of
// k / 128 CTAs with point-to-point signalling:
//
//   CTA c owns block c of the solution.  Forward (L y = b): once the partial products p[c][c'] = L[c, c'] y_c'
//   of every c' < c have arrived it forms rhs = b_c - sum_c' p[c][c'] (fixed order, fp64), y_c = Linv_c rhs (the
//   128 x 128 inverse of the diagonal factor block that potrf_diag_kernel left in the workspace), and then streams
//   the blocks L[j, c], j = c+1 .., below its diagonal block through shared memory (cp.async, double-buffered,
//   nearest row block first -- that one is on the critical path of CTA c+1), publishing p[j][c] and bumping row j's
//   arrival counter.  Backward (L^T x = y) mirrors it with the blocks L[c, j], j = c-1 .. 0, of its own block row,
//   used transposed.
//
// Per step the critical path is one 128 x 128 matrix-vector product from shared memory plus one signal (fence +
// atomic + polling load), ~3 us; every block of L is read exactly once (k^2 / 2 floats in each direction).
// Deterministic: fixed summation orders everywhere.  Requires a 16-byte aligned L with ld % 4 == 0 and all
// ceil(k / 128) CTAs co-resident (cooperative launch); k need not be a multiple of 128 (the last block is padded).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gadm_ptx.cuh"

namespace gadm {
namespace trsv {

constexpr int kB = 128;          // block size = potrf block size
constexpr int kThreads = 256;    // 8 warps
constexpr int kSmemBytes = 2 * kB * kB * 4 + 3 * kB * 4;

__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// 128 x 128 block (row pitch ld floats) -> shared memory [128][128], asynchronously, by the whole CTA; rows >= rows_valid
// (the last block row of a system whose size is not a multiple of 128) are written as zeros
__device__ __forceinline__ void fetch_block(float* dst, const float* src, int64_t ld, int rows_valid = kB) {
  const uint32_t d0 = smem_u32(dst);
  for (int i = threadIdx.x; i < kB * kB / 4; i += kThreads) {
    const int r = i >> 5, q = i & 31;  // row, 16-byte chunk
    if (r < rows_valid) cp16(d0 + (r * kB + q * 4) * 4, src + static_cast<int64_t>(r) * ld + q * 4);
    else *reinterpret_cast<float4*>(dst + r * kB + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_commit();
}

// out[r] = sum_s M[r][s] v[s] (kTrans = false) or sum_s M[s][r] v[s] (kTrans = true), M [128][128] in shared memory.
// Row form: a warp owns rows w, w + 8, ..., lanes run along s (conflict-free), shuffle reduction.  Column form:
// thread r (< 128) walks down column r (consecutive threads, consecutive addresses).
template <bool kTrans>
__device__ __forceinline__ void block_matvec(const float* M, const float* v, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (!kTrans) {
    const float v0 = v[lane], v1 = v[lane + 32], v2 = v[lane + 64], v3 = v[lane + 96];
    for (int r = warp; r < kB; r += kThreads / 32) {
      const float* row = M + r * kB;
      float s = fmaf(row[lane], v0, fmaf(row[lane + 32], v1, fmaf(row[lane + 64], v2, row[lane + 96] * v3)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) out[r] = s;
    }
  } else if (threadIdx.x < kB) {
    const int r = threadIdx.x;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
    for (int s = 0; s < kB; s += 4) {
      s0 = fmaf(M[(s + 0) * kB + r], v[s + 0], s0);
      s1 = fmaf(M[(s + 1) * kB + r], v[s + 1], s1);
      s2 = fmaf(M[(s + 2) * kB + r], v[s + 2], s2);
      s3 = fmaf(M[(s + 3) * kB + r], v[s + 3], s3);
    }
    out[r] = (s0 + s1) + (s2 + s3);
  }
}

__device__ __forceinline__ void wait_count(const uint32_t* counter, uint32_t target, uint32_t code) {
  if (threadIdx.x == 0) {
    const uint64_t t0 = globaltimer_ns();
    const uint64_t limit = *reinterpret_cast<volatile unsigned long long*>(&g_watchdog_ns);
    uint32_t v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if (limit != 0 && globaltimer_ns() - t0 > limit) watchdog_fire(code);
    } while (true);
  }
  __syncthreads();
}

__device__ __forceinline__ void signal(uint32_t* counter) {
  __syncthreads();  // every thread's partial products are written
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
  }
}

// grid = ceil(k / 128) CTAs (cooperative).  partial: [2][nblk][nblk][128] doubles; count: [2][nblk] words, zeroed.
// k % 128 != 0: the last block is padded -- potrf_diag_kernel pads its diagonal block with the identity, the padded
// rows of L read as zeros here, b as zero, and the padded part of x is not stored.
__global__ void __launch_bounds__(kThreads)
chol_solve_vec_kernel(const float* __restrict__ L, int64_t ld, int k, const float* __restrict__ linv,
                      const float* __restrict__ linv_t, const float* __restrict__ b, float* __restrict__ x,
                      double* __restrict__ partial, uint32_t* __restrict__ count) {
  extern __shared__ __align__(16) float trsv_smem[];
  float* buf0 = trsv_smem;
  float* buf1 = trsv_smem + kB * kB;
  float* rhs = trsv_smem + 2 * kB * kB;
  float* vec = rhs + kB;   // y_c, later x_c
  float* prod = vec + kB;  // one block product
  const int c = blockIdx.x, tid = threadIdx.x;
  const int nblk = (k + kB - 1) / kB;
  const int last_rows = k - (nblk - 1) * kB;  // valid rows of the last block
  auto rows_of = [&](int blk) { return blk == nblk - 1 ? last_rows : kB; };
  double* part_f = partial;
  double* part_b = partial + static_cast<size_t>(nblk) * nblk * kB;
  uint32_t* cnt_f = count;
  uint32_t* cnt_b = count + nblk;

  // ---------------- forward: L y = b
  fetch_block(buf0, linv + static_cast<size_t>(c) * kB * kB, kB);  // before the wait: it does not depend on anything
  if (c + 1 < nblk) fetch_block(buf1, L + static_cast<int64_t>(c + 1) * kB * ld + static_cast<int64_t>(c) * kB, ld, rows_of(c + 1));
  wait_count(cnt_f + c, static_cast<uint32_t>(c), 0x7501);
  if (tid < kB) {
    double s = tid < rows_of(c) ? static_cast<double>(b[c * kB + tid]) : 0.0;
    for (int j = 0; j < c; ++j) s -= __ldcg(part_f + (static_cast<size_t>(c) * nblk + j) * kB + tid);
    rhs[tid] = static_cast<float>(s);
  }
  if (c + 1 < nblk) cp_wait<1>(); else cp_wait<0>();
  __syncthreads();
  block_matvec<false>(buf0, rhs, vec);  // y_c = Linv_c rhs
  __syncthreads();
  for (int j = c + 1; j < nblk; ++j) {
    float* cur = ((j - c) & 1) ? buf1 : buf0;
    float* nxt = ((j - c) & 1) ? buf0 : buf1;
    if (j + 1 < nblk) {
      fetch_block(nxt, L + static_cast<int64_t>(j + 1) * kB * ld + static_cast<int64_t>(c) * kB, ld, rows_of(j + 1));
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    block_matvec<false>(cur, vec, prod);  // L[j, c] y_c
    __syncthreads();
    if (tid < kB) part_f[(static_cast<size_t>(j) * nblk + c) * kB + tid] = static_cast<double>(prod[tid]);
    signal(cnt_f + j);
  }

  // ---------------- backward: L^T x = y   (vec holds y_c)
  __syncthreads();
  fetch_block(buf0, linv_t + static_cast<size_t>(c) * kB * kB, kB);
  if (c > 0) fetch_block(buf1, L + static_cast<int64_t>(c) * kB * ld + static_cast<int64_t>(c - 1) * kB, ld, rows_of(c));
  wait_count(cnt_b + c, static_cast<uint32_t>(nblk - 1 - c), 0x7502);
  if (tid < kB) {
    double s = static_cast<double>(vec[tid]);
    for (int j = nblk - 1; j > c; --j) s -= __ldcg(part_b + (static_cast<size_t>(c) * nblk + j) * kB + tid);
    rhs[tid] = static_cast<float>(s);
  }
  if (c > 0) cp_wait<1>(); else cp_wait<0>();
  __syncthreads();
  block_matvec<false>(buf0, rhs, vec);  // x_c = Linv_c^T rhs  (linv_t holds the transposed inverse)
  __syncthreads();
  if (tid < rows_of(c)) x[c * kB + tid] = vec[tid];
  for (int j = c - 1; j >= 0; --j) {
    float* cur = ((c - j) & 1) ? buf1 : buf0;
    float* nxt = ((c - j) & 1) ? buf0 : buf1;
    if (j > 0) {
      fetch_block(nxt, L + static_cast<int64_t>(c) * kB * ld + static_cast<int64_t>(j - 1) * kB, ld, rows_of(c));
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    block_matvec<true>(cur, vec, prod);  // L[c, j]^T x_c
    __syncthreads();
    if (tid < kB) part_b[(static_cast<size_t>(j) * nblk + c) * kB + tid] = static_cast<double>(prod[tid]);
    signal(cnt_b + j);
  }
}

}  // namespace trsv
}  // namespace gadm
