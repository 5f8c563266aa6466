// fp32-grade GEMM on the 5th-gen tensor cores:  C[M, N] = alpha * A[M, K] * B[N, K]^T + beta * C
// A and B are plain fp32, row-major with the contraction index contiguous ("TN" form, both K-major).
//
// Used for every dense product of the TRAK scorer (text_to_image/traks.py:141-186 torch.matmul calls;
// src/attributions/methods/compute_gradient_score.py:75-79,108-126): Gram Phi^T Phi, Cholesky trailing
// updates and triangular solves, Z = Phi_gen K^-1 and the score GEMM S = Z Phi^T.
//
// Precision: the reference computes these in fp32 (fp64 for the unconditional inverse), so a single
// TF32 pass (10-bit mantissa) is not enough.  Each operand tile is split in shared memory into
// hi = x & 0xffffe000 (exactly representable in TF32) and lo = x - hi (exact in fp32), and the MMA warp
// issues three tcgen05.mma.kind::tf32 per K-step: hi*hi + hi*lo + lo*hi (the lo*lo term, <= 2^-20
// relative, is dropped) -- "3xTF32", fp32 accumulation in TMEM.
//
// Pipeline per CTA (one 128 x 128 output tile, 3 stages of K = 32):
//   warp 0     TMA: raw fp32 tiles of A and B (128B swizzle)              -> raw_full[s]
//   warps 12-15 split raw -> hi (in place) + lo, fence.proxy.async        -> split_full[s]
//   warp 1     MMA issue (one thread), tcgen05.commit                     -> empty[s], acc_full[buf]
//   warps 4-11 promote each 128-element K chunk TMEM -> fp32 registers    -> acc_empty[buf];
//              at the end alpha/beta/diag-shift -> global
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "gadm_ptx.cuh"

namespace gadm {
namespace gemm {

constexpr int kBM = 128;
constexpr int kBN = 128;
constexpr int kBK = 32;     // fp32 elements per smem row = 128 B
constexpr int kUmmaK = 8;   // tf32
constexpr int kStages = 3;
constexpr int kChunkKB = 4; // k-blocks (128 contraction elements) accumulated in TMEM before promotion
constexpr int kTileBytes = kBM * kBK * 4;                 // 16 KiB (A and B tiles have the same shape)
constexpr int kStageBytes = 4 * kTileBytes;               // rawA(hi) | rawB(hi) | loA | loB
constexpr int kEpiWarps = 8;                              // 2 per TMEM lane quarter, 64 columns each
constexpr int kSplitWarps = 4;
constexpr int kFirstEpiWarp = 4;
constexpr int kFirstSplitWarp = kFirstEpiWarp + kEpiWarps;
constexpr int kThreads = (kFirstSplitWarp + kSplitWarps) * 32;
constexpr int kTmemCols = 256;                            // two 128-column accumulator buffers
constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;

struct Args {
  float* C;
  int64_t ldc;
  int64_t stride_c;   // elements between the C matrices of consecutive batch entries (blockIdx.z)
  int32_t M, N, K;
  float alpha, beta;
  float diag_add;     // added to C[i, i] (global indices, after alpha/beta)
  int32_t lower_only; // skip tiles entirely above the diagonal
  int32_t tri_b;      // 1: B is lower-triangular (B[j, c] = 0 for c > j), 2: upper-triangular (0 for c < j):
                      //    the contraction of output tile column tile_n stops / starts at its diagonal block
};

// contraction k-block range [kb0, kb1) of an output tile
__device__ __forceinline__ void kblock_range(const Args& a, int tile_n, int& kb0, int& kb1) {
  const int nkb = (a.K + kBK - 1) / kBK;
  kb0 = 0;
  kb1 = nkb;
  if (a.tri_b == 1) { const int e = (tile_n + 1) * (kBN / kBK); kb1 = e < nkb ? e : nkb; }
  else if (a.tri_b == 2) { const int b = tile_n * (kBN / kBK); kb0 = b < nkb ? b : nkb; }
}

__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
  hi = x & 0xFFFFE000u;
  lo = __float_as_uint(__fsub_rn(__uint_as_float(x), __uint_as_float(hi)));
}

// Why chunks: the tensor core adds into its fp32 accumulator with truncation, so a long same-sign sum
// (the diagonal of a Gram matrix is one) drifts by ~(steps/2) * 2^-24 relative -- 5e-4 at 50 000
// examples, far above fp32 GEMM accuracy.  The MMA warp therefore accumulates only kChunkKB k-blocks in
// TMEM, ping-ponging between two accumulator buffers, and the epilogue warps promote every finished chunk
// into fp32 registers with round-to-nearest adds while the next chunk is being multiplied.
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_3xtf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  if (a.lower_only && tile_n * kBN > tile_m * kBM + (kBM - 1)) return;  // whole CTA exits together

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto raw_full = [&](int s) { return bar_base + 8u * s; };
  auto split_full = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto acc_full = [&](int b) { return bar_base + 8u * (3 * kStages + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (3 * kStages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * kStages + 4);
  auto hi_a = [&](int s) { return smem_base + s * kStageBytes; };
  auto hi_b = [&](int s) { return smem_base + s * kStageBytes + kTileBytes; };
  auto lo_a = [&](int s) { return smem_base + s * kStageBytes + 2 * kTileBytes; };
  auto lo_b = [&](int s) { return smem_base + s * kStageBytes + 3 * kTileBytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int kb0, kb1;
  kblock_range(a, tile_n, kb0, kb1);
  const int nkb = kb1 - kb0;  // >= 1: the diagonal block is always inside the range
  const int nchunks = (nkb + kChunkKB - 1) / kChunkKB;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tmap_a); prefetch_tensormap(&tmap_b); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(raw_full(s), 1);
      mbar_init(split_full(s), kSplitWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u, 0x1100 + s);
        mbar_arrive_expect_tx(raw_full(s), 2 * kTileBytes);
        tma_load_3d(hi_a(s), &tmap_a, raw_full(s), (kb0 + kb) * kBK, tile_m * kBM, blockIdx.z);
        tma_load_3d(hi_b(s), &tmap_b, raw_full(s), (kb0 + kb) * kBK, tile_n * kBN, blockIdx.z);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(UMMA_FMT_TF32, kBM, kBN);
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        mbar_wait(acc_empty(buf), ((c >> 1) & 1u) ^ 1u, 0x1500 + buf);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        const int kb_end = (c + 1) * kChunkKB < nkb ? (c + 1) * kChunkKB : nkb;
        for (int kb = c * kChunkKB; kb < kb_end; ++kb) {
          const int s = kb % kStages;
          const uint32_t ph = (kb / kStages) & 1u;
          mbar_wait(split_full(s), ph, 0x1200 + s);
          tcgen05_fence_after();
          const uint64_t dha = umma_desc_kmajor_sw128(hi_a(s)), dhb = umma_desc_kmajor_sw128(hi_b(s));
          const uint64_t dla = umma_desc_kmajor_sw128(lo_a(s)), dlb = umma_desc_kmajor_sw128(lo_b(s));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint32_t first = (kb == c * kChunkKB && k == 0) ? 0u : 1u;
            umma_tf32(d_tmem, dla + 2u * k, dhb + 2u * k, idesc, first);  // small terms first
            umma_tf32(d_tmem, dha + 2u * k, dlb + 2u * k, idesc, 1u);
            umma_tf32(d_tmem, dha + 2u * k, dhb + 2u * k, idesc, 1u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(acc_full(buf));
      }
    }
  } else if (warp >= kFirstEpiWarp && warp < kFirstSplitWarp) {
    const int e = warp - kFirstEpiWarp;
    const int q = warp & 3;       // TMEM lane quarter this warp may read
    const int half = e >> 2;      // which 64 of the 128 accumulator columns
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(acc_full(buf), (c >> 1) & 1u, 0x1300 + buf);
      tcgen05_fence_after();
      const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kBN + half * 64;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t0 + j * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[j * 32 + i] = __fadd_rn(acc[j * 32 + i], __uint_as_float(v[i]));
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
    }
    const int64_t row = static_cast<int64_t>(tile_m) * kBM + q * 32 + lane;
    const int64_t col0 = static_cast<int64_t>(tile_n) * kBN + half * 64;
    if (row < a.M) {
      float* crow = a.C + static_cast<int64_t>(blockIdx.z) * a.stride_c + row * a.ldc;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const int64_t col = col0 + i;
        if (col < a.N) {
          float r = a.alpha * acc[i];
          if (a.beta != 0.f) r += a.beta * crow[col];
          if (col == row) r += a.diag_add;
          crow[col] = r;
        }
      }
    }
  } else if (warp >= kFirstSplitWarp) {
    const int t = threadIdx.x - kFirstSplitWarp * 32;  // 0..127
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1u;
      mbar_wait(raw_full(s), ph, 0x1400 + s);
      // raw A|B are contiguous (2 * kTileBytes), lo A|B follow at +2*kTileBytes: purely elementwise, the
      // swizzle is a function of the address bits inside each 1024-B atom and is identical for hi and lo.
      const uint32_t raw = hi_a(s);
#pragma unroll 8
      for (int i = 0; i < (2 * kTileBytes) / 16 / (kSplitWarps * 32); ++i) {
        const uint32_t off = (static_cast<uint32_t>(i) * (kSplitWarps * 32) + t) * 16u;
        const uint4 x = ld_shared_v4(raw + off);
        uint4 hi, lo;
        split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y);
        split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
        // hi is NOT written back: kind::tf32 reads the fp32 bit pattern and ignores the low 13 mantissa bits, so the
        // raw tile already is the hi operand.  Shared-memory bandwidth is this kernel's bound (a 128x128x8 tf32 UMMA
        // reads 8 KiB per 64 clocks = 128 B/clk, the whole budget, and the split competes with it): one store less
        // per element takes the splitter's traffic from 96 to 64 KiB per k-block.
        st_shared_v4(raw + 2 * kTileBytes + off, lo);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(split_full(s));
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ "TS" variant: A operand in tensor memory
// Shared-memory bandwidth bounds the kernel above: per 32-wide k-block the MMAs read 12 x 8 KiB of operands, the
// splitter moves 64 KiB and TMA writes 32 KiB -- 192 KiB at 128 B/clk = 1536 clocks for 768 clocks of tf32 math
// (ncu: tensor pipe 43-50 %).  Here the splitter warps put the A tile (hi = raw and lo) into TENSOR memory with
// tcgen05.st (lane = row, 32 columns per k-block each) and the MMAs take A from there
// (tcgen05.mma ... [d], [a_tmem], b_desc): operand reads from smem halve (48 KiB), the lo-A smem tile and its store
// disappear (splitter 48 KiB), a stage shrinks to 48 KiB (4 stages) -- 128 KiB per k-block = 1024 clocks, a 75 % bound.
// Measured on the config-2 Gram (50 000 x 4096, lower tiles): tensor pipe 43 % -> 61 % of elapsed cycles (68 % inside
// the 3.57 -> 4 wave quantisation), 4.63 -> 4.17 ms; L2 at 27 % and LSU smem wavefronts at 43 % are not the limit any
// more -- what is left is the TMA -> split -> MMA round trip against only four stages in flight.
// TMEM map (512 columns): [0, 256) two accumulator buffers, [256 + 64 s, +32) hi-A and [+32, +64) lo-A of stage s.
constexpr int kTsStages = 4;
constexpr int kTsStageBytes = 3 * kTileBytes;  // rawA | rawB (= hi B, read in place) | lo B
constexpr int kTsSmemBytes = kTsStages * kTsStageBytes + 256 + 1024;
constexpr int kTsTmemCols = 512;
constexpr int kTsACol0 = 256;

__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_3xtf32_ts_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  if (a.lower_only && tile_n * kBN > tile_m * kBM + (kBM - 1)) return;  // whole CTA exits together

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kTsStages * kTsStageBytes;
  auto raw_full = [&](int s) { return bar_base + 8u * s; };
  auto split_full = [&](int s) { return bar_base + 8u * (kTsStages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kTsStages + s); };
  auto acc_full = [&](int b) { return bar_base + 8u * (3 * kTsStages + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (3 * kTsStages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * kTsStages + 4);
  auto raw_a = [&](int s) { return smem_base + s * kTsStageBytes; };
  auto hi_b = [&](int s) { return smem_base + s * kTsStageBytes + kTileBytes; };
  auto lo_b = [&](int s) { return smem_base + s * kTsStageBytes + 2 * kTileBytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int kb0, kb1;
  kblock_range(a, tile_n, kb0, kb1);
  const int nkb = kb1 - kb0;  // >= 1: the diagonal block is always inside the range
  const int nchunks = (nkb + kChunkKB - 1) / kChunkKB;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tmap_a); prefetch_tensormap(&tmap_b); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kTsStages; ++s) {
      mbar_init(raw_full(s), 1);
      mbar_init(split_full(s), kSplitWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, kTsTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kTsStages;
        const uint32_t ph = (kb / kTsStages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u, 0x1900 + s);
        mbar_arrive_expect_tx(raw_full(s), 2 * kTileBytes);
        tma_load_3d(raw_a(s), &tmap_a, raw_full(s), (kb0 + kb) * kBK, tile_m * kBM, blockIdx.z);
        tma_load_3d(hi_b(s), &tmap_b, raw_full(s), (kb0 + kb) * kBK, tile_n * kBN, blockIdx.z);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(UMMA_FMT_TF32, kBM, kBN);
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        mbar_wait(acc_empty(buf), ((c >> 1) & 1u) ^ 1u, 0x1d00 + buf);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        const int kb_end = (c + 1) * kChunkKB < nkb ? (c + 1) * kChunkKB : nkb;
        for (int kb = c * kChunkKB; kb < kb_end; ++kb) {
          const int s = kb % kTsStages;
          const uint32_t ph = (kb / kTsStages) & 1u;
          mbar_wait(split_full(s), ph, 0x1a00 + s);
          tcgen05_fence_after();
          const uint64_t dhb = umma_desc_kmajor_sw128(hi_b(s)), dlb = umma_desc_kmajor_sw128(lo_b(s));
          const uint32_t a_hi = tmem_base + kTsACol0 + s * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint32_t first = (kb == c * kChunkKB && k == 0) ? 0u : 1u;
            umma_tf32_ts(d_tmem, a_lo + k * kUmmaK, dhb + 2u * k, idesc, first);  // small terms first
            umma_tf32_ts(d_tmem, a_hi + k * kUmmaK, dlb + 2u * k, idesc, 1u);
            umma_tf32_ts(d_tmem, a_hi + k * kUmmaK, dhb + 2u * k, idesc, 1u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(acc_full(buf));
      }
    }
  } else if (warp >= kFirstEpiWarp && warp < kFirstSplitWarp) {
    const int e = warp - kFirstEpiWarp;
    const int q = warp & 3;       // TMEM lane quarter this warp may read
    const int half = e >> 2;      // which 64 of the 128 accumulator columns
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(acc_full(buf), (c >> 1) & 1u, 0x1b00 + buf);
      tcgen05_fence_after();
      const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kBN + half * 64;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t0 + j * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[j * 32 + i] = __fadd_rn(acc[j * 32 + i], __uint_as_float(v[i]));
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
    }
    const int64_t row = static_cast<int64_t>(tile_m) * kBM + q * 32 + lane;
    const int64_t col0 = static_cast<int64_t>(tile_n) * kBN + half * 64;
    if (row < a.M) {
      float* crow = a.C + static_cast<int64_t>(blockIdx.z) * a.stride_c + row * a.ldc;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const int64_t col = col0 + i;
        if (col < a.N) {
          float r = a.alpha * acc[i];
          if (a.beta != 0.f) r += a.beta * crow[col];
          if (col == row) r += a.diag_add;
          crow[col] = r;
        }
      }
    }
  } else if (warp >= kFirstSplitWarp) {
    const int q = warp & 3;                            // TMEM lane quarter this warp may write (kFirstSplitWarp % 4 == 0)
    const int t = threadIdx.x - kFirstSplitWarp * 32;  // 0..127
    const int row = q * 32 + lane;                     // row of the A tile owned by this thread
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kTsStages;
      const uint32_t ph = (kb / kTsStages) & 1u;
      mbar_wait(raw_full(s), ph, 0x1c00 + s);
      // ---- A: this thread's 128-byte row (8 swizzled 16-byte chunks) -> hi (= raw) and lo columns in TMEM
      {
        uint32_t x[32], lo[32];
        const uint32_t rbase = raw_a(s) + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = ld_shared_v4(rbase + ((c ^ (row & 7)) << 4));
          x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) { uint32_t hi; split_tf32(x[i], hi, lo[i]); }
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kTsACol0 + s * 64;
        tmem_st_32x32b_x32(ta, x);        // kind::tf32 ignores the low 13 mantissa bits: raw == hi
        tmem_st_32x32b_x32(ta + 32, lo);
      }
      // ---- B: elementwise lo tile next to the raw (= hi) tile
      const uint32_t raw = hi_b(s);
#pragma unroll 8
      for (int i = 0; i < kTileBytes / 16 / (kSplitWarps * 32); ++i) {
        const uint32_t off = (static_cast<uint32_t>(i) * (kSplitWarps * 32) + t) * 16u;
        const uint4 v = ld_shared_v4(raw + off);
        uint4 hi, lo;
        split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
        split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        st_shared_v4(raw + kTileBytes + off, lo);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(split_full(s));
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, kTsTmemCols);
}

// ------------------------------------------------------------------ "TS2" variant: CTA pair, 256 x 256 tiles
// What bounds the TS kernel above is the granularity of its MMA instructions.  Measured on the config-2 Gram: issuing
// every 128 x 128 x 8 MMA as two 128 x 64 x 8 halves (same arithmetic, bit-identical results) takes 6.16 ms instead of
// 4.33 -- interleaved over two accumulator column ranges or not -- i.e. ~47 clocks of fixed cost per instruction next
// to 64 clocks of math: 42 % of the kernel.  (A first CTA-pair variant with 256 x 128 tiles, half the shared-memory
// traffic, a 6-stage ring and two splitter groups measured 4.04 - 4.26 ms: none of those was the limit.)  This kernel
// doubles N: a cluster of two CTAs computes a 256 x 256 tile with tcgen05.mma.cta_group::2, M = 256, N = 256 --
// each CTA owns 128 rows (its A tile, split into its own tensor memory) and 128 of the 256 B rows (raw + lo tile in
// its shared memory), one thread of the leader issues the MMAs for both.
//   shared memory per stage: rawA 16 KiB | rawB half (= hi) 16 KiB | lo B half 16 KiB = 48 KiB, 4 stages;
//   tensor memory: [0, 256) ONE accumulator (no ping-pong: the MMAs of the next chunk wait until the epilogue has
//   read the chunk out, ~5 % of a chunk), [256 + 64 s, +64) hi / lo A of stage s;
//   epilogue: 8 warps x 128 fp32 running sums per thread -- the register file is re-split with setmaxnreg
//   (epilogue 184, splitter 104, control warps 40 registers per thread).
// Barriers: raw_full / empty / acc_full are per CTA (commits are multicast to both); split_full and acc_empty live in
// the leader and collect the arrivals of both CTAs' splitter / epilogue warps.
constexpr int kT2BN = 256;
constexpr int kT2Stages = 4;
constexpr int kT2StageBytes = 3 * kTileBytes;
constexpr int kT2SmemBytes = kT2Stages * kT2StageBytes + 256 + 1024;
constexpr int kT2Threads = kThreads;

template <int kRegs>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// grid = (2 * row-tile pairs, 256-column tiles, batch), cluster = (2, 1, 1)
__global__ void __launch_bounds__(kT2Threads, 1)
gemm_tn_3xtf32_ts2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                          const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;  // 128-row tile of this CTA, 256-column tile of the pair
  const uint32_t rank = cluster_ctarank();             // 0 = leader (even row tile), 1 = the row tile below it
  // the pair leaves together: only when even its lower row tile (tile_m | 1) lies above the diagonal
  if (a.lower_only && tile_n * kT2BN > (tile_m | 1) * kBM + (kBM - 1)) return;
  const bool store = !(a.lower_only && tile_n * kT2BN > tile_m * kBM + (kBM - 1));  // the leader's tile may be above it

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kT2Stages * kT2StageBytes;
  auto raw_full = [&](int s) { return bar_base + 8u * s; };
  auto split_full = [&](int s) { return bar_base + 8u * (kT2Stages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kT2Stages + s); };
  const uint32_t acc_full = bar_base + 8u * (3 * kT2Stages);
  const uint32_t acc_empty = bar_base + 8u * (3 * kT2Stages + 1);
  const uint32_t tmem_slot = bar_base + 8u * (3 * kT2Stages + 2);
  auto raw_a = [&](int s) { return smem_base + s * kT2StageBytes; };
  auto hi_b = [&](int s) { return smem_base + s * kT2StageBytes + kTileBytes; };
  auto lo_b = [&](int s) { return smem_base + s * kT2StageBytes + 2 * kTileBytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contraction range: tri_b refers to 128-column blocks of B; a 256-column tile spans two of them
  const int nkb_all = (a.K + kBK - 1) / kBK;
  int kb0 = 0, kb1 = nkb_all;
  if (a.tri_b == 1) { const int e = (2 * tile_n + 2) * (kBN / kBK); kb1 = e < nkb_all ? e : nkb_all; }
  else if (a.tri_b == 2) { const int b0 = 2 * tile_n * (kBN / kBK); kb0 = b0 < nkb_all ? b0 : nkb_all; }
  const int nkb = kb1 - kb0;
  const int nchunks = (nkb + kChunkKB - 1) / kChunkKB;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tmap_a); prefetch_tensormap(&tmap_b); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kT2Stages; ++s) {
      mbar_init(raw_full(s), 1);
      mbar_init(split_full(s), 2 * kSplitWarps);  // used in the leader only
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 2 * kEpiWarps);          // used in the leader only
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<2>(tmem_slot, kTsTmemCols);
  tcgen05_fence_before();
  cluster_arrive_wait();  // both CTAs' barriers are initialised before anybody arrives remotely
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < kFirstEpiWarp) {
    setmaxnreg_dec<40>();
    if (warp == 0 && lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kT2Stages;
        const uint32_t ph = (kb / kT2Stages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u, 0x2900 + s);
        mbar_arrive_expect_tx(raw_full(s), 2 * kTileBytes);
        tma_load_3d(raw_a(s), &tmap_a, raw_full(s), (kb0 + kb) * kBK, tile_m * kBM, blockIdx.z);
        tma_load_3d(hi_b(s), &tmap_b, raw_full(s), (kb0 + kb) * kBK, tile_n * kT2BN + static_cast<int>(rank) * kBN,
                    blockIdx.z);
      }
    } else if (warp == 1 && rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc(UMMA_FMT_TF32, 2 * kBM, kT2BN);
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(acc_empty, (c & 1u) ^ 1u, 0x2d00);  // both CTAs' epilogues have read chunk c - 1 out
        tcgen05_fence_after();
        const int kb_end = (c + 1) * kChunkKB < nkb ? (c + 1) * kChunkKB : nkb;
        for (int kb = c * kChunkKB; kb < kb_end; ++kb) {
          const int s = kb % kT2Stages;
          const uint32_t ph = (kb / kT2Stages) & 1u;
          mbar_wait(split_full(s), ph, 0x2a00 + s);
          tcgen05_fence_after();
          const uint64_t dhb = umma_desc_kmajor_sw128(hi_b(s)), dlb = umma_desc_kmajor_sw128(lo_b(s));
          const uint32_t a_hi = tmem_base + kTsACol0 + s * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint32_t first = (kb == c * kChunkKB && k == 0) ? 0u : 1u;
            umma_tf32_ts_cg2(tmem_base, a_lo + k * kUmmaK, dhb + 2u * k, idesc, first);  // small terms first
            umma_tf32_ts_cg2(tmem_base, a_hi + k * kUmmaK, dlb + 2u * k, idesc, 1u);
            umma_tf32_ts_cg2(tmem_base, a_hi + k * kUmmaK, dhb + 2u * k, idesc, 1u);
          }
          umma_commit_cg2_mcast(empty_bar(s), 0x3);
        }
        umma_commit_cg2_mcast(acc_full, 0x3);
      }
    }
  } else if (warp < kFirstSplitWarp) {
    setmaxnreg_inc<184>();
    const int e = warp - kFirstEpiWarp;
    const int q = warp & 3;    // TMEM lane quarter this warp may read
    const int half = e >> 2;   // which 128 of the 256 accumulator columns
    float acc[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) acc[i] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(acc_full, c & 1u, 0x2b00);
      tcgen05_fence_after();
      const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t0 + j * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[j * 32 + i] = __fadd_rn(acc[j * 32 + i], __uint_as_float(v[i]));
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(acc_empty); else mbar_arrive_cluster(mapa(acc_empty, 0));
      }
    }
    const int64_t row = static_cast<int64_t>(tile_m) * kBM + q * 32 + lane;
    const int64_t col0 = static_cast<int64_t>(tile_n) * kT2BN + half * 128;
    // lower_only keeps the single-CTA kernel's contract at 128-block granularity: blocks above the diagonal untouched
    const bool store_half = store && !(a.lower_only && (2 * tile_n + half) > tile_m);
    if (store_half && row < a.M) {
      float* crow = a.C + static_cast<int64_t>(blockIdx.z) * a.stride_c + row * a.ldc;
#pragma unroll
      for (int i = 0; i < 128; ++i) {
        const int64_t col = col0 + i;
        if (col < a.N) {
          float r = a.alpha * acc[i];
          if (a.beta != 0.f) r += a.beta * crow[col];
          if (col == row) r += a.diag_add;
          crow[col] = r;
        }
      }
    }
  } else {
    setmaxnreg_dec<104>();
    const int q = warp & 3;
    const int t = threadIdx.x - kFirstSplitWarp * 32;
    const int row = q * 32 + lane;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kT2Stages;
      const uint32_t ph = (kb / kT2Stages) & 1u;
      mbar_wait(raw_full(s), ph, 0x2c00 + s);
      {
        uint32_t x[32], lo[32];
        const uint32_t rbase = raw_a(s) + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = ld_shared_v4(rbase + ((c ^ (row & 7)) << 4));
          x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) { uint32_t hi; split_tf32(x[i], hi, lo[i]); }
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kTsACol0 + s * 64;
        tmem_st_32x32b_x32(ta, x);
        tmem_st_32x32b_x32(ta + 32, lo);
      }
      const uint32_t raw = hi_b(s);
#pragma unroll 8
      for (int i = 0; i < kTileBytes / 16 / (kSplitWarps * 32); ++i) {
        const uint32_t off = (static_cast<uint32_t>(i) * (kSplitWarps * 32) + t) * 16u;
        const uint4 v = ld_shared_v4(raw + off);
        uint4 hi, lo;
        split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
        split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        st_shared_v4(raw + kTileBytes + off, lo);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(split_full(s)); else mbar_arrive_cluster(mapa(split_full(s), 0));
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  cluster_arrive_wait();  // the peer's tensor memory / barriers are not touched after this point
  if (warp == 2) tmem_dealloc<2>(tmem_base, kTsTmemCols);
}

// ------------------------------------------------------------------ helpers around the GEMM

// out[c, r] = in[r, c]; in: [R, C] pitch ld_in; out: [C, R_pad] pitch ld_out (columns R..ld_out untouched).
// HBM-bound (8 B per element).  Tile = 128 rows x 64 columns: 256-byte read segments, 512-byte contiguous write
// segments per output row -- with 32 x 32 tiles every 128-byte segment opened its own DRAM page on both sides and the
// kernel ran at 300 GB/s (5.5 ms for the 50 000 x 4096 features of config 2); now 1.2 ms (1.35 TB/s).  Measured and
// rejected in round 2: 128 x 128 tiles (1.38 ms) and four consecutive row tiles per CTA (worse).
constexpr int kTrRows = 128, kTrCols = 64, kTrSub = 1;
constexpr int kTrSmemBytes = kTrCols * (kTrRows + 1) * 4;
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, int64_t R, int64_t Ccols, int64_t ld_in, float* __restrict__ out,
                 int64_t ld_out) {
  extern __shared__ float tr_tile[];  // [kTrCols][kTrRows + 1]
  auto tile = [&](int c, int r) -> float& { return tr_tile[c * (kTrRows + 1) + r]; };
  const int x = threadIdx.x, y = threadIdx.y;  // (32, 8)
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * kTrCols;
  const bool vec_ok = (ld_in % 2 == 0) && ((reinterpret_cast<uintptr_t>(in) & 7u) == 0);
  for (int sub = 0; sub < kTrSub; ++sub) {
  const int64_t r0 = (static_cast<int64_t>(blockIdx.y) * kTrSub + sub) * kTrRows;
  if (r0 >= R) break;
  if (sub) __syncthreads();  // the previous tile has been written out
#pragma unroll 4
  for (int i = y; i < kTrRows; i += 8) {
    const int64_t r = r0 + i;
#pragma unroll
    for (int h = 0; h < kTrCols / 64; ++h) {  // a warp reads 2 x 256 contiguous bytes of one row
      const int cl = h * 64 + 2 * x;
      const int64_t c = c0 + cl;
      float2 v = make_float2(0.f, 0.f);
      if (r < R) {
        if (vec_ok && c + 1 < Ccols) v = *reinterpret_cast<const float2*>(in + r * ld_in + c);
        else {
          if (c < Ccols) v.x = in[r * ld_in + c];
          if (c + 1 < Ccols) v.y = in[r * ld_in + c + 1];
        }
      }
      tile(cl, i) = v.x;
      tile(cl + 1, i) = v.y;
    }
  }
  __syncthreads();
#pragma unroll 2
  for (int cc = y; cc < kTrCols; cc += 8) {
    const int64_t c = c0 + cc;
    if (c >= Ccols) continue;
#pragma unroll
    for (int j = 0; j < kTrRows / 32; ++j) {
      const int64_t r = r0 + x + 32 * j;
      if (r < R) out[c * ld_out + r] = tile(cc, x + 32 * j);
    }
  }
  }
}

// Cholesky of one nb x nb diagonal block (nb <= 128), in place (lower; strict upper zeroed), plus the
// inverse of the factor and its transpose (dense 128 x 128, identity-padded) for the GEMM-based panel solves.
// One CTA of 512 threads, block resident in shared memory.  The block is processed as 4 x 4 sub-blocks of 32:
//   * the 32 x 32 diagonal sub-block is factored by ONE warp in registers (lane i owns row i, the column
//     broadcasts are warp shuffles: no block barrier inside the 32 columns) and inverted by the same warp
//     (lane j owns column j of the inverse, factor entries are smem broadcasts);
//   * the rows below it are solved against that factor by forward substitution (one thread per row, broadcast reads
//     of the factor) WHILE warp 0 inverts the sub-block, and the trailing part is updated by all 16 warps with
//     4 x 4 register tiles (round 1: multiply by the inverse after waiting for it, one output element per thread);
//   * the 128 x 128 inverse is assembled from the four 32 x 32 inverses by two levels of
//     X21 = -X22 (L21 X11)  (32 -> 64 -> 128), register-tiled as well.
// 3 block barriers per 32 columns.  The 32 launches of this kernel are strictly serial with the panel / trailing GEMMs
// of the blocked Cholesky: at 80-100 us each they were 3.6 of the 4.5 ms factorisation of the config-2 Gram matrix.
constexpr int kPotrfNb = 128;
constexpr int kPotrfLd = kPotrfNb + 1;
constexpr int kPotrfSb = 32;
constexpr int kPotrfThreads = 512;
constexpr int kPotrfTLd = 65;
constexpr int kPotrfSmem = (2 * kPotrfNb * kPotrfLd + 64 * kPotrfTLd) * 4;

// C[r, c] (+)= sign * sum_t A[r, t] * B[t, c] (kBT = false) or A[r, t] * B[c, t] (kBT = true) over an R x Cn block with
// inner length T; all operands are smem sub-matrices with leading dimensions lda / ldb / ldc.  Whole CTA.
// (r, c) pairs are dealt c-fastest so that a warp reads one A row (broadcast) and 32 B columns / rows.
template <bool kBT, bool kAccum>
__device__ __forceinline__ void smem_block_mm(const float* A, int lda, const float* B, int ldb, float* Cm, int ldc, int R,
                                              int Cn, int T, float sign) {
  for (int idx = threadIdx.x; idx < R * Cn; idx += kPotrfThreads) {
    const int r = idx / Cn, c = idx % Cn;
    float acc = 0.f;
#pragma unroll 8
    for (int t = 0; t < T; ++t) acc = fmaf(A[r * lda + t], kBT ? B[c * ldb + t] : B[t * ldb + c], acc);
    Cm[r * ldc + c] = kAccum ? fmaf(sign, acc, Cm[r * ldc + c]) : sign * acc;
  }
}

// The same product with a 4 x 4 register tile per thread (8 shared-memory loads per 16 FMAs instead of 2 per 1):
// R and Cn are multiples of 4.  kLowerOnly skips output tiles strictly above the diagonal (symmetric updates).
template <bool kBT, bool kAccum, bool kLowerOnly>
__device__ __forceinline__ void smem_tile_mm(const float* A, int lda, const float* B, int ldb, float* Cm, int ldc, int R,
                                             int Cn, int T, float sign) {
  // A thread owns rows 4 ti .. 4 ti + 3 and the four columns tj, tj + Cn/4, tj + 2 Cn/4, tj + 3 Cn/4: consecutive
  // threads then read consecutive B rows (kBT: pitch 129 -> consecutive banks) or consecutive B / C words.  With four
  // ADJACENT columns per thread (round 1) the B and C accesses of a warp were 4 words apart: 4-way bank conflicts on
  // 5 of every 8 shared-memory loads, and shared memory is what bounds these products (8 loads per 16 FMAs).
  const int tr = R >> 2, tcw = Cn >> 2;
  for (int idx = threadIdx.x; idx < tr * tcw; idx += kPotrfThreads) {
    const int ti = idx / tcw, tj = idx % tcw;
    int nj = 4;  // columns of this thread that reach the lower triangle (column <= last row of the tile)
    if (kLowerOnly) {
      if (tj > ti * 4 + 3) continue;
      nj = (ti * 4 + 3 - tj) / tcw + 1;
      if (nj > 4) nj = 4;
    }
    const float* a = A + (ti * 4) * lda;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = a[i * lda + t];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        bv[j] = (j < nj) ? (kBT ? B[(tj + j * tcw) * ldb + t] : B[t * ldb + tj + j * tcw]) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= nj) continue;
        float* c = Cm + (ti * 4 + i) * ldc + tj + j * tcw;
        *c = kAccum ? fmaf(sign, acc[i][j], *c) : sign * acc[i][j];
      }
  }
}

#ifdef GADM_POTRF_PROFILE
#define POTRF_T(i) do { __syncthreads(); if (threadIdx.x == 0) prof_t[i] = clock64(); } while (0)
#else
#define POTRF_T(i) do { } while (0)
#endif
// C (128 x 128, lower 4 x 4 tiles incl. the diagonal ones) -= P P^T for a 128 x 128 P, all in shared memory: the
// 528 needed tiles are enumerated along the triangle and dealt round-robin, so every thread owns one tile (16 threads
// two) -- the generic routine above deals the full 32 x 32 tile grid and skips the upper half, which leaves a quarter
// of the threads with two tiles and a quarter with none (this product sits on the critical path of every step).
__device__ __forceinline__ void smem_syrk_lower_128(const float* P, int ldp, float* Cm, int ldc) {
  constexpr int kTiles = 32 * 33 / 2;
  for (int idx = threadIdx.x; idx < kTiles; idx += kPotrfThreads) {
    int ti = static_cast<int>((sqrtf(8.f * static_cast<float>(idx) + 1.f) - 1.f) * 0.5f);
    while (ti * (ti + 1) / 2 > idx) --ti;            // guard the float square root at tile-row boundaries
    while ((ti + 1) * (ti + 2) / 2 <= idx) ++ti;
    const int tj = idx - ti * (ti + 1) / 2;          // tj <= ti
    const float* a = P + (ti * 4) * ldp;
    const float* b = P + (tj * 4) * ldp;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int t = 0; t < kPotrfNb; ++t) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = a[i * ldp + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = b[j * ldp + t];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float* c = Cm + (ti * 4 + i) * ldc + tj * 4 + j;
        *c = *c - acc[i][j];
      }
  }
}

__global__ void __launch_bounds__(kPotrfThreads, 1)
potrf_diag_kernel(float* __restrict__ A, int64_t ld, int nb, float* __restrict__ linv, float* __restrict__ linv_t,
                  int* __restrict__ info, int block_index, const float* __restrict__ prev) {
  extern __shared__ float potrf_smem[];
  float* L = potrf_smem;                        // [128][kPotrfLd]: the block, then its factor
  float* X = potrf_smem + kPotrfNb * kPotrfLd;  // inverse of the factor
  float* Tm = X + kPotrfNb * kPotrfLd;          // [64][kPotrfTLd] scratch
  __shared__ float dinv_s[kPotrfSb];            // 1 / diagonal of the current 32 x 32 factor
  __shared__ __align__(16) float col_s[2 * kPotrfSb];  // the factor column being eliminated (double-buffered)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef GADM_POTRF_PROFILE
  __shared__ long long prof_t[24];
  long long acc_f = 0, acc_p = 0, acc_u = 0;
#endif
  POTRF_T(0);
  // load; rows / columns >= nb are padded with the identity so that every sub-block step is well defined
  for (int idx = tid; idx < kPotrfNb * kPotrfNb; idx += kPotrfThreads) {
    const int r = idx / kPotrfNb, c = idx % kPotrfNb;
    L[r * kPotrfLd + c] = (r < nb && c < nb) ? A[static_cast<int64_t>(r) * ld + c] : ((r == c) ? 1.f : 0.f);
    // prev: this block row of the previous block column of the factor, L[b, b-1] (nb x 128, pitch ld).  The blocked
    // Cholesky leaves the last rank-128 update of this diagonal block to this kernel (A_bb -= L[b,b-1] L[b,b-1]^T):
    // as a separate GEMM launch it sat on the critical path between panel(b-1) and this kernel (29 us per step).
    X[r * kPotrfLd + c] = (prev != nullptr && r < nb) ? prev[static_cast<int64_t>(r) * ld + c] : 0.f;
  }
  __syncthreads();
  if (prev != nullptr) {
    smem_syrk_lower_128(X, kPotrfLd, L, kPotrfLd);
    __syncthreads();
    for (int idx = tid; idx < kPotrfNb * kPotrfNb; idx += kPotrfThreads) X[(idx / kPotrfNb) * kPotrfLd + idx % kPotrfNb] = 0.f;
    __syncthreads();
  }

  POTRF_T(1);
  for (int jb = 0; jb < kPotrfNb / kPotrfSb; ++jb) {
    const int j0 = jb * kPotrfSb;
    POTRF_T(4);
    if (warp == 0) {
      // ---- factor the 32 x 32 diagonal sub-block in registers: lane i owns row i.  Column c of the factor is
      // published through a double-buffered 32-float shared-memory line and read back as broadcast LDS.128 (8 loads)
      // instead of one shuffle per remaining row (31 per column): the shuffles were the issue-bound part of this
      // one-warp critical path (6.6 k cycles per sub-block, a third of the kernel).
      float r[kPotrfSb];
      float dinv = 1.f;
#pragma unroll
      for (int c = 0; c < kPotrfSb; ++c) r[c] = L[(j0 + lane) * kPotrfLd + j0 + c];
#pragma unroll
      for (int c = 0; c < kPotrfSb; ++c) {
        float d = __shfl_sync(0xffffffffu, r[c], c);
        if (!(d > 0.f)) {
          if (lane == 0 && info) atomicMax(info, block_index * kPotrfNb + j0 + c + 1);
          d = 1.f;
        }
        // 1 / sqrt(d): MUFU.RSQ + one Newton step (full fp32 accuracy) instead of an IEEE sqrt and a division per
        // column -- both sat on the one-warp critical path (the kernel ran 94 us per diagonal block)
        float y = rsqrtf(d);
        y = y * fmaf(-0.5f * d * y, y, 1.5f);
        const float lc = (lane == c) ? d * y : r[c] * y;  // column c of the factor, held by lane = row
        if (lane == c) dinv = y;                          // 1 / L[c][c], reused by the panel solve and the inverse
        r[c] = lc;
        if (c + 1 < kPotrfSb) {
          float* line = col_s + (c & 1) * kPotrfSb;
          line[lane] = lc;
          __syncwarp();
#pragma unroll
          for (int q = (c + 1) / 4; q < kPotrfSb / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(line + 4 * q);  // rows 4q .. 4q + 3 of column c
            if (4 * q + 0 > c) r[4 * q + 0] = fmaf(-lc, v.x, r[4 * q + 0]);  // only rows >= t are meaningful;
            if (4 * q + 1 > c) r[4 * q + 1] = fmaf(-lc, v.y, r[4 * q + 1]);  // the others are never read again
            if (4 * q + 2 > c) r[4 * q + 2] = fmaf(-lc, v.z, r[4 * q + 2]);
            if (4 * q + 3 > c) r[4 * q + 3] = fmaf(-lc, v.w, r[4 * q + 3]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kPotrfSb; ++c) L[(j0 + lane) * kPotrfLd + j0 + c] = (c <= lane) ? r[c] : 0.f;
      dinv_s[lane] = dinv;
    }
    __syncthreads();
    POTRF_T(5);
    const int rem = kPotrfNb - (j0 + kPotrfSb);  // rows below the diagonal sub-block
    if (warp == 0) {
      // ---- invert the sub-block (needed for the 128 x 128 inverse only: off the critical path, it runs while warps
      // 1..3 solve the panel): lane j owns column j of the inverse (forward substitution down the rows)
      float x[kPotrfSb];
#pragma unroll
      for (int i = 0; i < kPotrfSb; ++i) {
        float s0 = (i == lane) ? 1.f : 0.f;
#pragma unroll
        for (int t = 0; t < i; ++t) s0 = fmaf(-L[(j0 + i) * kPotrfLd + j0 + t], x[t], s0);  // x[t] = 0 for t < lane
        x[i] = (i >= lane) ? s0 * dinv_s[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < kPotrfSb; ++i) X[(j0 + i) * kPotrfLd + j0 + lane] = x[i];
    } else if (tid - 32 < rem) {
      // ---- panel by forward substitution, one thread per row: p D^T = a  (D = the 32 x 32 factor, broadcast reads)
      // instead of waiting for D^-1 and multiplying by it; four partial sums shorten the dependent FMA chain
      float* arow = L + (j0 + kPotrfSb + (tid - 32)) * kPotrfLd + j0;
      float pr[kPotrfSb];
#pragma unroll
      for (int c = 0; c < kPotrfSb; ++c) {
        const float* drow = L + (j0 + c) * kPotrfLd + j0;
        float s0 = arow[c], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int t = 0; t + 3 < c; t += 4) {
          s0 = fmaf(-pr[t], drow[t], s0);
          s1 = fmaf(-pr[t + 1], drow[t + 1], s1);
          s2 = fmaf(-pr[t + 2], drow[t + 2], s2);
          s3 = fmaf(-pr[t + 3], drow[t + 3], s3);
        }
#pragma unroll
        for (int t = c & ~3; t < c; ++t) s0 = fmaf(-pr[t], drow[t], s0);
        pr[c] = ((s0 + s1) + (s2 + s3)) * dinv_s[c];
      }
#pragma unroll
      for (int c = 0; c < kPotrfSb; ++c) arow[c] = pr[c];
    }
    __syncthreads();
    POTRF_T(6);
    if (rem > 0) {
      // ---- trailing update (lower 4 x 4 tiles incl. the diagonal ones): A22 -= P P^T, P = the panel just solved
      smem_tile_mm<true, true, true>(L + (j0 + kPotrfSb) * kPotrfLd + j0, kPotrfLd, L + (j0 + kPotrfSb) * kPotrfLd + j0,
                                     kPotrfLd, L + (j0 + kPotrfSb) * kPotrfLd + j0 + kPotrfSb, kPotrfLd, rem, rem, kPotrfSb,
                                     -1.f);
      __syncthreads();
      POTRF_T(7);
#ifdef GADM_POTRF_PROFILE
      if (tid == 0) { acc_p += prof_t[6] - prof_t[5]; acc_u += prof_t[7] - prof_t[6]; }
#endif
    }
#ifdef GADM_POTRF_PROFILE
    if (tid == 0) acc_f += prof_t[5] - prof_t[4];
#endif
  }
  POTRF_T(2);
  // ---- X = L^-1 from the four 32 x 32 diagonal inverses: X21 = -X22 (L21 X11), level 32 -> 64, then 64 -> 128
  for (int half = kPotrfSb; half < kPotrfNb; half *= 2) {
    for (int g0 = 0; g0 < kPotrfNb; g0 += 2 * half) {
      // groups are independent but share the scratch: serialised (2 groups at the first level, 1 at the second)
      smem_tile_mm<false, false, false>(L + (g0 + half) * kPotrfLd + g0, kPotrfLd, X + g0 * kPotrfLd + g0, kPotrfLd, Tm,
                                        kPotrfTLd, half, half, half, 1.f);                              // T = L21 X11
      __syncthreads();
      smem_tile_mm<false, false, false>(X + (g0 + half) * kPotrfLd + g0 + half, kPotrfLd, Tm, kPotrfTLd,
                                        X + (g0 + half) * kPotrfLd + g0, kPotrfLd, half, half, half, -1.f);  // X21 = -X22 T
      __syncthreads();
    }
  }
  POTRF_T(3);
  for (int idx = tid; idx < kPotrfNb * kPotrfNb; idx += kPotrfThreads) {
    const int r = idx / kPotrfNb, c = idx % kPotrfNb;
    if (r < nb && c < nb) A[static_cast<int64_t>(r) * ld + c] = (c <= r) ? L[r * kPotrfLd + c] : 0.f;
    // identity padding (rows >= nb) keeps the block invertible.  Both outputs are written with consecutive threads on
    // consecutive addresses; the transpose is taken on the shared-memory side (pitch 129: conflict-free)
    linv[r * kPotrfNb + c] = X[r * kPotrfLd + c];
    linv_t[r * kPotrfNb + c] = X[c * kPotrfLd + r];
  }
#ifdef GADM_POTRF_PROFILE
  POTRF_T(8);
  if (tid == 0 && block_index == 0)
    printf("potrf cycles: load %lld | factor+invert (warp 0) %lld | panel %lld | trailing %lld | loop total %lld | assemble X %lld | store %lld | all %lld\n",
           prof_t[1] - prof_t[0], acc_f, acc_p, acc_u, prof_t[2] - prof_t[1], prof_t[3] - prof_t[2], prof_t[8] - prof_t[3],
           prof_t[8] - prof_t[0]);
#endif
}

// X (lower) and Xt (upper) <- the 128 x 128 diagonal-block inverses that potrf_diag_kernel left in the workspace;
// everything else of the two k x k matrices is zeroed by the caller.  grid.x = diagonal block.
__global__ void tri_inverse_init_kernel(const float* __restrict__ linv, const float* __restrict__ linv_t, int64_t k,
                                        float* __restrict__ X, int64_t ldx, float* __restrict__ Xt, int64_t ldxt) {
  const int64_t b = blockIdx.x, i0 = b * kPotrfNb;
  for (int idx = threadIdx.x; idx < kPotrfNb * kPotrfNb; idx += blockDim.x) {
    const int r = idx / kPotrfNb, c = idx % kPotrfNb;
    if (i0 + r < k && i0 + c < k) {
      X[(i0 + r) * ldx + i0 + c] = linv[b * kPotrfNb * kPotrfNb + idx];
      Xt[(i0 + r) * ldxt + i0 + c] = linv_t[b * kPotrfNb * kPotrfNb + idx];
    }
  }
}

// out[r] = ||x[r, :]||_2 (or its reciprocal), fp32 data, fp64 accumulation; one warp per row
__global__ void row_norms_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld, int reciprocal,
                                 float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  double s = 0.0;
  for (int64_t c = lane; c < cols; c += 32) { const double v = x[r * ld + c]; s += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = static_cast<float>(reciprocal ? 1.0 / sqrt(s) : sqrt(s));
}

// out[n] = mean_t S[t, n] * row_scale[t] * col_scale[n]   (traks.py:146,157,162-168), fp64 accumulation.
// Block (32 columns, 32 row lanes): thread (x, y) sums rows y, y + 32, ... of column x, the 32 partial sums are added
// in lane order -- a fixed order, deterministic.  (One thread per column over all T rows left the mean over the 1000
// generated features of config 2 with 4096 threads and a serial chain of 1000 loads each: 158 us for 16 MB.)
constexpr int kColMeanLanes = 32;
__global__ void __launch_bounds__(32 * kColMeanLanes)
col_mean_scaled_kernel(const float* __restrict__ S, int64_t T, int64_t N, int64_t ld, const float* __restrict__ row_scale,
                       const float* __restrict__ col_scale, float* __restrict__ out) {
  __shared__ double part[kColMeanLanes][33];
  const int x = threadIdx.x, y = threadIdx.y;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * 32 + x;
  double s = 0.0;
  if (n < N) {
    for (int64_t t = y; t < T; t += kColMeanLanes) {
      const float v = S[t * ld + n];
      s += row_scale ? static_cast<double>(v) * row_scale[t] : static_cast<double>(v);
    }
  }
  part[y][x] = s;
  __syncthreads();
  if (y == 0 && n < N) {
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < kColMeanLanes; ++i) tot += part[i][x];
    tot /= static_cast<double>(T);
    if (col_scale) tot *= col_scale[n];
    out[n] = static_cast<float>(tot);
  }
}

// S[t, n] *= row_scale[t] * col_scale[n]  (compute_gradient_score.py:114-126 "scores / magnitude")
__global__ void scale_rows_cols_kernel(float* __restrict__ S, int64_t T, int64_t N, int64_t ld,
                                       const float* __restrict__ row_scale, const float* __restrict__ col_scale) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t t = blockIdx.y;
  if (n >= N || t >= T) return;
  float v = S[t * ld + n];
  if (row_scale) v *= row_scale[t];
  if (col_scale) v *= col_scale[n];
  S[t * ld + n] = v;
}

// out[n] = col_scale[n] * sum_j x[n, j] * v[j]: the mean-over-generated-images forms of traks.py:157,162-168 are
// linear in the generated features, so mean_t(gen_t K^-1 phi_n) = (mean_t gen_t) K^-1 phi_n needs one row of
// K^-1-solved features and this matrix-vector product instead of the [T, N] score GEMM.  HBM-bound (N * k * 4 bytes);
// one warp per row, float4 loads, fp64 accumulation in a fixed order.
__global__ void matvec_rows_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                                   const float* __restrict__ v, const float* __restrict__ col_scale,
                                   float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + r * ld;
  double s = 0.0;
  if ((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0) {
    const int64_t c4 = cols >> 2;
    for (int64_t c = lane; c < c4; c += 32) {
      const float4 a = reinterpret_cast<const float4*>(xr)[c];
      const float4 b = reinterpret_cast<const float4*>(v)[c];
      s += static_cast<double>(a.x) * b.x + static_cast<double>(a.y) * b.y + static_cast<double>(a.z) * b.z +
           static_cast<double>(a.w) * b.w;
    }
    for (int64_t c = (c4 << 2) + lane; c < cols; c += 32) s += static_cast<double>(xr[c]) * v[c];
  } else {
    for (int64_t c = lane; c < cols; c += 32) s += static_cast<double>(xr[c]) * v[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = static_cast<float>(col_scale ? s * col_scale[r] : s);
}

// out[0] = min_i L[i, i], out[1] = max_i L[i, i] of a Cholesky factor: (max / min)^2 is a lower bound of cond(K) and
// tells the host whether an fp32 factorisation can be trusted at all (one CTA; k reads).
__global__ void diag_minmax_kernel(const float* __restrict__ L, int64_t ld, int64_t k, float* __restrict__ out) {
  __shared__ float smin[32], smax[32];
  float lo = __int_as_float(0x7f800000), hi = 0.f;
  for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
    const float d = L[i * ld + i];
    lo = fminf(lo, d);
    hi = fmaxf(hi, d);
    if (!(d == d)) lo = d;  // NaN poisons the minimum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = (l2 != l2 || lo != lo) ? __int_as_float(0x7fc00000) : fminf(lo, l2);
    hi = fmaxf(hi, h2);
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned w = 1; w < (blockDim.x >> 5); ++w) {
      lo = (smin[w] != smin[w] || lo != lo) ? __int_as_float(0x7fc00000) : fminf(lo, smin[w]);
      hi = fmaxf(hi, smax[w]);
    }
    out[0] = lo;
    out[1] = hi;
  }
}

}  // namespace gemm
}  // namespace gadm
