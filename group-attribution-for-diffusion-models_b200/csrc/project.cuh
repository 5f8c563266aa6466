// JL projection  Phi[M, k] = G[M, D] * P[D, k]  with P generated on the fly (never stored).
//
// Replaces fast_jl.project_{normal,rademacher}_{8,16,32} behind trak.projectors.CudaProjector.project
// (reference call sites: src/attributions/methods/d_trak_grad.py:776,
// text_to_image/grad_text_to_image_lora.py:765,813).
//
// Kernel shape (sm_100a):
//   * one CTA pair (cluster of 2, tcgen05 cta_group::2) per "unit" = (256-column tile of Phi, D-split);
//     UMMA 256x256x16 bf16 -> fp32, two accumulators (2 x 256 TMEM columns) so that a generated P
//     tile feeds up to 512 staged gradient rows;
//   * A operand (staged bf16 gradients, K-major) arrives by TMA into 128B-swizzled smem; the staging
//     buffer is tile-major [D_pad/64][m_cap][64] so that a 128-row x 64-column tile is 16 KiB of
//     contiguous HBM (with a plain [M, D] layout every 128-byte row of a tile opens its own DRAM page
//     and the A feed, not the tensor pipe, capped the kernel at ~1.0 PFLOP/s);
//   * B operand (P tile, K-major, 128B swizzle) is *written by generator warps* straight into the
//     UMMA smem layout from Philox4x32-10 (see philox.cuh), then published to the async proxy;
//   * warp roles (struct Roles): generators first (4 groups of 2 or 4 warps, group g owns pipeline
//     slot g), then 4 epilogue warps (TMEM -> registers -> split-K partial tile in HBM), the TMEM
//     allocator warp, and last the MMA issuer (leader CTA, one thread) and the TMA producer;
//   * split-K partials are reduced in a fixed order by project_reduce_kernel => results do not
//     depend on the schedule, the SM count or atomics.
// kCtaGroup == 1 is the single-CTA variant of the same code (UMMA 128x256x16, 3 stages).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "gadm_ptx.cuh"
#include "philox.cuh"

namespace gadm {
namespace proj {

constexpr int kBlockK = 64;   // bf16 elements per smem row = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kTileN = 256;   // Phi columns per unit
constexpr int kAccRows = 128; // gradient rows per CTA per accumulator
constexpr int kNumAcc = 2;
constexpr int kMaxGenGroups = 4;  // one generator group per pipeline slot (see Cfg::kGenGroups)
constexpr int kTmemCols = 512;
#ifndef GADM_GEN_BACKOFF_NS
#define GADM_GEN_BACKOFF_NS 128
#endif
#ifndef GADM_EPI_BACKOFF_NS
#define GADM_EPI_BACKOFF_NS 512
#endif
constexpr int kGenBackoffNs = GADM_GEN_BACKOFF_NS;  // poll interval of generator warps waiting for their slot (mbar_wait)
constexpr int kEpiBackoffNs = GADM_EPI_BACKOFF_NS;  // ... of epilogue warps waiting for the end of a segment

// Register budget of the projection kernels: by default whatever fits one CTA per SM (__launch_bounds__); a build with
// -DGADM_PROJ_MAXNREG=N caps it instead, which leaves more of the register file to kernels that run beside the
// persistent projection (staging of the next pass).
#ifdef GADM_PROJ_MAXNREG
#define GADM_PROJ_BOUNDS __maxnreg__(GADM_PROJ_MAXNREG)
#else
#define GADM_PROJ_BOUNDS __launch_bounds__(Roles<kWarpsPerGroup>::kThreads, 1)
#endif

// Warp roles.  Generator warps come FIRST and the single-thread TMA / MMA roles LAST: the SM sub-partition
// arbiter favours the highest warp id among eligible warps, and a late MMA / TMA issue stalls the whole
// pipeline while a late generator instruction does not.
template <int kWarpsPerGroup>
struct Roles {
  static constexpr int kGenWarps = kMaxGenGroups * kWarpsPerGroup;
  static constexpr int kGroupThreads = kWarpsPerGroup * 32;
  static constexpr int kFirstEpiWarp = kGenWarps;       // multiple of 4 -> epilogue warp w reads TMEM lanes 32*(w%4)
  static constexpr int kAllocWarp = kGenWarps + 4;
  static constexpr int kMmaWarp = kGenWarps + 6;
  static constexpr int kTmaWarp = kGenWarps + 7;
  static constexpr int kThreads = (kGenWarps + 8) * 32;
};

template <int kCtaGroup>
struct Cfg {
  static constexpr int kBRows = kTileN / kCtaGroup;                    // P rows (Phi columns) generated per CTA
  static constexpr int kATileBytes = kAccRows * kBlockK * 2;           // 16 KiB
  static constexpr int kABytes = kNumAcc * kATileBytes;                // 32 KiB
  static constexpr int kBBytes = kBRows * kBlockK * 2;                 // 16 / 32 KiB
  static constexpr int kStageBytes = kABytes + kBBytes;                // 48 / 64 KiB
  static constexpr int kStages = (kCtaGroup == 2) ? 4 : 3;
  // One generator group per pipeline slot: a group only ever waits for the *next* phase of its own
  // slot's barriers (mbarrier parity waits must never run more than one phase ahead).
  static constexpr int kGenGroups = kStages;
  static constexpr int kBarBytes = 256;
  static constexpr int kEpiBytes = 4 * 32 * 144;                       // per-warp transpose buffers of the epilogue
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kEpiBytes + 1024;  // + alignment slack
  static constexpr int kUnitRows = kNumAcc * kAccRows * kCtaGroup;     // rows of a full partial tile (512 / 256)
};

struct Args {
  float* partial;       // [n_units][unit_rows][256] fp32 split-K partial tiles
  uint32_t n_tiles;     // k / 256
  uint32_t n_splits;    // D-splits
  uint32_t n_units;     // n_tiles * n_splits
  uint32_t nkb_total;   // D_pad / 64
  uint32_t n_acc;       // accumulators in use (1 or 2)
  uint32_t unit_rows;   // n_acc * 128 * cta_group
  uint32_t key0, key1;  // Philox key = seed64
  uint32_t proj_type;   // ProjType
  uint32_t p_base_div64;  // canonical index of staged column 0, / 64
  uint32_t debug;         // perf ablation only (GADM_PROJ_DEBUG): bit0 skip generation, bit1 skip TMA loads
  uint32_t* sync_counter; // zeroed device word for the inter-cluster lockstep (nullptr: disabled)
  uint32_t sync_iters;    // lockstep only while the cluster-local k-block counter is below this (multiple of sync_every)
  uint32_t sync_every;    // k-blocks between lockstep points
  uint32_t seg_kb;        // k-blocks accumulated in TMEM before the accumulators are promoted into the partial tile
  uint32_t m_rows;        // staged rows in use (rows beyond it are never scaled nor reduced)
  uint32_t a_fmt;         // UMMA format of the staged gradients AND of the generated P: UMMA_FMT_BF16 or UMMA_FMT_F16
  const float* inv_scale; // F16G staging: [m_cap][scale_groups] inverse power-of-two scales (nullptr: unscaled)
  uint32_t scale_groups;  // scale groups per row = ceil(nkb_total / group_kb)
  uint32_t group_kb;      // k-blocks per scale group (multiple of seg_kb)
};

// Why segments: tcgen05 adds each MMA result into the fp32 TMEM accumulator with truncation.  Over a
// 15 000-k-block unit (60 000 accumulations) that shrinks every feature systematically -- measured -1.1e-3
// relative at the C2 shape (both projection types), -3 % at C3 -- far outside fp32-accumulate accuracy.  A unit
// is therefore cut into segments of seg_kb k-blocks: after each segment the epilogue warps add the accumulators
// into the unit's partial tile in HBM with round-to-nearest fp32 adds and the MMA restarts from zero.
// The add is a fire-and-forget vector reduction (REDG.E.ADD.F32x4.RN), a plain store for the first segment: no
// load, so the accumulators are released after ~16 TMEM loads instead of an HBM round trip per 64 B.  One thread
// owns an address for the whole launch and its st / red operations on it are ordered (same-thread, same
// location), so the sum order is fixed and the result deterministic.
// Segment boundaries lie on the GLOBAL k-block grid (multiples of seg_kb), not relative to the unit's first k-block:
// the F16G staging format carries one scale per (row, group of group_kb k-blocks) and a segment must not straddle
// two groups (group_kb % seg_kb == 0).  A unit's first and last segment may be short.
__device__ __forceinline__ uint32_t seg_end(uint32_t seg0, uint32_t kb1, uint32_t seg_kb) {
  const uint32_t e = (seg0 / seg_kb + 1) * seg_kb;
  return e < kb1 ? e : kb1;
}


// The partial tiles in flight (one per cluster, 33-37 MB in total) are re-touched once per segment; in between the
// gradient stream pushes them out of L2, so every segment costs a DRAM read + write of the tile.  An evict_last
// hint cuts the re-reads to a third in the quad kernel (ncu, full size: DRAM 178 -> 150 GB per launch; the
// write-backs remain, L2 cleans dirty lines regardless).  In the faster pair kernel the pinned tiles crowd the
// window in which the 16 clusters of a D-split share gradient tiles (71 -> 82 GB), so it issues the same instructions without a hint (pol == 0).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void red_add_v4(float* p, uint4 v, uint64_t pol) {
  if (pol != 0)
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)), "l"(pol)
                 : "memory");
  else
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}
__device__ __forceinline__ void st_global_v4_hint(float* p, uint4 v, uint64_t pol) {
  if (pol != 0)
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w), "l"(pol)
                 : "memory");
  else
    *reinterpret_cast<uint4*>(p) = v;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)
               : "memory");
  return v;
}

// Epilogue of one segment for one accumulator: the 32 TMEM lanes of this warp -> 32 partial-tile rows
// (store for the first segment, add afterwards).  tcgen05.ld hands lane r the 32 columns of row r; written
// out like that every instruction would touch 32 different 128-byte lines with 16 bytes each, and those LSU
// transactions, not the data, were the cost of a drain.  The 32x32 block is therefore transposed through a
// 4.5 KiB per-warp smem buffer (row pitch 144 B: conflict-free 16-byte accesses both ways) so that one
// instruction covers 4 rows x 128 contiguous bytes.
constexpr int kEpiPitch = 144;
constexpr int kEpiWarpBytes = 32 * kEpiPitch;
// `scale` (nullable): inverse staging scale of this segment's group for row i * 4 + rr of the warp's 32 rows is
// scale[(i * 4 + rr) * scale_ld]; rows >= rows_valid are not scaled (their partials are never reduced).
__device__ __forceinline__ void drain_accumulator(uint32_t tmem_addr, float* __restrict__ dst_warp, bool first,
                                                  uint32_t buf, int lane, uint64_t pol, const float* __restrict__ scale,
                                                  uint32_t scale_ld, int rows_valid) {
  const uint32_t my_row = buf + lane * kEpiPitch;
  const int rr = lane >> 3, cc = lane & 7;
  float sc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    sc[i] = (scale != nullptr && i * 4 + rr < rows_valid) ? __ldg(scale + static_cast<size_t>(i * 4 + rr) * scale_ld) : 1.f;
#pragma unroll 1
  for (int c = 0; c < kTileN; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem_addr + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) st_shared_v4(my_row + i * 16, make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i * 4 + rr;
      uint4 w = ld_shared_v4(buf + row * kEpiPitch + cc * 16);
      if (scale != nullptr) {  // exact: the scales are powers of two
        w.x = __float_as_uint(__uint_as_float(w.x) * sc[i]); w.y = __float_as_uint(__uint_as_float(w.y) * sc[i]);
        w.z = __float_as_uint(__uint_as_float(w.z) * sc[i]); w.w = __float_as_uint(__uint_as_float(w.w) * sc[i]);
      }
      float* g = dst_warp + static_cast<size_t>(row) * kTileN + c + cc * 4;
      if (first) st_global_v4_hint(g, w, pol);
      else red_add_v4(g, w, pol);
    }
    __syncwarp();
  }
}

// Inter-cluster lockstep.  The 16 clusters that work on the same D-split (one per 256-column tile) stream
// the SAME gradient tiles; they only hit in L2 if they stay within the L2 retention window of each other
// (~300 k-blocks).  Left alone they drift apart over a 15 000-k-block unit (two dies, arbitration jitter)
// and the staged gradients were re-read 4.6-8.6x from HBM.  Every kSyncEvery k-blocks the leader CTA's TMA
// thread therefore passes a monotonic global counter barrier; the 4-slot pipeline hides the wait.
// All clusters are co-resident (grid <= SM count, cooperative launch), so the spin cannot deadlock.
constexpr uint32_t kSyncEvery = 128;  // default; GADM_PROJ_SYNC_EVERY overrides for tuning
__device__ __forceinline__ void grid_lockstep(uint32_t* counter, uint32_t target) {
  atomicAdd(counter, 1u);
  const uint64_t t0 = globaltimer_ns();
  const uint64_t limit = *reinterpret_cast<volatile unsigned long long*>(&g_watchdog_ns);
  uint32_t v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v >= target) break;
    __nanosleep(200);
    if (limit != 0 && globaltimer_ns() - t0 > limit) watchdog_fire(0x600);
  } while (true);
}

// Fill one B stage (kBRows x 64 bf16, K-major, 128B swizzle) with Rademacher signs.
//   smem_b: stage base (1024-aligned); p_div32: canonical p of stage column 0, / 32; j0: first Phi column
template <int kBRows, int kGroupThreads>
__device__ __forceinline__ void gen_rademacher_stage(uint32_t smem_b, uint32_t p_div32, uint32_t j0, uint32_t k0,
                                                     uint32_t k1, int tig, int lane, uint32_t ones) {
  constexpr int kJGroups = kBRows / 4;
  constexpr int kCalls = kJGroups * 2;
  const int rot = (lane >> 1) & 3;
#pragma unroll
  for (int c = tig; c < kCalls; c += kGroupThreads) {
    const int jg = c % kJGroups;
    const int pg = c / kJGroups;
    uint4 w = rademacher_call(p_div32 + pg, (j0 >> 2) + jg, k0, k1);
    // rotate the four words by `rot` so that the 8 lanes of a quarter-warp store to 8 distinct rows mod 8
    if (rot & 1) { const uint32_t t = w.x; w.x = w.y; w.y = w.z; w.z = w.w; w.w = t; }
    if (rot & 2) { uint32_t t = w.x; w.x = w.z; w.z = t; t = w.y; w.y = w.w; w.w = t; }
    const uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = 4 * jg + ((i + rot) & 3);
      const uint32_t row_addr = smem_b + row * 128;
      const uint32_t sw = row & 7;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const uint32_t byte = (words[i] >> (8 * cc)) & 0xFFu;
        const uint32_t chunk = 4 * pg + cc;
        st_shared_v4(row_addr + ((chunk ^ sw) << 4), rademacher_expand8(byte, ones));
      }
    }
  }
}

// Fill one B stage with bf16 N(0,1) values.  p_div8: canonical p of stage column 0, / 8.
template <int kBRows, int kGroupThreads, bool kF16>
__device__ __forceinline__ void gen_normal_stage(uint32_t smem_b, uint32_t p_div8, uint32_t j0, uint32_t k0,
                                                 uint32_t k1, int tig) {
  constexpr int kCalls = kBRows * 8;
#pragma unroll 4
  for (int e = tig; e < kCalls; e += kGroupThreads) {
    const int row = e % kBRows;
    const int c = e / kBRows;
    const uint4 v = normal_chunk<kF16>(p_div8 + c, j0 + row, k0, k1);
    st_shared_v4(smem_b + row * 128 + ((c ^ (row & 7)) << 4), v);
  }
}

template <int kCtaGroup, int kWarpsPerGroup>
__global__ void GADM_PROJ_BOUNDS
project_kernel(const __grid_constant__ CUtensorMap tmap_g, const Args a) {
  using C = Cfg<kCtaGroup>;
  using R = Roles<kWarpsPerGroup>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * C::kStages);
  const uint32_t tmem_empty_bar = tmem_full_bar + 8u;
  const uint32_t tmem_slot = tmem_empty_bar + 8u;
  auto smem_a = [&](int s, int acc) { return smem_base + s * C::kStageBytes + acc * C::kATileBytes; };
  auto smem_b = [&](int s) { return smem_base + s * C::kStageBytes + C::kABytes; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
  const uint32_t cid = blockIdx.x / kCtaGroup;
  const uint32_t n_clusters = gridDim.x / kCtaGroup;

  if (warp == R::kTmaWarp && lane == 0) prefetch_tensormap(&tmap_g);
  if (warp == R::kMmaWarp && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1 + kWarpsPerGroup * kCtaGroup);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 4 * kCtaGroup);
    fence_barrier_init();
  }
  if (warp == R::kAllocWarp) tmem_alloc<kCtaGroup>(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  if constexpr (kCtaGroup == 2) cluster_arrive_wait(); else __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto kb_begin = [&](uint32_t split) {
    return static_cast<uint32_t>((static_cast<uint64_t>(split) * a.nkb_total) / a.n_splits);
  };

  if (warp == R::kTmaWarp) {
    // ===================== TMA producer: staged gradient tiles (A operand)
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t u = cid; u < a.n_units; u += n_clusters) {
        const uint32_t split = u / a.n_tiles;
        const uint32_t kb0 = kb_begin(split), kb1 = kb_begin(split + 1);
        for (uint32_t kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u, 0x100 + s);
          if (rank == 0 && it != 0 && it < a.sync_iters && (it % a.sync_every) == 0)
            grid_lockstep(a.sync_counter, (it / a.sync_every) * n_clusters);
          if (a.debug & 2u) {  // ablation: no loads, operands are whatever the slot holds
            if (rank == 0) mbar_arrive(full_bar(s));
            continue;
          }
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), a.n_acc * C::kATileBytes * kCtaGroup);
          for (uint32_t acc = 0; acc < a.n_acc; ++acc) {
            const int32_t row = acc * (kAccRows * kCtaGroup) + rank * kAccRows;
            // staged layout [D_pad/64][m_cap][64]: one 128-row tile is 16 KiB of contiguous HBM
            if constexpr (kCtaGroup == 2)
              tma_load_3d_cg2(smem_a(s, acc), &tmap_g, mapa(full_bar(s), 0), 0, row, kb);
            else
              tma_load_3d(smem_a(s, acc), &tmap_g, full_bar(s), 0, row, kb);
          }
        }
      }
    }
  } else if (warp == R::kMmaWarp) {
    // ===================== MMA issuer (leader CTA, one thread)
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc(a.a_fmt, kAccRows * kCtaGroup, kTileN);
      uint32_t it = 0, seg_iter = 0;
      for (uint32_t u = cid; u < a.n_units; u += n_clusters) {
        const uint32_t split = u / a.n_tiles;
        const uint32_t kb0 = kb_begin(split), kb1 = kb_begin(split + 1);
        for (uint32_t seg0 = kb0, seg1; seg0 < kb1; seg0 = seg1, ++seg_iter) {
          seg1 = seg_end(seg0, kb1, a.seg_kb);
          if (seg_iter > 0) mbar_wait(tmem_empty_bar, (seg_iter - 1) & 1u, 0x200);
          tcgen05_fence_after();
          for (uint32_t kb = seg0; kb < seg1; ++kb, ++it) {
            const int s = it % C::kStages;
            const uint32_t ph = (it / C::kStages) & 1u;
            mbar_wait(full_bar(s), ph, 0x300 + s);
            tcgen05_fence_after();
            const uint64_t bdesc = umma_desc_kmajor_sw128(smem_b(s));
            for (uint32_t acc = 0; acc < a.n_acc; ++acc) {
              const uint64_t adesc = umma_desc_kmajor_sw128(smem_a(s, acc));
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                // advancing K by 16 bf16 = 32 B inside the 128B swizzle row: +2 in the (addr >> 4) field
                umma_f16<kCtaGroup>(tmem_base + acc * kTileN, adesc + 2u * k, bdesc + 2u * k, idesc,
                                    (kb > seg0 || k > 0) ? 1u : 0u);
              }
            }
            if constexpr (kCtaGroup == 2) umma_commit_cg2_mcast(empty_bar(s), 0x3); else umma_commit(empty_bar(s));
          }
          if constexpr (kCtaGroup == 2) umma_commit_cg2_mcast(tmem_full_bar, 0x3); else umma_commit(tmem_full_bar);
        }
      }
    }
  } else if (warp >= R::kFirstEpiWarp && warp < R::kFirstEpiWarp + 4) {
    // ===================== epilogue: TMEM -> registers -> split-K partial tile
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const uint64_t pol = 0;  // no L2 hint in the pair kernel (see red_add_v4)
    uint32_t seg_iter = 0;
    for (uint32_t u = cid; u < a.n_units; u += n_clusters) {
      const uint32_t split = u / a.n_tiles;
      const uint32_t kb0 = kb_begin(split), kb1 = kb_begin(split + 1);
      uint32_t seg = 0;
      for (uint32_t seg0 = kb0; seg0 < kb1; seg0 = seg_end(seg0, kb1, a.seg_kb), ++seg, ++seg_iter) {
        mbar_wait<kEpiBackoffNs>(tmem_full_bar, seg_iter & 1u, 0x400);
        tcgen05_fence_after();
        for (uint32_t acc = 0; acc < a.n_acc; ++acc) {
          const uint32_t row = acc * (kAccRows * kCtaGroup) + rank * kAccRows + q * 32;
          float* dst = a.partial + (static_cast<size_t>(u) * a.unit_rows + row) * kTileN;
          const float* sc = a.inv_scale ? a.inv_scale + static_cast<size_t>(row) * a.scale_groups + seg0 / a.group_kb : nullptr;
          drain_accumulator(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kTileN, dst, seg == 0,
                            bar_base + C::kBarBytes + q * kEpiWarpBytes, lane, pol, sc, a.scale_groups,
                            static_cast<int>(a.m_rows) - static_cast<int>(row));
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCtaGroup == 2) mbar_arrive_cluster(mapa(tmem_empty_bar, 0)); else mbar_arrive(tmem_empty_bar);
        }
      }
    }
  } else if (warp < R::kGenWarps) {
    // ===================== generators: P tile (B operand) straight into the UMMA smem layout
    const int gw = warp;
    const int group = gw / kWarpsPerGroup;  // groups >= kGenGroups (single-CTA variant) stay idle
    const int tig = (gw % kWarpsPerGroup) * 32 + lane;
    uint32_t it = 0;
    for (uint32_t u = cid; u < a.n_units; u += n_clusters) {
      const uint32_t split = u / a.n_tiles;
      const uint32_t tile = u % a.n_tiles;
      const uint32_t j0 = tile * kTileN + rank * C::kBRows;
      const uint32_t kb0 = kb_begin(split), kb1 = kb_begin(split + 1);
      for (uint32_t kb = kb0; kb < kb1; ++kb, ++it) {
        if (static_cast<int>(it % C::kGenGroups) != group) continue;
        const int s = group;
        const uint32_t ph = (it / C::kStages) & 1u;
        mbar_wait<kGenBackoffNs>(empty_bar(s), ph ^ 1u, 0x500 + s);
        const uint32_t p_div64 = a.p_base_div64 + kb;
        if (a.debug & 1u) {
          // ablation: publish the slot without generating
        } else if (a.proj_type == kProjRademacher)
          gen_rademacher_stage<C::kBRows, R::kGroupThreads>(smem_b(s), p_div64 * 2u, j0, a.key0, a.key1, tig, lane,
                                                            a.a_fmt == UMMA_FMT_F16 ? kOnesF16 : kOnesBf16);
        else if (a.a_fmt == UMMA_FMT_F16)
          gen_normal_stage<C::kBRows, R::kGroupThreads, true>(smem_b(s), p_div64 * 8u, j0, a.key0, a.key1, tig);
        else
          gen_normal_stage<C::kBRows, R::kGroupThreads, false>(smem_b(s), p_div64 * 8u, j0, a.key0, a.key1, tig);
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the UMMA (async proxy) reads
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCtaGroup == 2) mbar_arrive_cluster(mapa(full_bar(s), 0)); else mbar_arrive(full_bar(s));
        }
      }
    }
  }

  // ===================== teardown
  __syncwarp();  // single-lane roles: reconverge before the .aligned barriers below
  tcgen05_fence_before();
  if constexpr (kCtaGroup == 2) cluster_arrive_wait(); else __syncthreads();
  if (warp == R::kAllocWarp) tmem_dealloc<kCtaGroup>(tmem_base, kTmemCols);
}

// out[m, tile*256 + c] (+)= sum over splits of partial[split*n_tiles + tile][m][c], fixed order.
__global__ void project_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int64_t ld_out,
                                      uint32_t M, uint32_t n_tiles, uint32_t n_splits, uint32_t unit_rows,
                                      int accumulate) {
  const uint32_t col4 = blockIdx.x * blockDim.x + threadIdx.x;  // float4 column index over k/4
  const uint32_t m = blockIdx.y;
  if (col4 >= n_tiles * (kTileN / 4) || m >= M) return;
  const uint32_t tile = col4 / (kTileN / 4);
  const uint32_t c4 = col4 % (kTileN / 4);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t s = 0; s < n_splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(
        partial + (static_cast<size_t>(s * n_tiles + tile) * unit_rows + m) * kTileN + c4 * 4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float* o = out + static_cast<size_t>(m) * ld_out + tile * kTileN + c4 * 4;
  if (accumulate) { o[0] += acc.x; o[1] += acc.y; o[2] += acc.z; o[3] += acc.w; }
  else { o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w; }
}

// Oracle hook: P[row0 + r, j] for r < nrows, j < k as fp32, using the very device functions the
// projection kernel uses (so a host-side G @ P reproduces the kernel up to summation order).
__global__ void materialize_p_kernel(float* __restrict__ out, int64_t row0, int64_t nrows, int64_t k, uint32_t key0,
                                     uint32_t key1, int proj_type, int f16) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nrows * k) return;
  const int64_t r = idx / k, j = idx % k;
  const uint64_t p = static_cast<uint64_t>(row0 + r);
  float v;
  if (proj_type == kProjRademacher) {
    const uint4 w = rademacher_call(static_cast<uint32_t>(p >> 5), static_cast<uint32_t>(j >> 2), key0, key1);
    const uint32_t words[4] = {w.x, w.y, w.z, w.w};
    const uint32_t word = words[j & 3];
    const uint32_t byte = (word >> (8 * ((p & 31) >> 3))) & 0xFFu;
    const uint4 e = rademacher_expand8(byte, f16 ? kOnesF16 : kOnesBf16);
    const uint32_t pk[4] = {e.x, e.y, e.z, e.w};
    const uint32_t h = (pk[(p & 7) >> 1] >> (16 * (p & 1))) & 0xFFFFu;
    v = f16 ? __half2float(__ushort_as_half(static_cast<unsigned short>(h))) : __uint_as_float(h << 16);
  } else {
    const uint4 c = f16 ? normal_chunk<true>(static_cast<uint32_t>(p >> 3), static_cast<uint32_t>(j), key0, key1)
                        : normal_chunk<false>(static_cast<uint32_t>(p >> 3), static_cast<uint32_t>(j), key0, key1);
    const uint32_t pk[4] = {c.x, c.y, c.z, c.w};
    const uint32_t h = (pk[(p & 7) >> 1] >> (16 * (p & 1))) & 0xFFFFu;
    v = f16 ? __half2float(__ushort_as_half(static_cast<unsigned short>(h))) : __uint_as_float(h << 16);
  }
  out[idx] = v;
}

// fp32 / bf16 / fp16 gradient block -> bf16 staging buffer in the tile-major layout
// staged[kb][row][c] (kb = p / 64, c = p % 64, row pitch 64, slab pitch m_cap * 64), p = col0 + i.
// src: [B, numel] with row pitch src_stride (elements); rows land at row0 + b.
template <typename T>
__global__ void pack_block_kernel(const T* __restrict__ src, int64_t src_stride, int64_t numel, int64_t B,
                                  __nv_bfloat16* __restrict__ dst, int64_t m_cap, int64_t row0, int64_t col0,
                                  float scale) {
  const int64_t b = blockIdx.y;
  const T* s = src + b * src_stride;
  const int64_t row = row0 + b;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = col0 + i;
    dst[((p >> 6) * m_cap + row) * 64 + (p & 63)] = __float2bfloat16_rn(static_cast<float>(s[i]) * scale);
  }
}

}  // namespace proj
}  // namespace gadm
