// Gradient staging for the JL projection: per-example gradients (one tensor, or the dict of per-parameter
// tensors that vmap(grad(f)) returns) -> the projection kernel's 16-bit tile-major staging buffer, in ONE launch
// per batch, plus the fp32 timestep accumulator that precedes it.
//
// Replaces, on the reference's featurisation loop (src/attributions/methods/d_trak_grad.py:718-792,
// text_to_image/grad_text_to_image_lora.py:773-814):
//   vectorize_and_ignore_buffers (d_trak_grad.py:188-226: B x n_params flatten / cat launches + a B*D*4-byte copy),
//   emb += grads per timestep and emb / K (d_trak_grad.py:764-770),
//   the fp32 -> 16-bit conversion fast_jl performs inside project_*.
//
// Staging formats (include/gadm.h):
//   GADM_STAGE_BF16  bf16(v), no scaling (round-1 format);
//   GADM_STAGE_F16G  fp16(v * 2^s) with one power-of-two scale per (example row, group of 32768 columns), chosen so
//                    that the group's largest magnitude lands in [2^13, 2^14): 11 significant bits instead of 8 and
//                    no fp16 range problem (gradients span many decades between layers and examples).  The
//                    projection kernel multiplies each accumulation segment by the inverse scale (an exact power
//                    of two) when it promotes the TMEM accumulators, so results are independent of the scales up
//                    to fp16 rounding of the inputs.  A group is the unit because a kernel segment (256 / 512
//                    k-blocks) never crosses a group boundary.
//
// Bound: HBM.  Algorithmic bytes per staged element: sizeof(src) read + 2 written (the group is read twice --
// absmax pass, convert pass -- but it is 128 KiB per CTA, so the second read hits L2).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "gadm_ptx.cuh"

namespace gadm {
namespace stage {

constexpr int kGroupKb = 512;                 // k-blocks (of 64 columns) per scale group
constexpr int kGroupCols = kGroupKb * 64;     // 32768
// Two CTA shapes.  "narrow": 128 threads x <= 32 registers + a 12 KiB ring = what a resident projection CTA leaves
// free on its SM, so the CTA runs BESIDE the persistent projection kernel.  "wide": 256 threads, a 32 KiB ring -- it
// does not fit beside a projection CTA, so while the 4-CTA-cluster (quad) projection is running it is confined to
// the 16 SMs that grid cannot use (33 clusters = 132 of 148 SMs) and does not disturb the other 132 at all; alone it
// spreads over the whole GPU.
constexpr int kWideBulkDepth = 3, kWideBulkCols = 1024;
constexpr int kNarrowThreads = 128, kNarrowDepth = 3;
constexpr int kWideThreads = 256, kWideDepth = 4;
constexpr int kMaxBlocks = 1024;              // parameter blocks per launch (kernel-parameter space: 24 B each)

// Parameter blocks of one example, sorted by position in the flattened gradient.  Entry i covers columns
// [start[i], start[i + 1]); an entry with ptr == nullptr is a gap.  Gaps and columns >= start[n] are staged as zeros.
struct BlockTable {
  int32_t n;
  int32_t pad;
  int64_t start[kMaxBlocks + 1];  // start[n] = end of the last block
  int64_t stride[kMaxBlocks];     // elements between consecutive examples of the block
  const void* ptr[kMaxBlocks];
};

// index of the last block with start <= c (blocks are sorted); -1 if c lies before the first block
__device__ __noinline__ int find_block(const BlockTable& t, int64_t c) {
  int lo = 0, hi = t.n;  // invariant: start[lo - 1] <= c < start[hi] (with start[-1] = -inf)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t.start[mid] <= c) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p) { return static_cast<float>(*p); }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p) { return __half2float(*p); }

// Walks the columns [c_lo, c_hi) of example `ex` eight at a time on the global column grid (c_lo % 64 == 0, so an octet
// never straddles a 64-column row of the staging buffer): f(p, v[8]); columns no block covers read as 0.  Thread t
// takes octets t, t + blockDim.x, ...; the block that held the previous octet is cached, the table is only searched
// when an octet leaves it (~n_blocks times per row), and an octet that straddles a block boundary takes the
// per-element path.  Values are returned unscaled.
template <typename T>
struct BlockCursor {
  const BlockTable& tab;
  int64_t ex;
  int bi = -2;
  int64_t lo = 0, hi = 0;  // cached block covers [lo, hi)
  const T* base = nullptr;  // address of column lo of example ex
  __device__ __forceinline__ BlockCursor(const BlockTable& t, int64_t e) : tab(t), ex(e) {}
  __device__ __forceinline__ void seek(int64_t c) {
    bi = find_block(tab, c);
    if (bi >= 0 && tab.ptr[bi] != nullptr) {
      lo = tab.start[bi];
      hi = tab.start[bi + 1];
      base = reinterpret_cast<const T*>(tab.ptr[bi]) + ex * tab.stride[bi];
    } else {
      lo = hi = -1;
    }
  }
  __device__ __forceinline__ float at(int64_t c) {
    if (!(c >= lo && c < hi)) seek(c);
    return (c >= lo && c < hi) ? load_as_float(base + (c - lo)) : 0.f;
  }
  __device__ __forceinline__ void octet(int64_t p, float (&v)[8]) {
    if (!(p >= lo && p < hi)) seek(p);
    if (p >= lo && p + 8 <= hi) {
      const T* s = base + (p - lo);
      if (sizeof(T) == 4 && (reinterpret_cast<uintptr_t>(s) & 15) == 0) {
        const float4 q0 = reinterpret_cast<const float4*>(s)[0], q1 = reinterpret_cast<const float4*>(s)[1];
        v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
      } else if (sizeof(T) == 2 && (reinterpret_cast<uintptr_t>(s) & 15) == 0) {
        const uint4 q = *reinterpret_cast<const uint4*>(s);
        const T* h = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = load_as_float(h + i);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = load_as_float(s + i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = at(p + i);  // straddles a block boundary or a gap (rare)
    }
  }
};

// ---- asynchronous chunk stream -------------------------------------------------------------------------------------
// A staging CTA that runs beside the persistent projection kernel gets 4 warps and 32 registers per thread: with
// plain loads that is ~32 bytes in flight per thread, ~0.5 TB/s for the whole GPU, and staging (not the projection)
// sets the pace.  The loads therefore go through cp.async into a shared-memory ring (no registers held while in
// flight): a warp owns chunks of 256 consecutive columns, kDepth chunks ahead, copied with fully coalesced
// instructions -- 16-byte copies (512 contiguous bytes per instruction) when the chunk's source is 16-byte aligned,
// 4-byte copies (128 contiguous bytes per instruction) otherwise (fp32 rows of odd length) -- and lane l then reads
// columns 8 l .. 8 l + 7 of the chunk back.  Chunks that straddle a block boundary / gap are fetched synchronously.
constexpr int kChunk = 256;  // columns per warp per ring slot
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

template <typename T, int kThreads, int kDepth>
struct ChunkStream {
  static constexpr uint32_t kSlotBytes = kChunk * sizeof(T);  // per warp
  static constexpr uint32_t kRingBytes = kDepth * (kThreads / 32) * kSlotBytes;
  BlockCursor<T> cur;
  uint32_t ring;  // shared-memory address of this warp's slot 0; slot r lives (kThreads / 32) * kSlotBytes further
  int lane;
  __device__ __forceinline__ ChunkStream(const BlockTable& t, int64_t ex, uint32_t ring_base)
      : cur(t, ex), ring(ring_base + (threadIdx.x >> 5) * kSlotBytes), lane(threadIdx.x & 31) {}
  __device__ __forceinline__ uint32_t slot_addr(int r) const { return ring + r * ((kThreads / 32) * kSlotBytes); }

  // p0: first column of the chunk (the same for every lane of the warp)
  __device__ __forceinline__ void issue(int64_t p0, int r) {
    const uint32_t dst = slot_addr(r);
    if (!(p0 >= cur.lo && p0 < cur.hi)) cur.seek(p0);
    if (p0 >= cur.lo && p0 + kChunk <= cur.hi && (reinterpret_cast<uintptr_t>(cur.base + (p0 - cur.lo)) & 3) == 0) {
      const char* s = reinterpret_cast<const char*>(cur.base + (p0 - cur.lo));
      if ((reinterpret_cast<uintptr_t>(s) & 15) == 0) {
#pragma unroll
        for (uint32_t o = 0; o < kSlotBytes; o += 512) cp_async_16(dst + o + lane * 16, s + o + lane * 16);
      } else {
#pragma unroll
        for (uint32_t o = 0; o < kSlotBytes; o += 128) cp_async_4(dst + o + lane * 4, s + o + lane * 4);
      }
    } else {
      // straddles a block boundary / gap (or 2-byte elements at an odd offset): element i * 32 + lane, synchronously
#pragma unroll
      for (int i = 0; i < kChunk / 32; ++i) {
        const int64_t c = p0 + i * 32 + lane;
        if (!(c >= cur.lo && c < cur.hi)) cur.seek(c);
        const bool in_block = c >= cur.lo && c < cur.hi;
        const uint32_t d = dst + (i * 32 + lane) * sizeof(T);
        if constexpr (sizeof(T) == 4) {
          // asynchronous as well (part of this slot's commit group): a plain load here serialised 8 memory round
          // trips per boundary chunk and held the whole warp -- +25 % staging time on a 270-tensor block table
          if (in_block) cp_async_4(d, cur.base + (c - cur.lo));
          else asm volatile("st.shared.b32 [%0], %1;" ::"r"(d), "r"(0u) : "memory");
        } else {
          T v;
          if (in_block) v = cur.base[c - cur.lo];
          else v = static_cast<T>(0.f);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(d), "h"(*reinterpret_cast<const unsigned short*>(&v)) : "memory");
        }
      }
    }
    cp_async_commit();
  }

  // this lane's 8 columns (8 * lane ...) of the chunk in slot r
  __device__ __forceinline__ void consume(int r, float (&v)[8]) const {
    const uint32_t src = slot_addr(r) + lane * 8 * sizeof(T);
    if constexpr (sizeof(T) == 4) {
      uint4 q0, q1;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "r"(src) : "memory");
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "r"(src + 16) : "memory");
      v[0] = __uint_as_float(q0.x); v[1] = __uint_as_float(q0.y); v[2] = __uint_as_float(q0.z); v[3] = __uint_as_float(q0.w);
      v[4] = __uint_as_float(q1.x); v[5] = __uint_as_float(q1.y); v[6] = __uint_as_float(q1.z); v[7] = __uint_as_float(q1.w);
    } else {
      uint4 q;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(src) : "memory");
      const T* h = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = load_as_float(h + i);
    }
  }

  // f(p, v) for every octet of this warp's chunks; the warp takes the chunks s = warp, warp + kWarps, ... of a
  // sequence of n_seq chunks whose first columns are given by start(s) -- all chunks of the range in order, or the
  // sampled chunks of the scale guess -- with kDepth fetches in flight.  c_hi bounds the octets that are emitted.
  template <typename Start, typename F>
  __device__ __forceinline__ void for_each_chunk(int n_seq, int64_t c_hi, Start&& start, F&& f) {
    constexpr int kWarps = kThreads / 32;
    const int w = threadIdx.x >> 5;
    const int n = n_seq > w ? (n_seq - w + kWarps - 1) / kWarps : 0;
#pragma unroll
    for (int i = 0; i < kDepth; ++i)
      if (i < n) issue(start(w + i * kWarps), i);
    int r = 0;
    for (int i = 0; i < n; ++i) {
      if (n - i >= kDepth) cp_async_wait<kDepth - 1>(); else cp_async_wait<0>();
      __syncwarp();  // every lane's copies of this slot have landed
      float v[8];
      consume(r, v);
      const int64_t p = start(w + i * kWarps) + 8 * lane;
      if (p < c_hi) f(p, v);
      __syncwarp();  // every lane has read the slot before it is refilled
      if (i + kDepth < n) issue(start(w + (i + kDepth) * kWarps), r);
      r = (r + 1 == kDepth) ? 0 : r + 1;
    }
  }
  // every chunk of [c_lo, c_hi)  ((c_hi - c_lo) % 64 == 0)
  template <typename F>
  __device__ __forceinline__ void for_each(int64_t c_lo, int64_t c_hi, F&& f) {
    const int n_chunks = static_cast<int>((c_hi - c_lo + kChunk - 1) / kChunk);
    for_each_chunk(n_chunks, c_hi, [&](int s) { return c_lo + static_cast<int64_t>(s) * kChunk; }, f);
  }
  // the sample of the scale guess: the first 1024 of every 8192 columns of the range (chunks with id % 32 < 4) --
  // a fixed set of columns, independent of the CTA shape
  template <typename F>
  __device__ __forceinline__ void for_each_sampled(int64_t c_lo, int64_t c_hi, F&& f) {
    const int n_chunks = static_cast<int>((c_hi - c_lo + kChunk - 1) / kChunk);
    const int n_seq = (n_chunks / 32) * 4 + ((n_chunks % 32) < 4 ? (n_chunks % 32) : 4);
    for_each_chunk(n_seq, c_hi, [&](int s) { return c_lo + static_cast<int64_t>((s >> 2) * 32 + (s & 3)) * kChunk; }, f);
  }
};

// ---- TMA chunk stream (fp32 sources, wide CTAs) --------------------------------------------------------------------
// cp.async (LDGSTS) costs the SM 8 cycles per warp instruction whatever its width, i.e. 64 B/clk with 16-byte copies
// but 16 B/clk with the 4-byte copies that three out of four fp32 rows of odd length need -- ~40 GB/s per SM measured,
// which is what bounded staging whenever it had few SMs (beside a projection pass it gets 20).  Here the chunks are
// fetched by the TMA engine instead: one lane issues ONE 1-D bulk copy per 256-column chunk (mbarrier completion, no
// LSU issue slots, no registers).  Bulk copies need 16-byte aligned addresses and sizes, so a misaligned chunk is
// fetched as the aligned 1040 bytes that contain it and the consumer shifts by 1-3 words in registers (three
// LDS.128 and a warp-uniform select).  Chunks that are not interior to a block (the first / last few columns of a
// parameter tensor, gaps) are filled synchronously element by element as in ChunkStream.
template <int kThreads, int kDepth, int kCols>
struct BulkStream {
  static_assert(kDepth <= 8, "slot metadata is 4 bits per slot in one register");
  static_assert(kCols % kChunk == 0 && 1024 % kCols == 0, "a slot holds whole 256-column chunks and divides the sample unit");
  static constexpr int kWarps = kThreads / 32;
  static constexpr int kSub = kCols / kChunk;
  static constexpr uint32_t kSlotBytes = kCols * 4 + 16;
  static constexpr uint32_t kRingBytes = kDepth * kWarps * kSlotBytes;
  static constexpr uint32_t kSmemBytes = kRingBytes + kDepth * kWarps * 8;  // + one mbarrier per slot
  BlockCursor<float> cur;
  uint32_t ring;   // this warp's slot 0; slot r lives kWarps * kSlotBytes further
  uint32_t bars;   // this warp's mbarrier 0; slot r's is 8 * kWarps bytes further
  int lane;
  uint32_t phase = 0;  // bit r: parity the next wait on slot r expects
  uint32_t meta = 0;   // 4 bits per slot: bit 3 = bulk copy (wait on the mbarrier), bit 2 = element-wise cp.async group,
                       // bits 0-1 = word shift of a bulk copy
  __device__ __forceinline__ BulkStream(const BlockTable& t, int64_t ex, uint32_t smem_base)
      : cur(t, ex), ring(smem_base + (threadIdx.x >> 5) * kSlotBytes),
        bars(smem_base + kRingBytes + (threadIdx.x >> 5) * 8), lane(threadIdx.x & 31) {
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < kDepth; ++r) mbar_init(bar_addr(r), 1);
      fence_barrier_init();
    }
    __syncwarp();
  }
  __device__ __forceinline__ uint32_t slot_addr(int r) const { return ring + r * (kWarps * kSlotBytes); }
  __device__ __forceinline__ uint32_t bar_addr(int r) const { return bars + r * (kWarps * 8); }

  __device__ __forceinline__ void issue(int64_t p0, int r) {
    const uint32_t dst = slot_addr(r);
    if (!(p0 >= cur.lo && p0 < cur.hi)) cur.seek(p0);
    const bool inside = p0 >= cur.lo && p0 + kCols <= cur.hi;
    const float* s = cur.base + (p0 - cur.lo);
    const uint32_t sh = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(s) >> 2) & 3u;  // words past a 16-byte boundary
    // the aligned superset reads sh words before and 4 - sh words after the chunk: both must belong to this block's row
    const bool bulk = inside && (reinterpret_cast<uintptr_t>(s) & 3) == 0 &&
                      (sh == 0 || (p0 - cur.lo >= 4 && p0 + kCols + 4 <= cur.hi));
    uint32_t m;
    if (bulk) {
      if (lane == 0) {
        const uint32_t bytes = sh ? kSlotBytes : kCols * 4;
        mbar_arrive_expect_tx(bar_addr(r), bytes);
        bulk_copy_global_to_smem(dst, s - sh, bytes, bar_addr(r));
      }
      m = 8u | sh;
    } else {
      // not interior to a block (a tensor's first / last columns, tensors shorter than a slot, gaps): 4-byte
      // cp.async per element -- asynchronous like the bulk copy, tracked by a commit group instead of the mbarrier
#pragma unroll 4
      for (int i = 0; i < kCols / 32; ++i) {
        const int64_t c = p0 + i * 32 + lane;
        if (!(c >= cur.lo && c < cur.hi)) cur.seek(c);
        const uint32_t d = dst + (i * 32 + lane) * 4;
        if (c >= cur.lo && c < cur.hi) cp_async_4(d, cur.base + (c - cur.lo));
        else asm volatile("st.shared.b32 [%0], %1;" ::"r"(d), "r"(0u) : "memory");
      }
      cp_async_commit();
      m = 4u;
    }
    meta = (meta & ~(15u << (4 * r))) | (m << (4 * r));
  }

  // waits until slot r has landed; returns its word shift
  __device__ __forceinline__ uint32_t acquire(int r) {
    const uint32_t m = (meta >> (4 * r)) & 15u;
    if (m & 8u) {
      mbar_wait(bar_addr(r), (phase >> r) & 1u, 0x57a60000u | static_cast<uint32_t>(r));
      phase ^= 1u << r;
    } else if (m & 4u) {
      cp_async_wait<0>();  // groups complete in order: this slot's and (harmlessly) any later element-wise slot's
    }
    __syncwarp();  // element-wise slots: every lane's copies / zero stores are visible
    return m & 3u;
  }
  // this lane's 8 columns (sub * 256 + 8 * lane ...) of the chunk in slot r
  __device__ __forceinline__ void consume(int r, int sub, uint32_t sh, float (&v)[8]) const {
    const uint32_t src = slot_addr(r) + sub * (kChunk * 4) + lane * 32;
    uint32_t q[12];
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(src) : "memory");
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(src + 16) : "memory");
    if (sh == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(q[i]);
    } else {
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]) : "r"(src + 32) : "memory");
      if (sh == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(q[i + 1]);
      } else if (sh == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(q[i + 2]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(q[i + 3]);
      }
    }
  }

  // slots of kCols columns: the warp takes slots s = warp, warp + kWarps, ... of a sequence of n_seq
  template <typename Start, typename F>
  __device__ __forceinline__ void for_each_slot(int n_seq, int64_t c_hi, Start&& start, F&& f) {
    const int w = threadIdx.x >> 5;
    const int n = n_seq > w ? (n_seq - w + kWarps - 1) / kWarps : 0;
#pragma unroll
    for (int i = 0; i < kDepth; ++i)
      if (i < n) issue(start(w + i * kWarps), i);
    int r = 0;
    for (int i = 0; i < n; ++i) {
      const uint32_t sh = acquire(r);
      const int64_t p0 = start(w + i * kWarps);
#pragma unroll
      for (int sub = 0; sub < kSub; ++sub) {
        const int64_t p = p0 + sub * kChunk + 8 * lane;
        if (p < c_hi) {
          float v[8];
          consume(r, sub, sh, v);
          f(p, v);
        }
      }
      __syncwarp();  // every lane has read the slot before it is refilled
      if (i + kDepth < n) issue(start(w + (i + kDepth) * kWarps), r);
      r = (r + 1 == kDepth) ? 0 : r + 1;
    }
  }
  template <typename F>
  __device__ __forceinline__ void for_each(int64_t c_lo, int64_t c_hi, F&& f) {
    const int n_slots = static_cast<int>((c_hi - c_lo + kCols - 1) / kCols);
    for_each_slot(n_slots, c_hi, [&](int s) { return c_lo + static_cast<int64_t>(s) * kCols; }, f);
  }
  // the sample of the scale guess (see ChunkStream): the first 1024 of every 8192 columns
  template <typename F>
  __device__ __forceinline__ void for_each_sampled(int64_t c_lo, int64_t c_hi, F&& f) {
    constexpr int kPer = 1024 / kCols;   // sampled slots per 8192 columns
    constexpr int kEvery = 8192 / kCols;
    const int n_slots = static_cast<int>((c_hi - c_lo + kCols - 1) / kCols);
    const int n_seq = (n_slots / kEvery) * kPer + ((n_slots % kEvery) < kPer ? (n_slots % kEvery) : kPer);
    for_each_slot(n_seq, c_hi, [&](int s) { return c_lo + static_cast<int64_t>((s / kPer) * kEvery + (s % kPer)) * kCols; }, f);
  }
};

template <int kThreads>
__device__ __forceinline__ float block_max(float v, float* smem) {
  __syncthreads();  // the previous result has been read by everyone
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = smem[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) m = fmaxf(m, smem[w]);
  return m;
}

// scale exponent that puts a magnitude amax into [2^top, 2^(top + 1)); clamped so that both 2^s and 2^-s are normal
// fp32 numbers.  amax == 0 (or NaN / Inf, which then propagate) -> s = 0.
__device__ __forceinline__ int group_scale_exponent(float amax, int top) {
  if (!(amax > 0.f) || !(amax < __int_as_float(0x7f800000))) return 0;
  int e = static_cast<int>((__float_as_uint(amax) >> 23) & 0xffu) - 127;  // floor(log2(amax)) for normal amax
  if (e < -100) e = -100;
  if (e > 100) e = 100;
  return top - e;
}
__device__ __forceinline__ float exp2_int(int s) { return __uint_as_float(static_cast<uint32_t>(s + 127) << 23); }

// grid = (groups, batch).  CTA (g, b) stages columns [g * 32768, ...) of example b into row row0 + b.
//
// Scale of a group (F16G).  fp16 rounding is invariant under power-of-two scaling as long as nothing overflows or
// falls into the subnormal range, so the scale does not have to come from the exact group maximum -- it only has to
// put the maximum safely below 65504 and far above 2^-14.  The group is therefore staged in ONE pass with a scale
// GUESSED from a sample (columns [1024 j, 1024 j + 1024) of the group with j % 8 == 0, one eighth of it): sampled
// maximum -> [2^11, 2^12).  The pass computes the true maximum on the side; if the guess leaves it outside
// [2^8, 65504) (an outlier the sample missed, or an all-zero sample) the group is staged again with the exact
// scale (true maximum -> [2^13, 2^14)).  Deterministic: the scale is a function of the row's own values only.
// v * scale is never formed separately: |.| and rounding are monotonic, so max|v * scale| = |max|v| * scale|, and
// 2^s is folded into the multiplier (exact).
//
// Co-residency: the persistent projection kernel owns every SM (768 threads x 80 registers, ~215 KB of shared memory),
// which leaves 4096 registers, 1280 threads and ~15 KB of shared memory per SM.  A staging CTA is sized to fit into
// exactly that (128 threads, __maxnreg__(32), a 12 KiB cp.async ring), so that staging the next pass on another
// stream proceeds WHILE a pass is being projected instead of queueing behind it; alone, 16 such CTAs fill an SM.
// CTA (g, b) of the grid (groups, batch)
__device__ __forceinline__ int64_t stage_cta_example() { return blockIdx.y; }
template <typename Stream, bool kF16, int kThreads>
__device__ __forceinline__ void stage_groups_body(Stream& in, float* red, uint16_t* __restrict__ dst, int64_t m_cap,
                                                  int64_t row0, int64_t d_pad, float scale, float* __restrict__ inv_scale,
                                                  int64_t groups_per_row) {
  const int64_t g = blockIdx.x, b = stage_cta_example();
  const int64_t row = row0 + b;
  const int64_t c_lo = g * kGroupCols;
  const int64_t c_hi = (c_lo + kGroupCols < d_pad) ? c_lo + kGroupCols : d_pad;
  uint16_t* drow = dst + row * 64;
  float mul = scale;
  float amax = 0.f;
  auto convert_pass = [&]() {
    amax = 0.f;
    in.for_each(c_lo, c_hi, [&](int64_t p, const float (&v)[8]) {
      uint4 out;
      uint32_t* o = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if constexpr (kF16) {
          amax = fmaxf(amax, fmaxf(fabsf(v[2 * i]), fabsf(v[2 * i + 1])));  // NaNs are skipped here, stored as NaN
          const __half2 h = __floats2half2_rn(__fmul_rn(v[2 * i], mul), __fmul_rn(v[2 * i + 1], mul));
          o[i] = *reinterpret_cast<const uint32_t*>(&h);
        } else {
          const __nv_bfloat162 h = __floats2bfloat162_rn(__fmul_rn(v[2 * i], mul), __fmul_rn(v[2 * i + 1], mul));
          o[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
      *reinterpret_cast<uint4*>(drow + (p >> 6) * (m_cap * 64) + (p & 63)) = out;
    });
  };
  if constexpr (!kF16) {
    convert_pass();
    return;
  }
  in.for_each_sampled(c_lo, c_hi, [&](int64_t, const float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[i]));
  });
  int s = group_scale_exponent(block_max<kThreads>(amax, red) * fabsf(scale), 11);
  mul = scale * exp2_int(s);  // exact unless it leaves the normal range, which the exponent clamp excludes for
                              // |scale| in [2^-20, 2^20]
  convert_pass();
  const float true_max = block_max<kThreads>(amax, red) * fabsf(scale);
  const float landed = true_max * exp2_int(s);
  if (!(landed >= 256.f && landed < 65504.f) && !(true_max == 0.f)) {  // the sample misjudged the group: exact scale
    s = group_scale_exponent(true_max, 13);
    mul = scale * exp2_int(s);
    convert_pass();
  }
  if (threadIdx.x == 0) inv_scale[row * groups_per_row + g] = exp2_int(-s);
}

template <typename T, bool kF16>
__global__ void __maxnreg__(32)
stage_groups_kernel(const __grid_constant__ BlockTable tab, uint16_t* __restrict__ dst, int64_t m_cap, int64_t row0,
                    int64_t d_pad, float scale, float* __restrict__ inv_scale, int64_t groups_per_row) {
  using Stream = ChunkStream<T, kNarrowThreads, kNarrowDepth>;
  __shared__ float red[kNarrowThreads / 32];
  __shared__ __align__(16) uint8_t ring[Stream::kRingBytes];
  Stream in(tab, stage_cta_example(), static_cast<uint32_t>(__cvta_generic_to_shared(ring)));
  stage_groups_body<Stream, kF16, kNarrowThreads>(in, red, dst, m_cap, row0, d_pad, scale, inv_scale, groups_per_row);
}

// wide CTA: fp32 sources through the TMA stream, 2-byte sources through the cp.async stream; dynamic shared memory
template <typename T>
struct WideStream { using type = ChunkStream<T, kWideThreads, kWideDepth>; static constexpr uint32_t kSmemBytes = type::kRingBytes; };
template <>
struct WideStream<float> { using type = BulkStream<kWideThreads, kWideBulkDepth, kWideBulkCols>; static constexpr uint32_t kSmemBytes = type::kSmemBytes; };
template <typename T, bool kF16>
__global__ void __launch_bounds__(kWideThreads)
stage_groups_wide_kernel(const __grid_constant__ BlockTable tab, uint16_t* __restrict__ dst, int64_t m_cap, int64_t row0,
                         int64_t d_pad, float scale, float* __restrict__ inv_scale, int64_t groups_per_row) {
  using Stream = typename WideStream<T>::type;
  __shared__ float red[kWideThreads / 32];
  extern __shared__ __align__(16) uint8_t wide_ring[];
  Stream in(tab, stage_cta_example(), static_cast<uint32_t>(__cvta_generic_to_shared(wide_ring)));
  stage_groups_body<Stream, kF16, kWideThreads>(in, red, dst, m_cap, row0, d_pad, scale, inv_scale, groups_per_row);
}

// Timestep accumulator: slab[row0 + b, p] = (accumulate ? slab : 0) + scale * src   (fp32 slab [rows][d_pad]).
// grid = (ceil(d_pad / 8192), batch)
constexpr int kAccCols = 8192;
template <typename T>
__global__ void __launch_bounds__(kWideThreads)
accumulate_rows_kernel(const __grid_constant__ BlockTable tab, float* __restrict__ slab, int64_t d_pad, int64_t row0,
                       float scale, int accumulate) {
  const int64_t b = blockIdx.y;
  const int64_t c_lo = static_cast<int64_t>(blockIdx.x) * kAccCols;
  const int64_t c_hi = (c_lo + kAccCols < d_pad) ? c_lo + kAccCols : d_pad;
  float* out = slab + (row0 + b) * d_pad;
  __shared__ __align__(16) uint8_t ring[ChunkStream<T, kWideThreads, kWideDepth>::kRingBytes];
  ChunkStream<T, kWideThreads, kWideDepth> in(tab, b, static_cast<uint32_t>(__cvta_generic_to_shared(ring)));
  in.for_each(c_lo, c_hi, [&](int64_t p, const float (&v)[8]) {
    float4* o = reinterpret_cast<float4*>(out + p);  // d_pad % 64 == 0 and p % 8 == 0: 32-byte aligned
    // separate round-to-nearest multiply and add (no FMA contraction): bit-identical to emb += grads * scale in fp32
    float4 r0 = make_float4(__fmul_rn(v[0], scale), __fmul_rn(v[1], scale), __fmul_rn(v[2], scale), __fmul_rn(v[3], scale));
    float4 r1 = make_float4(__fmul_rn(v[4], scale), __fmul_rn(v[5], scale), __fmul_rn(v[6], scale), __fmul_rn(v[7], scale));
    if (accumulate) {
      const float4 a0 = o[0], a1 = o[1];
      r0.x = __fadd_rn(a0.x, r0.x); r0.y = __fadd_rn(a0.y, r0.y); r0.z = __fadd_rn(a0.z, r0.z); r0.w = __fadd_rn(a0.w, r0.w);
      r1.x = __fadd_rn(a1.x, r1.x); r1.y = __fadd_rn(a1.y, r1.y); r1.z = __fadd_rn(a1.z, r1.z); r1.w = __fadd_rn(a1.w, r1.w);
    }
    o[0] = r0;
    o[1] = r1;
  });
}

}  // namespace stage
}  // namespace gadm
