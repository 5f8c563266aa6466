// JL projection, "quad" variant: a cluster of FOUR CTAs = two tcgen05 CTA pairs that work on the same
// (256-column tile, D-split) unit for two different blocks of 512 staged rows and SHARE the generated P
// tiles.  Each CTA generates only half of its 128 P rows per pipeline slot (64 rows = 8 KiB, contiguous
// in the 128B-swizzled layout) and ships that half to the CTA of equal rank in the other pair with one
// cp.async.bulk shared::cta -> shared::cluster copy (async proxy on both ends, completion on an mbarrier
// of the destination CTA), so a Philox / Box-Muller element now feeds 1024 rows instead of 512.
// That halves the MUFU / issue load that bounds the normal projection (profiles/README.md); the price is
// cluster size 4 (33 clusters = 132 of 148 SMs co-resident on B200), so it is used for the normal type only.
//
// Barrier protocol per slot s (pair leader = even rank; "relay" = the otherwise idle MMA warp of the odd rank):
//   full[s]   (leader)   1 arrive.expect_tx by the leader's TMA thread (A bytes of both CTAs + the 8 KiB that the
//                        other pair ships into the leader's own smem) + one arrive per generator warp of the
//                        pair (2 CTAs x W warps) + 1 relay arrive
//   rfull[s]  (odd rank) the 8 KiB shipped into the odd-rank CTA complete here; the relay waits and forwards
//   empty[s]  (every CTA) 2 arrivals: tcgen05.commit of BOTH pair leaders, multicast to all four CTAs
#pragma once
#include "project.cuh"

namespace gadm {
namespace proj {

constexpr int kQuadOwnRows = 64;                     // P rows generated locally per slot per CTA
constexpr uint32_t kQuadShipBytes = kQuadOwnRows * kBlockK * 2;  // 8 KiB

template <int kRows, int kGroupThreads>
__device__ __forceinline__ void gen_rademacher_rows(uint32_t smem_b, int row_base, uint32_t p_div32, uint32_t j0,
                                                    uint32_t k0, uint32_t k1, int tig, int lane, uint32_t ones) {
  constexpr int kJGroups = kRows / 4;
  constexpr int kCalls = kJGroups * 2;
  const int rot = (lane >> 1) & 3;
  for (int c = tig; c < kCalls; c += kGroupThreads) {
    const int jg = row_base / 4 + c % kJGroups;
    const int pg = c / kJGroups;
    uint4 w = rademacher_call(p_div32 + pg, (j0 >> 2) + jg, k0, k1);
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

template <int kRows, int kGroupThreads, bool kF16>
__device__ __forceinline__ void gen_normal_rows(uint32_t smem_b, int row_base, uint32_t p_div8, uint32_t j0,
                                                uint32_t k0, uint32_t k1, int tig) {
  constexpr int kCalls = kRows * 8;
#pragma unroll 4
  for (int e = tig; e < kCalls; e += kGroupThreads) {
    const int row = row_base + e % kRows;
    const int c = e / kRows;
    const uint4 v = normal_chunk<kF16>(p_div8 + c, j0 + row, k0, k1);
    st_shared_v4(smem_b + row * 128 + ((c ^ (row & 7)) << 4), v);
  }
}

// kGroups generator groups of kGenWarps / kGroups warps: group g fills the k-blocks with it % kGroups == g.
// kGroups == 4: one group per pipeline slot (a slot's 8 KiB half tile takes one group ~4 k-block times);
// kGroups == 2: twice the warps per k-block, so a slot is ready in half the time -- the per-slot chain
// generate -> fence -> ship over DSMEM -> relay -> MMA -> commit has to fit into the 4-slot window.
template <int kWarpsPerGroup, int kGroups = kMaxGenGroups>
__global__ void GADM_PROJ_BOUNDS
project_quad_kernel(const __grid_constant__ CUtensorMap tmap_g, const Args a) {
  using C = Cfg<2>;
  using R = Roles<kWarpsPerGroup>;
  constexpr int kGroupWarps = R::kGenWarps / kGroups;
  constexpr int kGroupThreadsQ = kGroupWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto rfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (3 * C::kStages);
  const uint32_t tmem_empty_bar = tmem_full_bar + 8u;
  const uint32_t tmem_slot = tmem_empty_bar + 8u;
  auto smem_a = [&](int s, int acc) { return smem_base + s * C::kStageBytes + acc * C::kATileBytes; };
  auto smem_b = [&](int s) { return smem_base + s * C::kStageBytes + C::kABytes; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank4 = cluster_ctarank();
  const uint32_t pair = rank4 >> 1;           // which block of 512 rows
  const uint32_t rank = rank4 & 1u;           // position inside the tcgen05 CTA pair
  const uint32_t leader4 = pair * 2;          // cluster rank of this pair's leader
  const uint32_t partner4 = rank4 ^ 2u;       // same position in the other pair
  const uint32_t cid = blockIdx.x / 4;
  const uint32_t n_clusters = gridDim.x / 4;
  const uint32_t pair_rows = kNumAcc * kAccRows * 2;  // 512

  if (warp == R::kTmaWarp && lane == 0) prefetch_tensormap(&tmap_g);
  if (warp == R::kMmaWarp && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1 + kGroupWarps * 2 + 1);
      mbar_init(empty_bar(s), 2);
      mbar_init(rfull_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 4 * 2);
    fence_barrier_init();
  }
  if (warp == R::kAllocWarp) tmem_alloc<2>(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  cluster_arrive_wait();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto kb_begin = [&](uint32_t split) {
    return static_cast<uint32_t>((static_cast<uint64_t>(split) * a.nkb_total) / a.n_splits);
  };

  if (warp == R::kTmaWarp) {
    // ===================== TMA producer: this CTA's 2 x 128 staged gradient rows per slot
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t u = cid; u < a.n_units; u += n_clusters) {
        const uint32_t split = u / a.n_tiles;
        const uint32_t kb0 = kb_begin(split), kb1 = kb_begin(split + 1);
        for (uint32_t kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u, 0x2100 + s);
          if (rank4 == 0 && it != 0 && it < a.sync_iters && (it % a.sync_every) == 0)
            grid_lockstep(a.sync_counter, (it / a.sync_every) * n_clusters);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), kNumAcc * C::kATileBytes * 2 + kQuadShipBytes);
          for (uint32_t acc = 0; acc < kNumAcc; ++acc) {
            const int32_t row = pair * pair_rows + acc * (kAccRows * 2) + rank * kAccRows;
            tma_load_3d_cg2(smem_a(s, acc), &tmap_g, mapa(full_bar(s), leader4), 0, row, kb);
          }
        }
      }
    }
  } else if (warp == R::kMmaWarp) {
    if (lane == 0) {
      if (rank == 0) {
        // ===================== MMA issuer of this pair
        const uint32_t idesc = umma_idesc(a.a_fmt, kAccRows * 2, kTileN);
        const uint16_t pair_mask = static_cast<uint16_t>(0x3u << leader4);
        uint32_t it = 0, seg_iter = 0;
        for (uint32_t u = cid; u < a.n_units; u += n_clusters) {
          const uint32_t split = u / a.n_tiles;
          const uint32_t kb0 = kb_begin(split), kb1 = kb_begin(split + 1);
          for (uint32_t seg0 = kb0, seg1; seg0 < kb1; seg0 = seg1, ++seg_iter) {
            seg1 = seg_end(seg0, kb1, a.seg_kb);
            if (seg_iter > 0) mbar_wait(tmem_empty_bar, (seg_iter - 1) & 1u, 0x2200);
            tcgen05_fence_after();
            for (uint32_t kb = seg0; kb < seg1; ++kb, ++it) {
              const int s = it % C::kStages;
              const uint32_t ph = (it / C::kStages) & 1u;
              mbar_wait(full_bar(s), ph, 0x2300 + s);
              tcgen05_fence_after();
              const uint64_t bdesc = umma_desc_kmajor_sw128(smem_b(s));
              for (uint32_t acc = 0; acc < kNumAcc; ++acc) {
                const uint64_t adesc = umma_desc_kmajor_sw128(smem_a(s, acc));
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  umma_f16<2>(tmem_base + acc * kTileN, adesc + 2u * k, bdesc + 2u * k, idesc, (kb > seg0 || k > 0) ? 1u : 0u);
              }
              umma_commit_cg2_mcast(empty_bar(s), 0xF);  // both pairs must release a slot before anyone refills it
            }
            umma_commit_cg2_mcast(tmem_full_bar, pair_mask);
          }
        }
      } else {
        // ===================== relay (odd rank): the 8 KiB shipped into this CTA landed -> tell the pair leader
        uint32_t it = 0;
        for (uint32_t u = cid; u < a.n_units; u += n_clusters) {
          const uint32_t split = u / a.n_tiles;
          const uint32_t kb0 = kb_begin(split), kb1 = kb_begin(split + 1);
          for (uint32_t kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % C::kStages;
            const uint32_t ph = (it / C::kStages) & 1u;
            mbar_arrive_expect_tx(rfull_bar(s), kQuadShipBytes);
            mbar_wait(rfull_bar(s), ph, 0x2700 + s);
            mbar_arrive_cluster(mapa(full_bar(s), leader4));
          }
        }
      }
    }
  } else if (warp >= R::kFirstEpiWarp && warp < R::kFirstEpiWarp + 4) {
    // ===================== epilogue: TMEM -> registers -> split-K partial tile (rows of this pair)
    const int q = warp & 3;
    const uint64_t pol = l2_policy_evict_last();
    uint32_t seg_iter = 0;
    for (uint32_t u = cid; u < a.n_units; u += n_clusters) {
      const uint32_t split = u / a.n_tiles;
      const uint32_t kb0 = kb_begin(split), kb1 = kb_begin(split + 1);
      uint32_t seg = 0;
      for (uint32_t seg0 = kb0; seg0 < kb1; seg0 = seg_end(seg0, kb1, a.seg_kb), ++seg, ++seg_iter) {
        mbar_wait<kEpiBackoffNs>(tmem_full_bar, seg_iter & 1u, 0x2400);
        tcgen05_fence_after();
        for (uint32_t acc = 0; acc < kNumAcc; ++acc) {
          const uint32_t row = pair * pair_rows + acc * (kAccRows * 2) + rank * kAccRows + q * 32;
          float* dst = a.partial + (static_cast<size_t>(u) * a.unit_rows + row) * kTileN;
          const float* sc = a.inv_scale ? a.inv_scale + static_cast<size_t>(row) * a.scale_groups + seg0 / a.group_kb : nullptr;
          drain_accumulator(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kTileN, dst, seg == 0,
                            bar_base + C::kBarBytes + q * kEpiWarpBytes, lane, pol, sc, a.scale_groups,
                            static_cast<int>(a.m_rows) - static_cast<int>(row));
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa(tmem_empty_bar, leader4));
      }
    }
  } else if (warp < R::kGenWarps) {
    // ===================== generators: own 64 rows of the P tile -> local smem -> shipped to the partner CTA
    const int group = warp / kGroupWarps;
    const int tig = (warp % kGroupWarps) * 32 + lane;
    const int row_base = pair * kQuadOwnRows;
    uint32_t it = 0;
    for (uint32_t u = cid; u < a.n_units; u += n_clusters) {
      const uint32_t split = u / a.n_tiles;
      const uint32_t tile = u % a.n_tiles;
      const uint32_t j0 = tile * kTileN + rank * C::kBRows;
      const uint32_t kb0 = kb_begin(split), kb1 = kb_begin(split + 1);
      for (uint32_t kb = kb0; kb < kb1; ++kb, ++it) {
        if (static_cast<int>(it % kGroups) != group) continue;
        const int s = it % C::kStages;
        const uint32_t ph = (it / C::kStages) & 1u;
        mbar_wait<kGenBackoffNs>(empty_bar(s), ph ^ 1u, 0x2500 + s);
        const uint32_t p_div64 = a.p_base_div64 + kb;
        if (a.proj_type == kProjRademacher)
          gen_rademacher_rows<kQuadOwnRows, kGroupThreadsQ>(smem_b(s), row_base, p_div64 * 2u, j0, a.key0, a.key1, tig, lane,
                                                            a.a_fmt == UMMA_FMT_F16 ? kOnesF16 : kOnesBf16);
        else if (a.a_fmt == UMMA_FMT_F16)
          gen_normal_rows<kQuadOwnRows, kGroupThreadsQ, true>(smem_b(s), row_base, p_div64 * 8u, j0, a.key0, a.key1, tig);
        else
          gen_normal_rows<kQuadOwnRows, kGroupThreadsQ, false>(smem_b(s), row_base, p_div64 * 8u, j0, a.key0, a.key1, tig);
        fence_proxy_async_smem();                           // my generic writes -> async proxy (UMMA and the bulk copy)
        named_bar_sync(1 + group, kGroupThreadsQ);          // the whole 64-row half is written and fenced
        if (tig == 0) {
          const uint32_t src = smem_b(s) + row_base * 128;
          const uint32_t dst_bar = (rank == 0) ? full_bar(s) : rfull_bar(s);
          bulk_copy_smem_to_cluster(mapa(src, partner4), src, kQuadShipBytes, mapa(dst_bar, partner4));
        }
        if (lane == 0) mbar_arrive_cluster(mapa(full_bar(s), leader4));
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  cluster_arrive_wait();
  if (warp == R::kAllocWarp) tmem_dealloc<2>(tmem_base, kTmemCols);
}

}  // namespace proj
}  // namespace gadm
