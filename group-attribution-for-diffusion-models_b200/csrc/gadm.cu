// C ABI (include/gadm.h) over the sm_100a kernels.  Single translation unit: the kernel headers are
// included here so that device-side globals (watchdog word) exist exactly once.
#include "../../include/gadm.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "gadm_ptx.cuh"
#include "philox.cuh"
#include "project.cuh"
#include "project_quad.cuh"
#include "stage.cuh"
#include <type_traits>
#include <utility>
#include <vector>

#include "aggregate.cuh"
#include "ridge.cuh"
#include "datamodel.cuh"
#include "gemm.cuh"
#include "trsv.cuh"

struct gadm_ctx {
  int device = 0;
  int num_sms = 0;
  int64_t launches = 0;
  uint32_t* scratch = nullptr;  // device scratch owned by the handle: ring of kLockstepSlots lockstep counters
  uint32_t lockstep_seq = 0;    // next ring slot: every projection launch gets its own counter word, so launches in
                                // flight on different streams of one device never share a barrier
  int quad_clusters = -1;       // co-resident clusters of 4 CTAs for the quad projection kernel (lazy)
  bool attr_gemm = false, attr_gemm_ts = false, attr_potrf = false;  // per-device kernel attributes already set
  bool attr_trsv = false;
  bool attr_gemm_ts2 = false;
  uint32_t attr_stage_wide = 0; // same, dynamic shared-memory opt-in of the wide staging kernels
  uint32_t attr_stage = 0;      // bit per staging-kernel instantiation whose carveout preference has been set
  cudaStream_t hp_stream = nullptr;  // high-priority stream for the Cholesky critical path (lazy)
  cudaStream_t hp_stream2 = nullptr; // second one: the next-block-column updates of the look-ahead
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // start / panel / col / end / rest (even, odd step)
  PFN_cuTensorMapEncodeTiled_v12000 encode_tiled = nullptr;
};

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define GADM_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(GADM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define GADM_REQUIRE(cond, ...)                       \
  do {                                                \
    if (!(cond)) return fail(GADM_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define GADM_TRY_RC(expr)           \
  do {                              \
    int _rc = (expr);               \
    if (_rc != GADM_OK) return _rc; \
  } while (0)

#define GADM_LAUNCHED(h)           \
  do {                             \
    GADM_CUDA(cudaGetLastError()); \
    (h)->launches++;               \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// fp32 [batch][rows][cols] (row pitch / batch stride in bytes), box = box_cols x box_rows x 1: GEMM operands
int make_tmap_3d_f32(gadm_handle h, CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                     uint64_t pitch_bytes, uint64_t batch_stride_bytes, uint32_t box_cols, uint32_t box_rows) {
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstride[2] = {pitch_bytes, batch_stride_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  GADM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor base %p is not 16-byte aligned", base);
  GADM_REQUIRE(pitch_bytes % 16 == 0 && batch_stride_bytes % 16 == 0, "pitch / batch stride must be multiples of 16 B");
  GADM_REQUIRE(box_cols * 4 == 128, "box inner extent must be 128 B for SWIZZLE_128B");
  CUresult r = h->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstride, box,
                               estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GADM_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return GADM_OK;
}

// staged gradients: bf16 [nkb][m_cap][64]; box = 64 x 128 x 1 (one contiguous 16 KiB tile)
int make_tmap_staged(gadm_handle h, CUtensorMap* map, const void* base, uint64_t m_cap, uint64_t nkb, bool f16) {
  cuuint64_t gdim[3] = {64, m_cap, nkb};
  cuuint64_t gstride[2] = {128, m_cap * 128};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  GADM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 127) == 0, "staging buffer %p is not 128-byte aligned", base);
  CUresult r = h->encode_tiled(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                               const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GADM_ERR_CUDA, "cuTensorMapEncodeTiled(staged) failed with CUresult %d", (int)r);
  return GADM_OK;
}

struct ProjPlan {
  uint32_t n_tiles, n_splits, n_units, nkb_total, n_acc, unit_rows, n_clusters;
  int64_t ws_bytes;
};

int quad_cluster_count(gadm_handle h);

int plan_projection(gadm_handle h, int64_t m_rows, int64_t d_pad, int64_t proj_dim, int cta_group, ProjPlan* p) {
  GADM_REQUIRE(cta_group == 1 || cta_group == 2 || cta_group == 4, "cta_group must be 1, 2 or 4, got %d", cta_group);
  GADM_REQUIRE(proj_dim > 0 && proj_dim % gadm::proj::kTileN == 0, "proj_dim %lld must be a positive multiple of 256",
               (long long)proj_dim);
  GADM_REQUIRE(d_pad > 0 && d_pad % gadm::proj::kBlockK == 0, "d_pad %lld must be a positive multiple of 64",
               (long long)d_pad);
  const int64_t max_rows = (int64_t)gadm::proj::kNumAcc * gadm::proj::kAccRows * cta_group;
  GADM_REQUIRE(m_rows > 0 && m_rows <= max_rows, "m_rows %lld out of range (1..%lld) for cta_group %d",
               (long long)m_rows, (long long)max_rows, cta_group);
  p->n_tiles = (uint32_t)(proj_dim / gadm::proj::kTileN);
  p->nkb_total = (uint32_t)(d_pad / gadm::proj::kBlockK);
  p->n_acc = (cta_group == 4 || m_rows > (int64_t)gadm::proj::kAccRows * cta_group) ? 2u : 1u;
  p->unit_rows = p->n_acc * gadm::proj::kAccRows * cta_group;
  if (cta_group == 4) {
    const int qc = quad_cluster_count(h);
    if (qc <= 0) return fail(GADM_ERR_CUDA, "no co-resident 4-CTA clusters available for the quad projection kernel");
    p->n_clusters = (uint32_t)qc;
    // a cluster count that is a multiple of the column-tile count gives every cluster exactly one (tile, D-split)
    // unit with the fewest splits (C2: 32 clusters = 16 tiles x 2 halves instead of 33 x 16 partial tiles through the
    // workspace) -- measured faster than the full 33 under the power cap, and it leaves whole SMs to the staging CTAs
    const uint32_t even = (uint32_t)qc / p->n_tiles * p->n_tiles;
    if (even > 0 && even * 16u >= (uint32_t)qc * 15u) p->n_clusters = even;
  } else {
    p->n_clusters = (uint32_t)(h->num_sms / cta_group);
  }
  // smallest D-split count whose unit count fills whole waves of clusters to >= 97% (units of one
  // wave cover consecutive splits x all column tiles, so a gradient tile is re-read from L2, not HBM)
  uint32_t best = 1;
  double best_eff = 0.0;
  const uint32_t max_splits = p->nkb_total < 128u ? p->nkb_total : 128u;
  for (uint32_t s = 1; s <= max_splits; ++s) {
    const uint64_t units = (uint64_t)p->n_tiles * s;
    const uint64_t waves = (units + p->n_clusters - 1) / p->n_clusters;
    const double eff = (double)units / (double)(waves * p->n_clusters);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    if (eff >= 0.97) break;
  }
  p->n_splits = best;
  p->n_units = p->n_tiles * p->n_splits;
  p->ws_bytes = (int64_t)p->n_units * p->unit_rows * gadm::proj::kTileN * (int64_t)sizeof(float);
  return GADM_OK;
}

constexpr uint32_t kLockstepSlots = 1024;  // ring of counter words in the handle scratch (4 KiB)

// Inter-cluster lockstep set-up shared by the pair and quad launchers.  Picks this launch's own counter word from
// the handle's ring (two projections in flight on different streams must not share a barrier), zeroes it on the
// launch stream and bounds the lockstep to the k-block count every launched cluster reaches (clusters own whole
// units round-robin).  The spin barrier needs every cluster resident at once, which only a cooperative launch
// guarantees: GADM_PROJ_COOPERATIVE=0 (plain launch, e.g. for profilers that cannot replay cooperative grids)
// therefore also switches the lockstep off unless GADM_PROJ_UNSAFE_LOCKSTEP=1 vouches that the kernel runs alone.
int setup_lockstep(gadm_handle h, gadm::proj::Args* a, uint32_t clusters, cudaStream_t stream, bool* cooperative) {
  uint64_t min_iters = ~0ull;
  for (uint32_t c = 0; c < clusters; ++c) {
    uint64_t iters = 0;
    for (uint32_t u = c; u < a->n_units; u += clusters) {
      const uint64_t split = u / a->n_tiles;
      iters += (split + 1) * a->nkb_total / a->n_splits - split * a->nkb_total / a->n_splits;
    }
    if (iters < min_iters) min_iters = iters;
  }
  const char* nosync = getenv("GADM_PROJ_NO_LOCKSTEP");
  const char* coop = getenv("GADM_PROJ_COOPERATIVE");
  const char* unsafe = getenv("GADM_PROJ_UNSAFE_LOCKSTEP");
  *cooperative = !(coop && atoi(coop) == 0);
  a->sync_every = gadm::proj::kSyncEvery;
  if (const char* e = getenv("GADM_PROJ_SYNC_EVERY")) { const int v = atoi(e); if (v >= 4) a->sync_every = (uint32_t)v; }
  a->sync_iters = (nosync && atoi(nosync)) ? 0u : (uint32_t)((min_iters / a->sync_every) * a->sync_every);
  if (!*cooperative && !(unsafe && atoi(unsafe))) a->sync_iters = 0;
  a->sync_counter = h->scratch + (h->lockstep_seq++ % kLockstepSlots);
  if (a->sync_iters) GADM_CUDA(cudaMemsetAsync(a->sync_counter, 0, sizeof(uint32_t), stream));
  return GADM_OK;
}

template <typename Kernel>
int launch_clusters(gadm_handle h, Kernel kernel, int cluster_size, int threads, int smem_bytes, const CUtensorMap& tmap,
                    const gadm::proj::Args& args, uint32_t n_clusters, cudaStream_t stream) {
  GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  cudaLaunchConfig_t cfg{};
  const uint32_t clusters = args.n_units < n_clusters ? args.n_units : n_clusters;
  cfg.gridDim = dim3(clusters * cluster_size);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;  // co-residency guarantee for the inter-cluster lockstep
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  gadm::proj::Args a = args;
  bool cooperative = true;
  int rc = setup_lockstep(h, &a, clusters, stream, &cooperative);
  if (rc != GADM_OK) return rc;
  // (A launch-completion event on this launch, waited on by the next staging launches so that the pass owns its SMs
  // first, was measured in round 2: the pass itself got 3 % slower with the attribute set -- not used.)
  auto launch = [&](bool coop) -> cudaError_t {
    cfg.numAttrs = coop ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, tmap, a);
  };
  if (a.sync_iters && cooperative) {
    if (launch(true) == cudaSuccess) {
      h->launches++;
      return GADM_OK;
    }
    (void)cudaGetLastError();  // cooperative launch refused (GPU shared / too large): run without the lockstep
    a.sync_iters = 0;
  }
  GADM_CUDA(launch(false));
  h->launches++;
  return GADM_OK;
}

template <int kCtaGroup, int kWarpsPerGroup>
int launch_project(gadm_handle h, const CUtensorMap& tmap, const gadm::proj::Args& args, uint32_t n_clusters,
                   cudaStream_t stream) {
  return launch_clusters(h, gadm::proj::project_kernel<kCtaGroup, kWarpsPerGroup>, kCtaGroup,
                         gadm::proj::Roles<kWarpsPerGroup>::kThreads, gadm::proj::Cfg<kCtaGroup>::kSmemBytes, tmap, args,
                         n_clusters, stream);
}

int quad_cluster_count(gadm_handle h) {
  if (h->quad_clusters >= 0) return h->quad_clusters;
  using C = gadm::proj::Cfg<2>;
  auto kernel = gadm::proj::project_quad_kernel<4>;
  int n = 0;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) == cudaSuccess) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(h->num_sms / 4 * 4));
    cfg.blockDim = dim3(gadm::proj::Roles<4>::kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
  } else {
    (void)cudaGetLastError();
  }
  if (const char* e = getenv("GADM_QUAD_CLUSTERS")) { const int v = atoi(e); if (v >= 1 && v < n) n = v; }  // experiment
  h->quad_clusters = n;
  return n;
}

template <int kWarpsPerGroup, int kGroups = gadm::proj::kMaxGenGroups>
int launch_project_quad(gadm_handle h, const CUtensorMap& tmap, const gadm::proj::Args& args, uint32_t n_clusters,
                        cudaStream_t stream) {
  return launch_clusters(h, gadm::proj::project_quad_kernel<kWarpsPerGroup, kGroups>, 4,
                         gadm::proj::Roles<kWarpsPerGroup>::kThreads, gadm::proj::Cfg<2>::kSmemBytes, tmap, args,
                         n_clusters, stream);
}

// host-side block list -> kernel-parameter table (sorted by position; gaps are staged as zeros)
int fill_block_table(const gadm_block* blocks, int n_blocks, int64_t batch, int64_t d_pad, gadm::stage::BlockTable* tab) {
  GADM_REQUIRE(blocks && n_blocks > 0, "empty block list");
  GADM_REQUIRE(n_blocks <= gadm::stage::kMaxBlocks, "%d parameter blocks exceed the %d one launch takes; concatenate "
               "neighbouring blocks first", n_blocks, gadm::stage::kMaxBlocks);
  int64_t prev_end = 0;
  int n = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const gadm_block& blk = blocks[b];
    if (blk.numel_per_example == 0) continue;
    GADM_REQUIRE(blk.ptr && blk.numel_per_example > 0 && blk.row_offset >= prev_end &&
                     blk.row_offset + blk.numel_per_example <= d_pad &&
                     (batch == 1 || blk.example_stride >= blk.numel_per_example),
                 "block %d (offset %lld, numel %lld, stride %lld) is unsorted, overlaps its predecessor or leaves the "
                 "staged gradient length %lld", b, (long long)blk.row_offset, (long long)blk.numel_per_example,
                 (long long)blk.example_stride, (long long)d_pad);
    if (blk.row_offset > prev_end) {  // gap: a table entry without a source, staged as zeros
      GADM_REQUIRE(n < gadm::stage::kMaxBlocks, "too many parameter blocks (gaps count as blocks)");
      tab->start[n] = prev_end;
      tab->stride[n] = 0;
      tab->ptr[n] = nullptr;
      ++n;
    }
    GADM_REQUIRE(n < gadm::stage::kMaxBlocks, "too many parameter blocks (gaps count as blocks)");
    tab->start[n] = blk.row_offset;
    tab->stride[n] = blk.example_stride;
    tab->ptr[n] = blk.ptr;
    prev_end = blk.row_offset + blk.numel_per_example;
    ++n;
    tab->start[n] = prev_end;  // columns from here to d_pad (or to the next block) read as zeros
  }
  GADM_REQUIRE(n > 0, "all blocks are empty");
  tab->n = n;
  tab->pad = 0;
  return GADM_OK;
}

inline int64_t stage_scale_count(int64_t d_pad) {
  return (d_pad / gadm::proj::kBlockK + gadm::stage::kGroupKb - 1) / gadm::stage::kGroupKb;
}

template <typename T>
int launch_stage(gadm_handle h, const gadm::stage::BlockTable& tab, int64_t batch, float scale, void* staged,
                 int stage_dtype, int64_t d_pad, int64_t m_cap, int64_t row0, float* inv_scale, int coresident,
                 cudaStream_t st) {
  const int64_t groups = stage_scale_count(d_pad);
  dim3 grid((unsigned)groups, (unsigned)batch);
  auto* dst = reinterpret_cast<uint16_t*>(staged);
  const bool f16 = stage_dtype == GADM_STAGE_F16G;
  float* sc = f16 ? inv_scale : nullptr;
  const uint32_t bit = 1u << (2 * sizeof(T) + (f16 ? 1 : 0) + (std::is_same<T, __half>::value ? 8 : 0));
  if (!coresident) {  // wide CTAs: the whole GPU when alone, the SMs a 4-CTA-cluster projection grid strands otherwise
    constexpr uint32_t smem = gadm::stage::WideStream<T>::kSmemBytes;
    auto opt_in = [&](auto kernel) -> int {
      if (h->attr_stage_wide & bit) return GADM_OK;
      GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      h->attr_stage_wide |= bit;
      return GADM_OK;
    };
    if (f16) {
      GADM_TRY_RC(opt_in(gadm::stage::stage_groups_wide_kernel<T, true>));
      gadm::stage::stage_groups_wide_kernel<T, true><<<grid, gadm::stage::kWideThreads, smem, st>>>(tab, dst, m_cap, row0, d_pad,
                                                                                                    scale, sc, groups);
    } else {
      GADM_TRY_RC(opt_in(gadm::stage::stage_groups_wide_kernel<T, false>));
      gadm::stage::stage_groups_wide_kernel<T, false><<<grid, gadm::stage::kWideThreads, smem, st>>>(tab, dst, m_cap, row0, d_pad,
                                                                                                     scale, sc, groups);
    }
    GADM_LAUNCHED(h);
    return GADM_OK;
  }
  // Narrow CTAs run beside a persistent projection CTA.  Same shared-memory carveout preference as that kernel
  // (maximum): an SM serves one carveout configuration at a time, and a staging CTA that asks for a small one would
  // wait for the projection CTA to leave instead of running beside it.
  auto prefer_max_smem = [&](auto kernel, uint32_t bit) -> int {
    if (h->attr_stage & bit) return GADM_OK;
    GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    h->attr_stage |= bit;
    return GADM_OK;
  };
  if (f16) {
    GADM_TRY_RC(prefer_max_smem(gadm::stage::stage_groups_kernel<T, true>, bit));
    gadm::stage::stage_groups_kernel<T, true><<<grid, gadm::stage::kNarrowThreads, 0, st>>>(tab, dst, m_cap, row0, d_pad, scale,
                                                                                            sc, groups);
  } else {
    GADM_TRY_RC(prefer_max_smem(gadm::stage::stage_groups_kernel<T, false>, bit));
    gadm::stage::stage_groups_kernel<T, false><<<grid, gadm::stage::kNarrowThreads, 0, st>>>(tab, dst, m_cap, row0, d_pad, scale,
                                                                                             sc, groups);
  }
  GADM_LAUNCHED(h);
  return GADM_OK;
}

}  // namespace

extern "C" {

int gadm_version(void) { return 100; }

const char* gadm_last_error(void) { return g_last_error.c_str(); }

int gadm_create(gadm_handle* out, int device) {
  if (!out) return fail(GADM_ERR_INVALID, "out is null");
  *out = nullptr;
  int count = 0;
  GADM_CUDA(cudaGetDeviceCount(&count));
  GADM_REQUIRE(device >= 0 && device < count, "device %d out of range (%d devices)", device, count);
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  GADM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(GADM_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)",
                device, prop.major, prop.minor);
  GADM_CUDA(cudaFree(0));  // make sure the primary context exists
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  GADM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(GADM_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  gadm_ctx* h = new gadm_ctx();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  if (cudaMalloc(&h->scratch, kLockstepSlots * sizeof(uint32_t)) != cudaSuccess) {
    delete h;
    return fail(GADM_ERR_CUDA, "cudaMalloc of the handle scratch failed");
  }
  *out = h;
  return GADM_OK;
}

int gadm_destroy(gadm_handle h) {
  if (h && h->scratch) cudaFree(h->scratch);
  if (h && h->hp_stream) cudaStreamDestroy(h->hp_stream);
  if (h && h->hp_stream2) cudaStreamDestroy(h->hp_stream2);
  if (h) for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  delete h;
  return GADM_OK;
}

int64_t gadm_launch_count(gadm_handle h) { return h ? h->launches : 0; }

int gadm_watchdog_code(gadm_handle h, unsigned int* code) {
  GADM_REQUIRE(h && code, "null argument");
  DeviceGuard guard(h->device);
  unsigned int zero = 0;
  GADM_CUDA(cudaMemcpyFromSymbol(code, gadm::g_watchdog_code, sizeof(unsigned int)));
  GADM_CUDA(cudaMemcpyToSymbol(gadm::g_watchdog_code, &zero, sizeof(unsigned int)));
  return GADM_OK;
}

int gadm_set_watchdog_ns(gadm_handle h, uint64_t ns) {
  GADM_REQUIRE(h, "null handle");
  DeviceGuard guard(h->device);
  unsigned long long v = ns;
  GADM_CUDA(cudaMemcpyToSymbol(gadm::g_watchdog_ns, &v, sizeof(v)));
  return GADM_OK;
}

int64_t gadm_project_workspace_bytes(gadm_handle h, int64_t m_rows, int64_t d_pad, int64_t proj_dim, int cta_group) {
  if (!h) return fail(GADM_ERR_INVALID, "null handle");
  ProjPlan p;
  int rc = plan_projection(h, m_rows, d_pad, proj_dim, cta_group, &p);
  if (rc != GADM_OK) return rc;
  return p.ws_bytes;
}

int gadm_pack_block(gadm_handle h, const void* src, int dtype, int64_t batch, int64_t numel, int64_t src_stride,
                    void* staged, int64_t d_pad, int64_t m_cap, int64_t row0, int64_t col0, float scale, void* stream) {
  GADM_REQUIRE(h && src && staged, "null argument");
  GADM_REQUIRE(batch > 0 && numel > 0 && src_stride >= numel && d_pad % 64 == 0 && d_pad >= col0 + numel && row0 >= 0 &&
                   col0 >= 0 && row0 + batch <= m_cap,
               "bad block geometry (batch %lld numel %lld stride %lld d_pad %lld m_cap %lld row0 %lld col0 %lld)",
               (long long)batch, (long long)numel, (long long)src_stride, (long long)d_pad, (long long)m_cap,
               (long long)row0, (long long)col0);
  DeviceGuard guard(h->device);
  const int threads = 256;
  int64_t bx = (numel + threads * 8 - 1) / (threads * 8);
  if (bx > 4096) bx = 4096;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)batch);
  auto* dst = reinterpret_cast<__nv_bfloat16*>(staged);
  if (dtype == GADM_DTYPE_F32)
    gadm::proj::pack_block_kernel<float><<<grid, threads, 0, as_stream(stream)>>>(
        reinterpret_cast<const float*>(src), src_stride, numel, batch, dst, m_cap, row0, col0, scale);
  else if (dtype == GADM_DTYPE_BF16)
    gadm::proj::pack_block_kernel<__nv_bfloat16><<<grid, threads, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(src), src_stride, numel, batch, dst, m_cap, row0, col0, scale);
  else if (dtype == GADM_DTYPE_F16)
    gadm::proj::pack_block_kernel<__half><<<grid, threads, 0, as_stream(stream)>>>(
        reinterpret_cast<const __half*>(src), src_stride, numel, batch, dst, m_cap, row0, col0, scale);
  else
    return fail(GADM_ERR_INVALID, "unknown dtype %d", dtype);
  GADM_CUDA(cudaGetLastError());
  h->launches++;
  return GADM_OK;
}

int64_t gadm_stage_scale_count(int64_t d_pad) { return stage_scale_count(d_pad); }

int gadm_stage_rows(gadm_handle h, const gadm_block* blocks, int n_blocks, int dtype, int64_t batch, float scale,
                    void* staged, int stage_dtype, int64_t d_pad, int64_t m_cap, int64_t row0, float* inv_scale,
                    int coresident, void* stream) {
  GADM_REQUIRE(h && staged, "null argument");
  GADM_REQUIRE(stage_dtype == GADM_STAGE_BF16 || stage_dtype == GADM_STAGE_F16G, "unknown stage_dtype %d", stage_dtype);
  GADM_REQUIRE(stage_dtype == GADM_STAGE_BF16 || inv_scale, "the F16G staging format needs the inv_scale array");
  GADM_REQUIRE(batch > 0 && batch < 65536 && d_pad > 0 && d_pad % 64 == 0 && row0 >= 0 && row0 + batch <= m_cap,
               "bad staging geometry (batch %lld d_pad %lld m_cap %lld row0 %lld)", (long long)batch, (long long)d_pad,
               (long long)m_cap, (long long)row0);
  GADM_REQUIRE((reinterpret_cast<uintptr_t>(staged) & 127) == 0, "staging buffer must be 128-byte aligned");
  static thread_local gadm::stage::BlockTable tab;  // 24 KiB: kept off the stack
  GADM_TRY_RC(fill_block_table(blocks, n_blocks, batch, d_pad, &tab));
  DeviceGuard guard(h->device);
  cudaStream_t st = as_stream(stream);
  if (dtype == GADM_DTYPE_F32) return launch_stage<float>(h, tab, batch, scale, staged, stage_dtype, d_pad, m_cap, row0, inv_scale, coresident, st);
  if (dtype == GADM_DTYPE_BF16) return launch_stage<__nv_bfloat16>(h, tab, batch, scale, staged, stage_dtype, d_pad, m_cap, row0, inv_scale, coresident, st);
  if (dtype == GADM_DTYPE_F16) return launch_stage<__half>(h, tab, batch, scale, staged, stage_dtype, d_pad, m_cap, row0, inv_scale, coresident, st);
  return fail(GADM_ERR_INVALID, "unknown dtype %d", dtype);
}

int gadm_accumulate_rows(gadm_handle h, const gadm_block* blocks, int n_blocks, int dtype, int64_t batch, float scale,
                         float* slab, int64_t d_pad, int64_t slab_rows, int64_t row0, int accumulate, void* stream) {
  GADM_REQUIRE(h && slab, "null argument");
  GADM_REQUIRE(batch > 0 && batch < 65536 && d_pad > 0 && d_pad % 64 == 0 && row0 >= 0 && row0 + batch <= slab_rows,
               "bad slab geometry (batch %lld d_pad %lld rows %lld row0 %lld)", (long long)batch, (long long)d_pad,
               (long long)slab_rows, (long long)row0);
  GADM_REQUIRE((reinterpret_cast<uintptr_t>(slab) & 31) == 0, "slab must be 32-byte aligned");
  static thread_local gadm::stage::BlockTable tab;
  GADM_TRY_RC(fill_block_table(blocks, n_blocks, batch, d_pad, &tab));
  DeviceGuard guard(h->device);
  cudaStream_t st = as_stream(stream);
  dim3 grid((unsigned)((d_pad + gadm::stage::kAccCols - 1) / gadm::stage::kAccCols), (unsigned)batch);
  if (dtype == GADM_DTYPE_F32)
    gadm::stage::accumulate_rows_kernel<float><<<grid, gadm::stage::kWideThreads, 0, st>>>(tab, slab, d_pad, row0, scale, accumulate);
  else if (dtype == GADM_DTYPE_BF16)
    gadm::stage::accumulate_rows_kernel<__nv_bfloat16><<<grid, gadm::stage::kWideThreads, 0, st>>>(tab, slab, d_pad, row0, scale, accumulate);
  else if (dtype == GADM_DTYPE_F16)
    gadm::stage::accumulate_rows_kernel<__half><<<grid, gadm::stage::kWideThreads, 0, st>>>(tab, slab, d_pad, row0, scale, accumulate);
  else
    return fail(GADM_ERR_INVALID, "unknown dtype %d", dtype);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_project_staged(gadm_handle h, const void* staged, int stage_dtype, const float* inv_scale, int64_t m_rows,
                        int64_t d_pad, int64_t m_cap, int64_t p_base, int64_t proj_dim, uint64_t seed64, int proj_type,
                        float* out, int64_t ld_out, int accumulate, void* workspace, int64_t workspace_bytes,
                        int cta_group, void* stream) {
  GADM_REQUIRE(h && staged && out && workspace, "null argument");
  GADM_REQUIRE(stage_dtype == GADM_STAGE_BF16 || stage_dtype == GADM_STAGE_F16G, "unknown stage_dtype %d", stage_dtype);
  GADM_REQUIRE(stage_dtype == GADM_STAGE_BF16 || inv_scale, "the F16G staging format needs the inv_scale array");
  GADM_REQUIRE(proj_type == GADM_PROJ_NORMAL || proj_type == GADM_PROJ_RADEMACHER, "unknown proj_type %d", proj_type);
  GADM_REQUIRE(p_base >= 0 && p_base % 64 == 0, "p_base %lld must be a non-negative multiple of 64", (long long)p_base);
  GADM_REQUIRE(m_cap >= m_rows, "m_cap %lld must be >= m_rows %lld", (long long)m_cap, (long long)m_rows);
  GADM_REQUIRE(ld_out >= proj_dim && ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "out must be 16-byte aligned with pitch %% 4 == 0");
  GADM_REQUIRE((p_base + d_pad) / 8 < (1ll << 32), "parameter index exceeds the 2^35 counter range");
  ProjPlan p;
  int rc = plan_projection(h, m_rows, d_pad, proj_dim, cta_group, &p);
  if (rc != GADM_OK) return rc;
  if (workspace_bytes < p.ws_bytes)
    return fail(GADM_ERR_WORKSPACE, "workspace %lld B < required %lld B", (long long)workspace_bytes,
                (long long)p.ws_bytes);
  GADM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  DeviceGuard guard(h->device);
  CUtensorMap tmap;
  rc = make_tmap_staged(h, &tmap, staged, (uint64_t)m_cap, (uint64_t)(d_pad / gadm::proj::kBlockK),
                        stage_dtype == GADM_STAGE_F16G);
  if (rc != GADM_OK) return rc;
  gadm::proj::Args a;
  a.partial = reinterpret_cast<float*>(workspace);
  a.n_tiles = p.n_tiles;
  a.n_splits = p.n_splits;
  a.n_units = p.n_units;
  a.nkb_total = p.nkb_total;
  a.n_acc = p.n_acc;
  a.unit_rows = p.unit_rows;
  a.key0 = (uint32_t)(seed64 & 0xFFFFFFFFull);
  a.key1 = (uint32_t)(seed64 >> 32);
  a.proj_type = (uint32_t)proj_type;
  a.p_base_div64 = (uint32_t)(p_base / 64);
  // k-blocks per TMEM accumulation segment (see project.cuh).  Measured at the C2 shape (fp64 reference, full D):
  // normal 256 -> 1.9e-5 relative error (unsegmented 1.3e-3), Rademacher 512 -> 1.3e-6; each costs ~2 % throughput.
  // GADM_PROJ_SEG_KB overrides for tuning.
  // fp16 inputs carry 11 significant bits, so each truncating TMEM add loses more than with bf16 inputs: Rademacher at
  // 512-k-block segments measured 7.8e-6 with F16G staging (1.3e-6 with bf16) -> 256 for both types there.
  a.seg_kb = (proj_type == gadm::kProjRademacher && stage_dtype == GADM_STAGE_BF16) ? 512u : 256u;
  if (const char* e = getenv("GADM_PROJ_SEG_KB")) { const long v = atol(e); if (v >= 1) a.seg_kb = (uint32_t)v; }
  a.m_rows = (uint32_t)m_rows;
  a.a_fmt = (stage_dtype == GADM_STAGE_F16G) ? gadm::UMMA_FMT_F16 : gadm::UMMA_FMT_BF16;
  a.inv_scale = (stage_dtype == GADM_STAGE_F16G) ? inv_scale : nullptr;
  a.scale_groups = (uint32_t)gadm_stage_scale_count(d_pad);
  a.group_kb = (uint32_t)gadm::stage::kGroupKb;
  if (a.inv_scale) {
    // a segment must not straddle two scale groups, and the scale groups are laid out from column 0 of the buffer
    GADM_REQUIRE(a.group_kb % a.seg_kb == 0, "segment length %u does not divide the scale group (%u k-blocks)", a.seg_kb,
                 a.group_kb);
  }
  {
    const char* dbg = getenv("GADM_PROJ_DEBUG");  // perf ablation only; results are garbage when set
    a.debug = dbg ? (uint32_t)atoi(dbg) : 0u;
  }
  cudaStream_t st = as_stream(stream);
  // generator warps per pipeline slot: 2 for Rademacher (cheap bits -> signs), 4 for the MUFU-heavy Box-Muller
  int gw = (proj_type == GADM_PROJ_NORMAL) ? 4 : 2;
  if (const char* e = getenv("GADM_PROJ_GEN_WARPS")) gw = (atoi(e) == 4) ? 4 : 2;  // tuning override
  if (cta_group == 4) {
    int groups = 4;  // generator groups of the quad kernel (project_quad.cuh); GADM_QUAD_GEN_GROUPS = 2 | 1 for tuning
    if (const char* e = getenv("GADM_QUAD_GEN_GROUPS")) groups = atoi(e);
    if (gw == 4 && groups == 2) rc = launch_project_quad<4, 2>(h, tmap, a, p.n_clusters, st);
    else if (gw == 4 && groups == 1) rc = launch_project_quad<4, 1>(h, tmap, a, p.n_clusters, st);
    else rc = (gw == 4) ? launch_project_quad<4>(h, tmap, a, p.n_clusters, st) : launch_project_quad<2>(h, tmap, a, p.n_clusters, st);
  }
  else if (cta_group == 2)
    rc = (gw == 4) ? launch_project<2, 4>(h, tmap, a, p.n_clusters, st) : launch_project<2, 2>(h, tmap, a, p.n_clusters, st);
  else
    rc = (gw == 4) ? launch_project<1, 4>(h, tmap, a, p.n_clusters, st) : launch_project<1, 2>(h, tmap, a, p.n_clusters, st);
  if (rc != GADM_OK) return rc;
  dim3 rgrid((unsigned)((proj_dim / 4 + 127) / 128), (unsigned)m_rows);
  gadm::proj::project_reduce_kernel<<<rgrid, 128, 0, st>>>(a.partial, out, ld_out, (uint32_t)m_rows, p.n_tiles,
                                                          p.n_splits, p.unit_rows, accumulate);
  GADM_CUDA(cudaGetLastError());
  h->launches++;
  return GADM_OK;
}

int gadm_materialize_p(gadm_handle h, int64_t row0, int64_t nrows, int64_t proj_dim, uint64_t seed64, int proj_type,
                       int stage_dtype, float* out, void* stream) {
  GADM_REQUIRE(h && out, "null argument");
  GADM_REQUIRE(row0 >= 0 && nrows > 0 && proj_dim > 0, "bad shape");
  GADM_REQUIRE(proj_type == GADM_PROJ_NORMAL || proj_type == GADM_PROJ_RADEMACHER, "unknown proj_type %d", proj_type);
  DeviceGuard guard(h->device);
  const int64_t total = nrows * proj_dim;
  const int threads = 256;
  gadm::proj::materialize_p_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
      out, row0, nrows, proj_dim, (uint32_t)(seed64 & 0xFFFFFFFFull), (uint32_t)(seed64 >> 32), proj_type,
      stage_dtype == GADM_STAGE_F16G ? 1 : 0);
  GADM_CUDA(cudaGetLastError());
  h->launches++;
  return GADM_OK;
}

// ------------------------------------------------------------------ scorer

#define GADM_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != GADM_OK) return _rc; \
  } while (0)

int gadm_gemm_tn_batched(gadm_handle h, const float* a, int64_t lda, int64_t stride_a, const float* b, int64_t ldb,
                         int64_t stride_b, float* c, int64_t ldc, int64_t stride_c, int64_t m, int64_t n, int64_t k,
                         int64_t batch, float alpha, float beta, float diag_add, int lower_only, void* stream) {
  GADM_REQUIRE(h && a && b && c, "null argument");
  GADM_REQUIRE(m > 0 && n > 0 && k > 0 && m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31), "bad shape");
  GADM_REQUIRE(lda >= k && ldb >= k && ldc >= n && lda % 4 == 0 && ldb % 4 == 0, "bad leading dimension");
  GADM_REQUIRE(batch >= 1 && batch < 65536, "bad batch count");
  if (batch > 1) GADM_REQUIRE(stride_a % 4 == 0 && stride_b % 4 == 0 && stride_a > 0 && stride_b > 0, "bad batch stride");
  DeviceGuard guard(h->device);
  CUtensorMap ta, tb;
  const uint64_t sa = (uint64_t)(batch > 1 ? stride_a : lda * m) * 4, sb = (uint64_t)(batch > 1 ? stride_b : ldb * n) * 4;
  GADM_TRY(make_tmap_3d_f32(h, &ta, a, (uint64_t)k, (uint64_t)m, (uint64_t)batch, (uint64_t)lda * 4, sa, gadm::gemm::kBK,
                            gadm::gemm::kBM));
  GADM_TRY(make_tmap_3d_f32(h, &tb, b, (uint64_t)k, (uint64_t)n, (uint64_t)batch, (uint64_t)ldb * 4, sb, gadm::gemm::kBK,
                            gadm::gemm::kBN));
  gadm::gemm::Args args;
  args.C = c; args.ldc = ldc; args.stride_c = batch > 1 ? stride_c : 0;
  args.M = (int32_t)m; args.N = (int32_t)n; args.K = (int32_t)k;
  args.alpha = alpha; args.beta = beta; args.diag_add = diag_add;
  args.lower_only = lower_only & 1;                                   // flag bit 0
  args.tri_b = (lower_only & 2) ? 1 : ((lower_only & 4) ? 2 : 0);      // bit 1: B lower-triangular, bit 2: upper
  // default: A operand staged in tensor memory (gemm.cuh, "TS" variant); GADM_GEMM_TS=0 selects the smem-smem kernel
  static const bool use_ts = [] { const char* e = getenv("GADM_GEMM_TS"); return !(e && atoi(e) == 0); }();
  dim3 grid((unsigned)((n + gadm::gemm::kBN - 1) / gadm::gemm::kBN), (unsigned)((m + gadm::gemm::kBM - 1) / gadm::gemm::kBM),
            (unsigned)batch);
  GADM_REQUIRE(grid.y < 65536, "too many row tiles (%u)", grid.y);
  // CTA-pair variant (256 x 256 tiles, N = 256 per MMA instruction): default whenever the problem has more than one
  // 128-tile in both directions; GADM_GEMM_2CTA=0 selects the single-CTA kernel everywhere
  static const bool use_2cta = [] { const char* e = getenv("GADM_GEMM_2CTA"); return !(e && atoi(e) == 0); }();
  // (long contractions only: for the rank-128 updates of the blocked Cholesky and the small merges of the triangular
  // inverse the larger tiles mean fewer, longer CTAs and measured slower -- 2.84 vs 2.67 ms for the factorisation)
  static const int64_t min_k = [] { const char* e = getenv("GADM_GEMM_2CTA_MINK"); return e ? (int64_t)atol(e) : (int64_t)2048; }();
  if (use_ts && use_2cta && m > gadm::gemm::kBM && n > gadm::gemm::kBN && k >= min_k) {
    // B box: the 128 rows of the 256-column tile that one CTA of the pair holds = the same tensor map as the 1-CTA kernel
    auto kernel = gadm::gemm::gemm_tn_3xtf32_ts2_kernel;
    if (!h->attr_gemm_ts2) {
      GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gadm::gemm::kT2SmemBytes));
      h->attr_gemm_ts2 = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2u * (unsigned)((m + 2 * gadm::gemm::kBM - 1) / (2 * gadm::gemm::kBM)),
                       (unsigned)((n + gadm::gemm::kT2BN - 1) / gadm::gemm::kT2BN), (unsigned)batch);
    cfg.blockDim = dim3(gadm::gemm::kT2Threads);
    cfg.dynamicSmemBytes = gadm::gemm::kT2SmemBytes;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GADM_REQUIRE(cfg.gridDim.y < 65536, "too many column tiles (%u)", cfg.gridDim.y);
    GADM_CUDA(cudaLaunchKernelEx(&cfg, kernel, ta, tb, args));
    h->launches++;
    return GADM_OK;
  }
  if (use_ts) {
    auto kernel = gadm::gemm::gemm_tn_3xtf32_ts_kernel;
    if (!h->attr_gemm_ts) {  // the attribute is per device: remembered in the handle, not in a process-wide static
      GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gadm::gemm::kTsSmemBytes));
      h->attr_gemm_ts = true;
    }
    kernel<<<grid, gadm::gemm::kThreads, gadm::gemm::kTsSmemBytes, as_stream(stream)>>>(ta, tb, args);
  } else {
    auto kernel = gadm::gemm::gemm_tn_3xtf32_kernel;
    if (!h->attr_gemm) {
      GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gadm::gemm::kSmemBytes));
      h->attr_gemm = true;
    }
    kernel<<<grid, gadm::gemm::kThreads, gadm::gemm::kSmemBytes, as_stream(stream)>>>(ta, tb, args);
  }
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_gemm_tn(gadm_handle h, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc,
                 int64_t m, int64_t n, int64_t k, float alpha, float beta, float diag_add, int lower_only,
                 void* stream) {
  return gadm_gemm_tn_batched(h, a, lda, 0, b, ldb, 0, c, ldc, 0, m, n, k, 1, alpha, beta, diag_add, lower_only, stream);
}

int gadm_transpose(gadm_handle h, const float* in, int64_t rows, int64_t cols, int64_t ld_in, float* out,
                   int64_t ld_out, void* stream) {
  GADM_REQUIRE(h && in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "bad argument");
  DeviceGuard guard(h->device);
  const int64_t rows_per_cta = (int64_t)gadm::gemm::kTrRows * gadm::gemm::kTrSub;
  dim3 grid((unsigned)((cols + gadm::gemm::kTrCols - 1) / gadm::gemm::kTrCols),
            (unsigned)((rows + rows_per_cta - 1) / rows_per_cta));
  GADM_REQUIRE(grid.y < 65536, "too many row tiles");
  GADM_CUDA(cudaFuncSetAttribute(gadm::gemm::transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 gadm::gemm::kTrSmemBytes));
  gadm::gemm::transpose_kernel<<<grid, dim3(32, 8), gadm::gemm::kTrSmemBytes, as_stream(stream)>>>(in, rows, cols, ld_in, out,
                                                                                                  ld_out);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int64_t gadm_cholesky_workspace_bytes(int64_t k) {
  const int64_t nblk = (k + gadm::gemm::kPotrfNb - 1) / gadm::gemm::kPotrfNb;
  return 2 * nblk * gadm::gemm::kPotrfNb * gadm::gemm::kPotrfNb * (int64_t)sizeof(float);
}

int gadm_cholesky(gadm_handle h, float* a, int64_t ld, int64_t k, void* blocks, int64_t blocks_bytes, int* info,
                  void* stream) {
  GADM_REQUIRE(h && a && blocks && k > 0 && ld >= k && ld % 4 == 0, "bad argument");
  if (blocks_bytes < gadm_cholesky_workspace_bytes(k))
    return fail(GADM_ERR_WORKSPACE, "blocks workspace %lld B < required %lld B", (long long)blocks_bytes,
                (long long)gadm_cholesky_workspace_bytes(k));
  DeviceGuard guard(h->device);
  constexpr int NB = gadm::gemm::kPotrfNb;
  const int64_t nblk = (k + NB - 1) / NB;
  float* linv = reinterpret_cast<float*>(blocks);
  float* linv_t = linv + nblk * NB * NB;
  auto potrf = gadm::gemm::potrf_diag_kernel;
  if (!h->attr_potrf) {
    GADM_CUDA(cudaFuncSetAttribute(potrf, cudaFuncAttributeMaxDynamicSharedMemorySize, gadm::gemm::kPotrfSmem));
    h->attr_potrf = true;
  }
  if (info) GADM_CUDA(cudaMemsetAsync(info, 0, sizeof(int), as_stream(stream)));
  // Look-ahead.  Step s of the right-looking factorisation is split into
  //   potrf(s)  diagonal block (+ its last rank-128 update A_ss -= L[s,s-1] L[s,s-1]^T, fused into the kernel)
  //   panel(s)  L[i,s] = A[i,s] L_ss^-T                        for i > s
  //   col(s)    A[i,s+1] -= L[i,s] L[s+1,s]^T                   for i >= s+2   (block column s+1 below its diagonal block)
  //   rest(s)   A[i,l]   -= L[i,s] L[l,s]^T                     for l >= s+2, i >= l (lower tiles)
  // potrf and panel run on a high-priority stream of the handle (the critical path: 46 + 16 us per step), col on a
  // second one and rest on the caller's stream, so that the small col(s) product (29 us, latency-bound) neither sits
  // on the critical path nor in front of rest(s).  Ordering (events are re-recorded every step; cudaStreamWaitEvent
  // captures the record that is current when it is called):
  //   col(s), rest(s)  after panel(s)                                               -- ev[1]
  //   panel(s+1)       after col(s)      (col(s) writes what it reads)              -- ev[2]
  //   potrf(s+2)       after rest(s)     (rest(s) writes A[s+2,s+2])                -- ev[4 + s % 2]
  //   col(s+1)         after rest(s)     (both write block column s+2)              -- ev[4 + s % 2]
  static const bool lookahead = [] { const char* e = getenv("GADM_CHOL_LOOKAHEAD"); return !(e && atoi(e) == 0); }();
  cudaStream_t user = as_stream(stream);
  cudaStream_t crit = user;
  if (lookahead && nblk > 2) {
    if (!h->hp_stream) {
      int least = 0, greatest = 0;
      GADM_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      GADM_CUDA(cudaStreamCreateWithPriority(&h->hp_stream, cudaStreamNonBlocking, greatest));
      GADM_CUDA(cudaStreamCreateWithPriority(&h->hp_stream2, cudaStreamNonBlocking, greatest));
      for (auto& e : h->ev) GADM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    crit = h->hp_stream;
    GADM_CUDA(cudaEventRecord(h->ev[0], user));
    GADM_CUDA(cudaStreamWaitEvent(crit, h->ev[0], 0));
    GADM_CUDA(cudaStreamWaitEvent(h->hp_stream2, h->ev[0], 0));
  }
  const bool split = crit != user;
  cudaStream_t cols = split ? h->hp_stream2 : user;
  bool any_col = false;
  bool col_pending = false;      // col(b-1) recorded in ev[2], not yet waited for by panel(b)
  bool rest_pending = false;     // rest(b-2) recorded in ev[4], not yet waited for by potrf(b)
  bool rest_prev = false;        // rest(b-1) recorded in ev[5]: becomes rest_pending at the next step
  for (int64_t b = 0; b < nblk; ++b) {
    const int64_t j0 = b * NB;
    const int nb = (int)((k - j0) < NB ? (k - j0) : NB);
    if (split && rest_pending) GADM_CUDA(cudaStreamWaitEvent(crit, h->ev[4 + (b & 1)], 0));  // rest(b-2) wrote A_bb
    const float* prev = (split && b > 0) ? a + j0 * ld + (j0 - NB) : nullptr;
    potrf<<<1, gadm::gemm::kPotrfThreads, gadm::gemm::kPotrfSmem, crit>>>(a + j0 * ld + j0, ld, nb, linv + b * NB * NB,
                                                                         linv_t + b * NB * NB, info, (int)b, prev);
    GADM_LAUNCHED(h);
    const int64_t rem = k - (j0 + nb);
    if (rem > 0) {
      float* panel = a + (j0 + nb) * ld + j0;
      if (split && col_pending) GADM_CUDA(cudaStreamWaitEvent(crit, h->ev[2], 0));  // col(b-1) updated this panel's input
      col_pending = false;
      // panel <- panel * L_jj^-T   (in place: one column tile, each CTA owns its rows)
      GADM_TRY(gadm_gemm_tn(h, panel, ld, linv + b * NB * NB, NB, panel, ld, rem, nb, nb, 1.f, 0.f, 0.f, 0, crit));
      float* trail = a + (j0 + nb) * ld + (j0 + nb);
      if (!split) {
        // trailing update (lower tiles only): A22 -= panel * panel^T
        GADM_TRY(gadm_gemm_tn(h, panel, ld, panel, ld, trail, ld, rem, rem, nb, -1.f, 1.f, 0.f, 1, user));
        continue;
      }
      GADM_CUDA(cudaEventRecord(h->ev[1], crit));  // panel(b) done
      GADM_CUDA(cudaStreamWaitEvent(user, h->ev[1], 0));
      const int64_t nn = rem < NB ? rem : NB;  // width of the next block column
      const bool had_rest = rest_prev;         // rest(b-1) exists (recorded in ev[4 + (b-1) % 2])
      rest_pending = rest_prev;                // what potrf(b+1) has to wait for is rest(b-1)
      rest_prev = false;
      if (rem > nn) {
        float* p2 = panel + nn * ld;  // panel rows from block b+2 on
        // col(b): block column b+1 below its diagonal block -= panel[b+2..] * panel[b+1]^T
        GADM_CUDA(cudaStreamWaitEvent(cols, h->ev[1], 0));
        if (had_rest) GADM_CUDA(cudaStreamWaitEvent(cols, h->ev[4 + ((b - 1) & 1)], 0));  // rest(b-1) wrote these blocks too
        GADM_TRY(gadm_gemm_tn(h, p2, ld, panel, ld, trail + nn * ld, ld, rem - nn, nn, nb, -1.f, 1.f, 0.f, 0, cols));
        GADM_CUDA(cudaEventRecord(h->ev[2], cols));
        col_pending = true;
        any_col = true;
        // rest(b): the lower triangle right of that column
        GADM_TRY(gadm_gemm_tn(h, p2, ld, p2, ld, trail + nn * ld + nn, ld, rem - nn, rem - nn, nb, -1.f, 1.f, 0.f, 1, user));
        GADM_CUDA(cudaEventRecord(h->ev[4 + (b & 1)], user));  // waited for by potrf(b+2): same parity
        rest_prev = true;
      }
    }
  }
  if (split) {  // join: everything after this call on the caller's stream sees the finished factor
    if (any_col) GADM_CUDA(cudaStreamWaitEvent(user, h->ev[2], 0));  // the last col(b) (panel(b+1) already waited for it)
    GADM_CUDA(cudaEventRecord(h->ev[3], crit));
    GADM_CUDA(cudaStreamWaitEvent(user, h->ev[3], 0));
  }
  return GADM_OK;
}

// x = (L L^T)^-1 b for one right-hand side, through the factor and the diagonal-block inverses gadm_cholesky left
// (csrc/trsv.cuh: one cooperative launch of k / 128 CTAs with point-to-point signalling).
int64_t gadm_cholesky_solve_vec_workspace_bytes(int64_t k) {
  const int64_t nblk = (k + 127) / 128;
  return 1024 + 2 * nblk * nblk * 128 * (int64_t)sizeof(double);  // 2 x nblk arrival counters (nblk <= 64), then the partial products
}

int gadm_cholesky_solve_vec(gadm_handle h, const float* l, int64_t ld, const void* blocks, int64_t k, const float* b,
                            float* x, void* workspace, int64_t workspace_bytes, void* stream) {
  GADM_REQUIRE(h && l && blocks && b && x && workspace && k > 0, "bad argument");
  GADM_REQUIRE(ld >= k && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(l) & 15) == 0,
               "gadm_cholesky_solve_vec needs ld %% 4 == 0 and a 16-byte aligned factor (k = %lld, ld = %lld)",
               (long long)k, (long long)ld);
  const int nblk = (int)((k + 127) / 128);
  GADM_REQUIRE(nblk <= h->num_sms && nblk <= 64, "k = %lld: the %d CTAs of the substitution must be co-resident (and <= 64)",
               (long long)k, nblk);
  if (workspace_bytes < gadm_cholesky_solve_vec_workspace_bytes(k))
    return fail(GADM_ERR_WORKSPACE, "workspace %lld B < required %lld B", (long long)workspace_bytes,
                (long long)gadm_cholesky_solve_vec_workspace_bytes(k));
  GADM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  DeviceGuard guard(h->device);
  cudaStream_t st = as_stream(stream);
  auto kernel = gadm::trsv::chol_solve_vec_kernel;
  if (!h->attr_trsv) {
    GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gadm::trsv::kSmemBytes));
    h->attr_trsv = true;
  }
  uint32_t* count = reinterpret_cast<uint32_t*>(workspace);
  double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 1024);
  GADM_CUDA(cudaMemsetAsync(count, 0, 1024, st));
  const float* linv = reinterpret_cast<const float*>(blocks);
  const float* linv_t = linv + (int64_t)nblk * 128 * 128;
  int nb = (int)k;  // the kernel derives the block count and the size of the last block from k
  void* args[] = {(void*)&l, (void*)&ld, (void*)&nb, (void*)&linv, (void*)&linv_t, (void*)&b, (void*)&x, (void*)&partial,
                  (void*)&count};
  GADM_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)nblk), dim3(gadm::trsv::kThreads), args,
                                        (size_t)gadm::trsv::kSmemBytes, st));
  GADM_LAUNCHED(h);
  return GADM_OK;
}

// L^-1 (and its transpose) of the factor left by gadm_cholesky, by recursive doubling over the diagonal blocks:
// for two adjacent diagonal groups with inverses X11, X22 and the factor's off-diagonal block L21,
//   X21 = -X22 (L21 X11)   -- three GEMMs per merge (T^T = X11^T L21^T, X21 = -X22 T, X21^T = -T^T X22^T).
// With L^-1 explicit, K^-1 is applied to any number of rows by two full-size GEMMs (y L^-T, then (.) L^-1) instead of
// 2 * k/128 dependent panel steps that each ran on m/128 CTAs (the blocked substitution of gadm_solve_rows was 5.5 ms
// of the 21 ms TRAK score at config 2 for 1000 generated images).
int64_t gadm_tri_inverse_workspace_bytes(int64_t k) {
  const int64_t kp = (k + 127) / 128 * 128;
  return kp * kp / 4 * (int64_t)sizeof(float) + 65536;
}

int gadm_tri_inverse(gadm_handle h, const float* l, int64_t ldl, const void* blocks, int64_t k, float* x, int64_t ldx,
                     float* xt, int64_t ldxt, void* workspace, int64_t workspace_bytes, void* stream) {
  GADM_REQUIRE(h && l && blocks && x && xt && workspace && k > 0, "bad argument");
  GADM_REQUIRE(ldl >= k && ldx >= k && ldxt >= k && ldl % 4 == 0 && ldx % 4 == 0 && ldxt % 4 == 0, "bad leading dimension");
  if (workspace_bytes < gadm_tri_inverse_workspace_bytes(k))
    return fail(GADM_ERR_WORKSPACE, "workspace %lld B < required %lld B", (long long)workspace_bytes,
                (long long)gadm_tri_inverse_workspace_bytes(k));
  DeviceGuard guard(h->device);
  constexpr int NB = gadm::gemm::kPotrfNb;
  const int64_t nblk = (k + NB - 1) / NB;
  const float* linv = reinterpret_cast<const float*>(blocks);
  const float* linv_t = linv + nblk * NB * NB;
  cudaStream_t st = as_stream(stream);
  GADM_CUDA(cudaMemset2DAsync(x, ldx * sizeof(float), 0, k * sizeof(float), k, st));
  GADM_CUDA(cudaMemset2DAsync(xt, ldxt * sizeof(float), 0, k * sizeof(float), k, st));
  gadm::gemm::tri_inverse_init_kernel<<<(unsigned)nblk, 256, 0, st>>>(linv, linv_t, k, x, ldx, xt, ldxt);
  GADM_LAUNCHED(h);
  float* work = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  // groups of diagonal blocks [g.first, g.second) in elements; adjacent pairs are merged level by level
  std::vector<std::pair<int64_t, int64_t>> groups;
  for (int64_t b = 0; b < nblk; ++b) groups.emplace_back(b * NB, (b + 1) * NB < k ? (b + 1) * NB : k);
  while (groups.size() > 1) {
    std::vector<std::pair<int64_t, int64_t>> next;
    int64_t woff = 0;
    // leading run of pairs whose four blocks all have the same full size b: one batched launch per product
    // (their X / Xt / L blocks lie 2b rows and 2b columns apart -> constant batch strides)
    const int64_t b = groups[0].second - groups[0].first;
    size_t uniform_pairs = 0;
    for (size_t g = 0; g + 1 < groups.size(); g += 2) {
      if (groups[g].second - groups[g].first != b || groups[g + 1].second - groups[g + 1].first != b) break;
      ++uniform_pairs;
    }
    if (uniform_pairs >= 2) {
      const int64_t a0 = groups[0].first, c0 = groups[1].first, P = (int64_t)uniform_pairs;
      float* tt = work;  // T^T of pair p at tt + p * b * b (pitch b)
      woff = P * b * b;
      GADM_TRY(gadm_gemm_tn_batched(h, xt + a0 * ldxt + a0, ldxt, 2 * b * ldxt + 2 * b, l + c0 * ldl + a0, ldl,
                                    2 * b * ldl + 2 * b, tt, b, b * b, b, b, b, P, 1.f, 0.f, 0.f, 0, stream));
      GADM_TRY(gadm_gemm_tn_batched(h, x + c0 * ldx + c0, ldx, 2 * b * ldx + 2 * b, tt, b, b * b, x + c0 * ldx + a0, ldx,
                                    2 * b * ldx + 2 * b, b, b, b, P, -1.f, 0.f, 0.f, 0, stream));
      GADM_TRY(gadm_gemm_tn_batched(h, tt, b, b * b, x + c0 * ldx + c0, ldx, 2 * b * ldx + 2 * b, xt + a0 * ldxt + c0, ldxt,
                                    2 * b * ldxt + 2 * b, b, b, b, P, -1.f, 0.f, 0.f, 0, stream));
      for (size_t p = 0; p < uniform_pairs; ++p) next.emplace_back(groups[2 * p].first, groups[2 * p + 1].second);
    } else {
      uniform_pairs = 0;
    }
    for (size_t g = 2 * uniform_pairs; g + 1 < groups.size(); g += 2) {
      const int64_t a0 = groups[g].first, b0 = groups[g].second - a0;       // first group: offset, size
      const int64_t c0 = groups[g + 1].first, b1 = groups[g + 1].second - c0;  // second group
      const int64_t ldt = (b1 + 3) / 4 * 4;
      float* tt = work + woff;  // T^T [b0, b1]
      woff += b0 * ldt;
      // T^T = X11^T L21^T = gemm_tn(Xt11 [b0, b0], L21 [b1, b0])
      GADM_TRY(gadm_gemm_tn(h, xt + a0 * ldxt + a0, ldxt, l + c0 * ldl + a0, ldl, tt, ldt, b0, b1, b0, 1.f, 0.f, 0.f, 0, stream));
      // X21 = -X22 T = -gemm_tn(X22 [b1, b1], T^T [b0, b1])
      GADM_TRY(gadm_gemm_tn(h, x + c0 * ldx + c0, ldx, tt, ldt, x + c0 * ldx + a0, ldx, b1, b0, b1, -1.f, 0.f, 0.f, 0, stream));
      // X21^T = -T^T X22^T = -gemm_tn(T^T [b0, b1], X22 [b1, b1])
      GADM_TRY(gadm_gemm_tn(h, tt, ldt, x + c0 * ldx + c0, ldx, xt + a0 * ldxt + c0, ldxt, b0, b1, b1, -1.f, 0.f, 0.f, 0, stream));
      next.emplace_back(a0, groups[g + 1].second);
    }
    if (groups.size() % 2) next.push_back(groups.back());
    groups.swap(next);
  }
  return GADM_OK;
}

int gadm_solve_rows(gadm_handle h, const float* l, int64_t ldl, const float* u, int64_t ldu, const void* blocks,
                    int64_t k, float* y, int64_t ldy, int64_t m, void* stream) {
  GADM_REQUIRE(h && l && u && blocks && y && k > 0 && m > 0, "bad argument");
  GADM_REQUIRE(ldl >= k && ldu >= k && ldy >= k && ldl % 4 == 0 && ldu % 4 == 0 && ldy % 4 == 0, "bad leading dimension");
  constexpr int NB = gadm::gemm::kPotrfNb;
  const int64_t nblk = (k + NB - 1) / NB;
  const float* linv = reinterpret_cast<const float*>(blocks);
  const float* linv_t = linv + nblk * NB * NB;
  // forward: y <- y L^-T
  for (int64_t b = 0; b < nblk; ++b) {
    const int64_t i0 = b * NB;
    const int64_t nb = (k - i0) < NB ? (k - i0) : NB;
    if (i0 > 0) GADM_TRY(gadm_gemm_tn(h, y, ldy, l + i0 * ldl, ldl, y + i0, ldy, m, nb, i0, -1.f, 1.f, 0.f, 0, stream));
    GADM_TRY(gadm_gemm_tn(h, y + i0, ldy, linv + b * NB * NB, NB, y + i0, ldy, m, nb, nb, 1.f, 0.f, 0.f, 0, stream));
  }
  // backward: y <- y L^-1
  for (int64_t b = nblk - 1; b >= 0; --b) {
    const int64_t i0 = b * NB;
    const int64_t nb = (k - i0) < NB ? (k - i0) : NB;
    const int64_t i1 = i0 + nb;
    if (i1 < k)
      GADM_TRY(gadm_gemm_tn(h, y + i1, ldy, u + i0 * ldu + i1, ldu, y + i0, ldy, m, nb, k - i1, -1.f, 1.f, 0.f, 0, stream));
    GADM_TRY(gadm_gemm_tn(h, y + i0, ldy, linv_t + b * NB * NB, NB, y + i0, ldy, m, nb, nb, 1.f, 0.f, 0.f, 0, stream));
  }
  return GADM_OK;
}

int gadm_row_norms(gadm_handle h, const float* x, int64_t rows, int64_t cols, int64_t ld, int reciprocal, float* out,
                   void* stream) {
  GADM_REQUIRE(h && x && out && rows > 0 && cols > 0 && ld >= cols, "bad argument");
  DeviceGuard guard(h->device);
  gadm::gemm::row_norms_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, rows, cols, ld, reciprocal,
                                                                                        out);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_matvec_rows(gadm_handle h, const float* x, int64_t rows, int64_t cols, int64_t ld, const float* v,
                     const float* col_scale, float* out, void* stream) {
  GADM_REQUIRE(h && x && v && out && rows > 0 && cols > 0 && ld >= cols, "bad argument");
  DeviceGuard guard(h->device);
  gadm::gemm::matvec_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, rows, cols, ld, v, col_scale,
                                                                                           out);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_diag_minmax(gadm_handle h, const float* l, int64_t ld, int64_t k, float* out2, void* stream) {
  GADM_REQUIRE(h && l && out2 && k > 0 && ld >= k, "bad argument");
  DeviceGuard guard(h->device);
  gadm::gemm::diag_minmax_kernel<<<1, 1024, 0, as_stream(stream)>>>(l, ld, k, out2);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_col_mean_scaled(gadm_handle h, const float* s, int64_t t, int64_t n, int64_t ld, const float* row_scale,
                         const float* col_scale, float* out, void* stream) {
  GADM_REQUIRE(h && s && out && t > 0 && n > 0 && ld >= n, "bad argument");
  DeviceGuard guard(h->device);
  gadm::gemm::col_mean_scaled_kernel<<<(unsigned)((n + 31) / 32), dim3(32, gadm::gemm::kColMeanLanes), 0, as_stream(stream)>>>(
      s, t, n, ld, row_scale, col_scale, out);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_scale_rows_cols(gadm_handle h, float* s, int64_t t, int64_t n, int64_t ld, const float* row_scale,
                         const float* col_scale, void* stream) {
  GADM_REQUIRE(h && s && t > 0 && n > 0 && ld >= n && t < 65536, "bad argument");
  DeviceGuard guard(h->device);
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)t);
  gadm::gemm::scale_rows_cols_kernel<<<grid, 256, 0, as_stream(stream)>>>(s, t, n, ld, row_scale, col_scale);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

// ------------------------------------------------------------------ aggregation

int gadm_pack_masks(gadm_handle h, const uint8_t* x, int64_t n, int64_t d, uint32_t* rowbits, uint32_t* colbits,
                    void* stream) {
  GADM_REQUIRE(h && x && rowbits && colbits && n > 0 && d > 0, "bad argument");
  DeviceGuard guard(h->device);
  const int64_t wd = (d + 31) / 32, wn = (n + 31) / 32;
  const int64_t total = n * wd + d * wn;
  gadm::agg::pack_masks_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, n, d, rowbits, wd,
                                                                                             colbits, wn);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_mask_gram(gadm_handle h, const uint32_t* colbits, int64_t n, int64_t d, int mode, double* a, void* stream) {
  GADM_REQUIRE(h && colbits && a && n > 0 && d > 0 && mode >= 0 && mode <= 2, "bad argument");
  DeviceGuard guard(h->device);
  gadm::agg::mask_gram_kernel<<<(unsigned)((d * d + 255) / 256), 256, 0, as_stream(stream)>>>(colbits, d, (n + 31) / 32,
                                                                                            n, mode, a);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_mask_xty(gadm_handle h, const uint32_t* rowbits, const double* y, int64_t n, int64_t d, int64_t k,
                  const double* shift, double half, double scale, double* out, void* stream) {
  GADM_REQUIRE(h && rowbits && y && out && n > 0 && d > 0 && k > 0, "bad argument");
  DeviceGuard guard(h->device);
  const int64_t wd = (d + 31) / 32;
  dim3 grid((unsigned)((k + gadm::agg::kXtyCols - 1) / gadm::agg::kXtyCols),
            (unsigned)((d + gadm::agg::kXtyPlayers - 1) / gadm::agg::kXtyPlayers));
  GADM_CUDA(cudaFuncSetAttribute(gadm::agg::mask_xty_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 gadm::agg::kXtySmemBytes));
  GADM_CUDA(cudaFuncSetAttribute(gadm::agg::mask_xty_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 gadm::agg::kXtySmemBytes));
  if (shift)
    gadm::agg::mask_xty_kernel<true><<<grid, gadm::agg::kXtyThreads, gadm::agg::kXtySmemBytes, as_stream(stream)>>>(
        rowbits, wd, y, n, d, k, shift, half, scale, out);
  else
    gadm::agg::mask_xty_kernel<false><<<grid, gadm::agg::kXtyThreads, gadm::agg::kXtySmemBytes, as_stream(stream)>>>(
        rowbits, wd, y, n, d, k, nullptr, half, scale, out);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_mask_times_matrix(gadm_handle h, const uint32_t* colbits, const double* mat, int64_t m, int64_t d, int64_t k,
                           double* out, void* stream) {
  GADM_REQUIRE(h && colbits && mat && out && m > 0 && d > 0 && k > 0, "bad argument");
  DeviceGuard guard(h->device);
  const int64_t wm = (m + 31) / 32;  // words per player in the column bit planes
  dim3 grid((unsigned)((k + gadm::agg::kXtyCols - 1) / gadm::agg::kXtyCols),
            (unsigned)((m + gadm::agg::kXtyPlayers - 1) / gadm::agg::kXtyPlayers));
  GADM_CUDA(cudaFuncSetAttribute(gadm::agg::mask_xty_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 gadm::agg::kXtySmemBytes));
  gadm::agg::mask_xty_kernel<false><<<grid, gadm::agg::kXtyThreads, gadm::agg::kXtySmemBytes, as_stream(stream)>>>(
      colbits, wm, mat, /*summed=*/d, /*outputs=*/m, k, nullptr, 0.0, 1.0, out);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int64_t gadm_sym_pinv_workspace_bytes(int64_t d) {
  const int64_t dp = (d + 1) & ~1ll;
  return (2 * dp * dp + dp) * (int64_t)sizeof(double);
}

int gadm_sym_pinv(gadm_handle h, const double* a, int64_t d, double rcond, double* out, void* workspace,
                  int64_t workspace_bytes, int* info, void* stream) {
  GADM_REQUIRE(h && a && out && workspace && d > 0 && d <= 8192, "bad argument");
  const int64_t need = gadm_sym_pinv_workspace_bytes(d);
  if (workspace_bytes < need)
    return fail(GADM_ERR_WORKSPACE, "workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)need);
  DeviceGuard guard(h->device);
  const int use_smem = need <= 200 * 1024;
  auto kernel = gadm::agg::sym_pinv_kernel;
  if (use_smem) GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
  kernel<<<1, gadm::agg::kPinvThreads, use_smem ? (size_t)need : 0, as_stream(stream)>>>(
      a, (int)d, rcond, out, reinterpret_cast<double*>(workspace), use_smem, info);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_dgemm_dk(gadm_handle h, const double* a, const double* b, int64_t d, int64_t k, double zero_below, double* c,
                  void* stream) {
  GADM_REQUIRE(h && a && b && c && d > 0 && k > 0, "bad argument");
  DeviceGuard guard(h->device);
  dim3 grid((unsigned)((k + 31) / 32), (unsigned)((d + gadm::agg::kDgemmTile - 1) / gadm::agg::kDgemmTile));
  gadm::agg::dgemm_dk_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(a, b, d, k, zero_below, c);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

// ------------------------------------------------------------------ RidgeCV datamodel estimator (lds.py:411-421)
int gadm_center_columns(gadm_handle h, const double* x, int64_t n, int64_t d, double* xc, double* mean, void* stream) {
  GADM_REQUIRE(h && x && xc && mean && n > 0 && d > 0, "bad argument");
  DeviceGuard guard(h->device);
  gadm::ridge::center_columns_kernel<<<(unsigned)((d + 63) / 64), 64, 0, as_stream(stream)>>>(x, n, d, xc, mean);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_dgemm(gadm_handle h, int trans_a, const double* a, int64_t lda, const double* b, int64_t ldb, int64_t m,
               int64_t j, int64_t n, double* c, int64_t ldc, void* stream) {
  GADM_REQUIRE(h && a && b && c && m > 0 && j > 0 && n > 0 && lda > 0 && ldb >= n && ldc >= n, "bad argument");
  DeviceGuard guard(h->device);
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((m + 31) / 32));
  if (trans_a) gadm::ridge::dgemm_kernel<true><<<grid, dim3(32, 8), 0, as_stream(stream)>>>(a, lda, b, ldb, m, j, n, c, ldc);
  else gadm::ridge::dgemm_kernel<false><<<grid, dim3(32, 8), 0, as_stream(stream)>>>(a, lda, b, ldb, m, j, n, c, ldc);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int64_t gadm_sym_eig_workspace_bytes(int64_t d) {
  const int64_t dp = (d + 1) & ~1ll;
  return 2 * dp * dp * (int64_t)sizeof(double);
}

int gadm_sym_eig(gadm_handle h, const double* a, int64_t d, double* evals, double* v, void* workspace,
                 int64_t workspace_bytes, int* info, void* stream) {
  GADM_REQUIRE(h && a && evals && v && workspace && d > 0 && d <= 8192, "bad argument");
  const int64_t need = gadm_sym_eig_workspace_bytes(d);
  if (workspace_bytes < need)
    return fail(GADM_ERR_WORKSPACE, "workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)need);
  DeviceGuard guard(h->device);
  const int use_smem = need <= 200 * 1024;
  auto kernel = gadm::ridge::sym_eig_kernel;
  if (use_smem) GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
  kernel<<<1, gadm::agg::kPinvThreads, use_smem ? (size_t)need : 0, as_stream(stream)>>>(
      a, (int)d, evals, v, reinterpret_cast<double*>(workspace), use_smem, info);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int64_t gadm_ridge_gcv_workspace_bytes(int64_t n, int64_t d, int64_t k, int64_t n_alphas) {
  const int64_t tiles = (n + gadm::ridge::kGcvRows - 1) / gadm::ridge::kGcvRows;
  // q [d] | den [A, n] | Z^T [d, n] | per-tile partial sums [A, tiles, k]; every part starts 16-byte aligned
  return ((d + 1) / 2 * 2 + (n_alphas * n + 1) / 2 * 2 + (d * n + 1) / 2 * 2 + n_alphas * tiles * k) * (int64_t)sizeof(double) + 256;
}

int gadm_ridge_gcv(gadm_handle h, const double* z, const double* t, const double* yc, const double* evals,
                   const double* alphas, int64_t n, int64_t d, int64_t k, int64_t n_alphas, void* workspace,
                   int64_t workspace_bytes, double* score, void* stream) {
  GADM_REQUIRE(h && z && t && yc && evals && alphas && workspace && score, "null argument");
  GADM_REQUIRE(n > 0 && d > 0 && k > 0 && n_alphas > 0, "bad size");
  if (workspace_bytes < gadm_ridge_gcv_workspace_bytes(n, d, k, n_alphas))
    return fail(GADM_ERR_WORKSPACE, "workspace %lld B < required %lld B", (long long)workspace_bytes,
                (long long)gadm_ridge_gcv_workspace_bytes(n, d, k, n_alphas));
  DeviceGuard guard(h->device);
  cudaStream_t st = as_stream(stream);
  const int64_t tiles = (n + gadm::ridge::kGcvRows - 1) / gadm::ridge::kGcvRows;
  double* q_work = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  double* den_work = q_work + (d + 1) / 2 * 2;
  double* zt = den_work + (n_alphas * n + 1) / 2 * 2;
  double* partial = zt + (d * n + 1) / 2 * 2;
  gadm::ridge::column_sums_kernel<<<(unsigned)((d + 63) / 64), 64, 0, st>>>(z, n, d, q_work);
  GADM_LAUNCHED(h);
  const int64_t jobs = n_alphas * n;
  gadm::ridge::ridge_denominator_kernel<<<(unsigned)((jobs + 7) / 8), 256, 0, st>>>(z, evals, q_work, alphas, n, d,
                                                                                    n_alphas, den_work);
  GADM_LAUNCHED(h);
  dim3 tgrid((unsigned)((d + 31) / 32), (unsigned)((n + 31) / 32));
  gadm::ridge::transpose_f64_kernel<<<tgrid, dim3(32, 8), 0, st>>>(z, n, d, zt);
  GADM_LAUNCHED(h);
  auto kernel = gadm::ridge::ridge_gcv_score_kernel;
  const int64_t gcv_smem = gadm::ridge::kGcvSmemBytes + d * (int64_t)sizeof(double);
  GADM_REQUIRE(gcv_smem <= 227 * 1024, "d = %lld: the w(alpha) table does not fit in shared memory", (long long)d);
  GADM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gcv_smem));
  dim3 grid((unsigned)((k + gadm::ridge::kGcvCols - 1) / gadm::ridge::kGcvCols), (unsigned)tiles);
  GADM_REQUIRE(grid.y < 65536, "too many row tiles");
  kernel<<<grid, gadm::ridge::kGcvThreads, (size_t)gcv_smem, st>>>(zt, t, yc, evals, den_work, alphas, n, d, k, n_alphas,
                                                                  partial);
  GADM_LAUNCHED(h);
  gadm::ridge::ridge_gcv_reduce_kernel<<<(unsigned)((n_alphas * k + 255) / 256), 256, 0, st>>>(partial, n_alphas, tiles, k, n,
                                                                                              score);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_ridge_select(gadm_handle h, const double* score, int64_t n_alphas, int64_t k, int per_target,
                      const double* alphas, const double* evals, const double* t, int64_t d, int32_t* best,
                      double* best_score, double* t_scaled, void* stream) {
  GADM_REQUIRE(h && score && alphas && evals && t && best && best_score && t_scaled && n_alphas > 0 && k > 0 && d > 0,
               "bad argument");
  DeviceGuard guard(h->device);
  gadm::ridge::ridge_select_kernel<<<(unsigned)((k + 127) / 128), 128, 0, as_stream(stream)>>>(
      score, n_alphas, k, per_target, alphas, evals, t, d, best, best_score, t_scaled);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_ridge_intercept(gadm_handle h, const double* coef, const double* xmean, const double* ymean, int64_t d,
                         int64_t k, double* intercept, void* stream) {
  GADM_REQUIRE(h && coef && xmean && ymean && intercept && d > 0 && k > 0, "bad argument");
  DeviceGuard guard(h->device);
  gadm::ridge::ridge_intercept_kernel<<<(unsigned)((k + 127) / 128), 128, 0, as_stream(stream)>>>(coef, xmean, ymean, d,
                                                                                                  k, intercept);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

// ------------------------------------------------------------------ bootstrapped datamodel (datamodel.py:8-37)
static_assert(sizeof(gadm_ridge_system) == sizeof(gadm::dm::System), "ABI struct and kernel struct must agree");

int64_t gadm_datamodel_slot_bytes(int64_t n) { return (n * n + 4 * n) * (int64_t)sizeof(double); }

int gadm_datamodel_ridge_systems(gadm_handle h, const double* g0, const double* y, const int32_t* idx, int64_t n,
                                 const gadm_ridge_system* systems, int64_t n_systems, void* workspace,
                                 int64_t workspace_bytes, double* scores, double* wdual, void* stream) {
  GADM_REQUIRE(h && g0 && y && idx && systems && workspace && scores && wdual && n > 1 && n_systems > 0, "bad argument");
  GADM_REQUIRE(n <= 16384, "n = %lld: one CTA factors an n x n fp64 system in place", (long long)n);
  const int64_t slots = workspace_bytes / gadm_datamodel_slot_bytes(n);
  if (slots < 1)
    return fail(GADM_ERR_WORKSPACE, "workspace %lld B < one slot of %lld B", (long long)workspace_bytes,
                (long long)gadm_datamodel_slot_bytes(n));
  DeviceGuard guard(h->device);
  int64_t ctas = n_systems < slots ? n_systems : slots;
  if (ctas > 2 * h->num_sms) ctas = 2 * h->num_sms;
  gadm::dm::ridge_fold_kernel<<<(unsigned)ctas, gadm::dm::kThreads, 0, as_stream(stream)>>>(
      g0, y, idx, (int)n, reinterpret_cast<const gadm::dm::System*>(systems), (int)n_systems,
      reinterpret_cast<double*>(workspace), scores, wdual);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_shapley_rhs(gadm_handle h, const double* ainv, const double* b, int64_t d, int64_t k, const double* v1,
                     const double* v0, double* colsum_work, double* rhs, void* stream) {
  GADM_REQUIRE(h && ainv && b && v1 && v0 && colsum_work && rhs && d > 0 && k > 0, "bad argument");
  DeviceGuard guard(h->device);
  gadm::agg::shapley_colsum_kernel<<<1, 256, 0, as_stream(stream)>>>(ainv, d, colsum_work);
  GADM_LAUNCHED(h);
  gadm::agg::shapley_rhs_kernel<<<(unsigned)((k + 127) / 128), 128, 0, as_stream(stream)>>>(colsum_work, b, d, k, v1, v0,
                                                                                          rhs);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_lds_spearman(gadm_handle h, const double* pred, const double* y, int64_t m, int64_t k, const int32_t* idx,
                      int64_t n_eval, int64_t rows_per_eval, double* rho, void* stream) {
  GADM_REQUIRE(h && pred && y && rho && m > 0 && k > 0 && n_eval > 0, "bad argument");
  if (!idx) GADM_REQUIRE(n_eval == 1 && rows_per_eval == m, "identity evaluation needs n_eval = 1, rows_per_eval = m");
  GADM_REQUIRE(rows_per_eval > 0 && rows_per_eval <= gadm::agg::kLdsMaxRows, "rows_per_eval %lld out of range (1..%d)",
               (long long)rows_per_eval, gadm::agg::kLdsMaxRows);
  DeviceGuard guard(h->device);
  const bool count = rows_per_eval <= gadm::agg::kLdsCountRows;  // small sets: counting ranks; larger: sorting network
  const size_t per_warp = count ? (size_t)2 * rows_per_eval * sizeof(double) : gadm::agg::lds_warp_smem_bytes(rows_per_eval);
  int warps = 8;
  while (warps > 1 && (size_t)warps * per_warp > 48 * 1024) warps /= 2;
  const size_t smem = (size_t)warps * per_warp;
  if (smem > 48 * 1024)  // one warp per CTA and a large evaluation set: opt in to the large carve-out
    GADM_CUDA(cudaFuncSetAttribute(gadm::agg::lds_spearman_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t jobs = n_eval * k;
  const unsigned grid = (unsigned)((jobs + warps - 1) / warps);
  if (count)
    gadm::agg::lds_spearman_count_kernel<<<grid, warps * 32, smem, as_stream(stream)>>>(pred, y, m, k, idx, n_eval,
                                                                                        rows_per_eval, rho);
  else
    gadm::agg::lds_spearman_kernel<<<grid, warps * 32, smem, as_stream(stream)>>>(pred, y, m, k, idx, n_eval,
                                                                                  rows_per_eval, rho);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_lds_mean(gadm_handle h, const double* rho, int64_t n_eval, int64_t k, double* out, void* stream) {
  GADM_REQUIRE(h && rho && out && n_eval > 0 && k > 0, "bad argument");
  DeviceGuard guard(h->device);
  gadm::agg::lds_mean_kernel<<<(unsigned)((n_eval + 63) / 64), 64, 0, as_stream(stream)>>>(rho, n_eval, k, out);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_group_reduce(gadm_handle h, const void* values, int dtype, const int32_t* group, int64_t n, int64_t n_groups,
                      int mode, double* out, void* stream) {
  GADM_REQUIRE(h && values && group && out && n > 0 && n_groups > 0 && mode >= 0 && mode <= 2, "bad argument");
  DeviceGuard guard(h->device);
  const unsigned blocks = (unsigned)((n_groups + 31) / 32);
  if (dtype == GADM_DTYPE_F32)
    gadm::agg::group_reduce_kernel<float><<<blocks, 32, 0, as_stream(stream)>>>(
        reinterpret_cast<const float*>(values), group, n, n_groups, mode, out);
  else if (dtype == GADM_DTYPE_F64)
    gadm::agg::group_reduce_kernel<double><<<blocks, 32, 0, as_stream(stream)>>>(
        reinterpret_cast<const double*>(values), group, n, n_groups, mode, out);
  else
    return fail(GADM_ERR_INVALID, "group_reduce supports f32 / f64 values, got dtype %d", dtype);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_stable_rank_desc(gadm_handle h, const double* x, int64_t n, int64_t* rank, void* stream) {
  GADM_REQUIRE(h && x && rank && n > 0, "bad argument");
  DeviceGuard guard(h->device);
  gadm::agg::stable_rank_desc_kernel<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(x, n, rank);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

int gadm_row_mean(gadm_handle h, const double* x, int64_t n, int64_t k, double* out, void* stream) {
  GADM_REQUIRE(h && x && out && n > 0 && k > 0, "bad argument");
  DeviceGuard guard(h->device);
  gadm::agg::row_mean_kernel<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(x, n, k, out);
  GADM_LAUNCHED(h);
  return GADM_OK;
}

// ------------------------------------------------------------------ composite entry points

int gadm_project(gadm_handle h, const gadm_block* blocks, int n_blocks, int dtype, int64_t batch, float scale,
                 void* staged, int stage_dtype, float* inv_scale, int64_t d_pad, int64_t m_cap, int64_t proj_dim,
                 uint64_t seed64, int proj_type, float* out, int64_t ld_out, int accumulate, void* workspace,
                 int64_t workspace_bytes, int cta_group, void* stream) {
  GADM_REQUIRE(h && blocks && n_blocks > 0 && staged && out && batch > 0, "bad argument");
  GADM_TRY(gadm_stage_rows(h, blocks, n_blocks, dtype, batch, scale, staged, stage_dtype, d_pad, m_cap, 0, inv_scale, 0, stream));
  return gadm_project_staged(h, staged, stage_dtype, inv_scale, batch, d_pad, m_cap, 0, proj_dim, seed64, proj_type, out,
                             ld_out, accumulate, workspace, workspace_bytes, cta_group, stream);
}

int gadm_gram(gadm_handle h, const float* phi, int64_t n, int64_t k, int64_t ld_phi, float* phi_t_work, int64_t ld_t,
              float* g, int64_t ldg, float diag_add, int accumulate, void* stream) {
  GADM_REQUIRE(h && phi && phi_t_work && g && n > 0 && k > 0 && ld_t >= n && ld_t % 4 == 0, "bad argument");
  GADM_TRY(gadm_transpose(h, phi, n, k, ld_phi, phi_t_work, ld_t, stream));
  return gadm_gemm_tn(h, phi_t_work, ld_t, phi_t_work, ld_t, g, ldg, k, k, n, 1.f, accumulate ? 1.f : 0.f, diag_add, 1, stream);
}

int gadm_score(gadm_handle h, const float* gen, int64_t t, int64_t ld_gen, const float* x, int64_t ldx, const float* xt,
               int64_t ldxt, int64_t k, const float* train, int64_t n_loc, int64_t ld_train, float* z_work, int64_t ldz,
               float* s, int64_t lds, const float* col_scale, float* mean_out, void* stream) {
  GADM_REQUIRE(h && gen && x && xt && train && z_work && s && t > 0 && k > 0 && n_loc > 0 && ldz >= k && ldz % 4 == 0,
               "bad argument");
  float* z1 = z_work;
  float* z2 = z_work + t * ldz;
  GADM_TRY(gadm_gemm_tn(h, gen, ld_gen, x, ldx, z1, ldz, t, k, k, 1.f, 0.f, 0.f, 0, stream));   // gen L^-T
  GADM_TRY(gadm_gemm_tn(h, z1, ldz, xt, ldxt, z2, ldz, t, k, k, 1.f, 0.f, 0.f, 0, stream));     // ... L^-1
  GADM_TRY(gadm_gemm_tn(h, z2, ldz, train, ld_train, s, lds, t, n_loc, k, 1.f, 0.f, 0.f, 0, stream));
  if (mean_out) GADM_TRY(gadm_col_mean_scaled(h, s, t, n_loc, lds, nullptr, col_scale, mean_out, stream));
  return GADM_OK;
}

int64_t gadm_shapley_workspace_bytes(int64_t d, int64_t k) {
  // A [d, d] | Ainv [d, d] | b [d, k] | rhs [d, k] | colsum [d + 1] | info | pinv workspace
  return (2 * d * d + 2 * d * k + d + 1 + 2) * (int64_t)sizeof(double) + gadm_sym_pinv_workspace_bytes(d) + 256;
}

static int mask_regression(gadm_handle h, int mode, const uint32_t* rowbits, const uint32_t* colbits, int64_t n, int64_t d,
                           const double* y, int64_t k, const double* v1, const double* v0, void* workspace,
                           int64_t workspace_bytes, double* phi, void* stream) {
  GADM_REQUIRE(h && rowbits && colbits && y && workspace && phi && n > 0 && d > 0 && k > 0, "bad argument");
  if (workspace_bytes < gadm_shapley_workspace_bytes(d, k))
    return fail(GADM_ERR_WORKSPACE, "workspace %lld B < required %lld B", (long long)workspace_bytes,
                (long long)gadm_shapley_workspace_bytes(d, k));
  double* a = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  double* ainv = a + d * d;
  double* b = ainv + d * d;
  double* rhs = b + d * k;
  double* colsum = rhs + d * k;
  int* info = reinterpret_cast<int*>(colsum + d + 1);
  void* pws = reinterpret_cast<void*>(colsum + d + 3);
  GADM_TRY(gadm_mask_gram(h, colbits, n, d, mode, a, stream));
  if (mode == 0) {  // Shapley: datashapley.py:29-45
    GADM_REQUIRE(v1 && v0, "v1 / v0 missing");
    GADM_TRY(gadm_mask_xty(h, rowbits, y, n, d, k, v0, 0.0, 1.0 / (double)n, b, stream));
    GADM_TRY(gadm_sym_pinv(h, a, d, 1e-15, ainv, pws, gadm_sym_pinv_workspace_bytes(d), info, stream));
    GADM_TRY(gadm_shapley_rhs(h, ainv, b, d, k, v1, v0, colsum, rhs, stream));
    return gadm_dgemm_dk(h, ainv, rhs, d, k, 1e-10, phi, stream);
  }
  // Banzhaf: databanzhaf.py:19-25 (lstsq(rcond=None) cut-off = eps * d)
  GADM_TRY(gadm_mask_xty(h, rowbits, y, n, d, k, nullptr, 0.5, 1.0, b, stream));
  GADM_TRY(gadm_sym_pinv(h, a, d, 2.220446049250313e-16 * (double)d, ainv, pws, gadm_sym_pinv_workspace_bytes(d), info, stream));
  return gadm_dgemm_dk(h, ainv, b, d, k, 0.0, phi, stream);
}

int gadm_shapley(gadm_handle h, const uint32_t* rowbits, const uint32_t* colbits, int64_t n, int64_t d, const double* y,
                 int64_t k, const double* v1, const double* v0, void* workspace, int64_t workspace_bytes, double* phi,
                 void* stream) {
  return mask_regression(h, 0, rowbits, colbits, n, d, y, k, v1, v0, workspace, workspace_bytes, phi, stream);
}

int gadm_banzhaf(gadm_handle h, const uint32_t* rowbits, const uint32_t* colbits, int64_t n, int64_t d, const double* y,
                 int64_t k, void* workspace, int64_t workspace_bytes, double* phi, void* stream) {
  return mask_regression(h, 1, rowbits, colbits, n, d, y, k, nullptr, nullptr, workspace, workspace_bytes, phi, stream);
}

int gadm_lds(gadm_handle h, const uint32_t* test_colbits, int64_t m, int64_t d, const double* y_test, const double* phi,
             int64_t k, const int32_t* idx, int64_t n_eval, int64_t rows_per_eval, void* workspace,
             int64_t workspace_bytes, double* lds_out, void* stream) {
  GADM_REQUIRE(h && test_colbits && y_test && phi && workspace && lds_out && m > 0 && d > 0 && k > 0 && n_eval > 0,
               "bad argument");
  if (workspace_bytes < (m + n_eval) * k * (int64_t)sizeof(double))
    return fail(GADM_ERR_WORKSPACE, "workspace %lld B < required %lld B", (long long)workspace_bytes,
                (long long)((m + n_eval) * k * (int64_t)sizeof(double)));
  double* pred = reinterpret_cast<double*>(workspace);
  double* rho = pred + m * k;
  GADM_TRY(gadm_mask_times_matrix(h, test_colbits, phi, m, d, k, pred, stream));
  GADM_TRY(gadm_lds_spearman(h, pred, y_test, m, k, idx, n_eval, rows_per_eval, rho, stream));
  return gadm_lds_mean(h, rho, n_eval, k, lds_out, stream);
}

}  // extern "C"
