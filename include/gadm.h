/* gadm -- C ABI of the B200-native TRAK / Shapley attribution hot path.
 *
 * Every entry point takes plain pointers and sizes; device pointers are CUDA device addresses owned by
 * the caller (PyTorch in practice: tensor.data_ptr()), `stream` is a cudaStream_t passed as void*
 * (torch.cuda.current_stream().cuda_stream).  Calls are asynchronous on that stream, never allocate
 * user-visible memory and never retain pointers after returning.  Return value: 0 on success, a
 * negative code on failure with the message available from gadm_last_error() (thread-local).
 * A handle belongs to one (process, device); it is not thread-safe, distinct handles are.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   gadm_pack_block / gadm_project_staged / gadm_materialize_p
 *       trak.projectors.CudaProjector(...).project(grads, model_id)  -> fast_jl.project_*  (third-party,
 *       requirements.txt:10-11); call sites src/attributions/methods/d_trak_grad.py:504-511,776 and
 *       text_to_image/grad_text_to_image_lora.py:561-568,765,813; vectorize_and_ignore_buffers
 *       (d_trak_grad.py:188-226) is subsumed by per-block packing.
 *   gadm_gemm_tn / gadm_gram / gadm_cholesky / gadm_solve_rows / gadm_transpose / gadm_row_norms /
 *   gadm_col_mean_scaled
 *       text_to_image/traks.py:141-186 (torch.matmul / torch.inverse / norms / mean) and
 *       src/attributions/methods/compute_gradient_score.py:75-79,108-130.
 *   gadm_group_reduce / gadm_stable_rank_desc
 *       text_to_image/traks.py:188-225, src/attributions/methods/attribution_utils.py:15-48,
 *       text_to_image/shapley_lds.py:294.
 *   gadm_pack_masks / gadm_mask_gram / gadm_mask_xty / gadm_sym_pinv / gadm_dgemm / gadm_shapley_finish
 *       src/attributions/methods/datashapley.py:8-48, src/attributions/methods/databanzhaf.py:5-26.
 *   gadm_lds_spearman
 *       evaluate_lds: text_to_image/shapley_lds.py:138-150, lds.py:158-170 and bootstrap statistic
 *       lds.py:458-485.
 */
#ifndef GADM_H_
#define GADM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gadm_ctx* gadm_handle;

enum { GADM_PROJ_NORMAL = 0, GADM_PROJ_RADEMACHER = 1 };
enum { GADM_DTYPE_F32 = 0, GADM_DTYPE_BF16 = 1, GADM_DTYPE_F16 = 2 };
enum {
  GADM_OK = 0,
  GADM_ERR_INVALID = -1,   /* bad argument (shape, alignment, enum) */
  GADM_ERR_CUDA = -2,      /* CUDA runtime / driver error, incl. kernel watchdog */
  GADM_ERR_WORKSPACE = -3, /* workspace too small */
  GADM_ERR_UNSUPPORTED = -4 /* device is not sm_100 */
};

int gadm_version(void);
const char* gadm_last_error(void);
int gadm_create(gadm_handle* out, int device);
int gadm_destroy(gadm_handle h);
/* number of kernels this handle has launched since creation (for bench.py's gpu_launches) */
int64_t gadm_launch_count(gadm_handle h);
/* reads and clears the in-kernel watchdog word (0 = no barrier timeout fired) */
int gadm_watchdog_code(gadm_handle h, unsigned int* code);

/* ---------------------------------------------------------------- JL projection */

/* Bytes of scratch gadm_project_staged needs for split-K partial tiles. */
int64_t gadm_project_workspace_bytes(gadm_handle h, int64_t m_rows, int64_t d_pad, int64_t proj_dim, int cta_group);

/* Convert one gradient block to bf16 inside the staging buffer.
 *   src: [batch, numel] of `dtype`, consecutive examples `src_stride` elements apart
 *   staged: bf16 [rows, ld] row-major; block lands at rows row0..row0+batch, columns col0..col0+numel
 *   the value written is bf16(src * scale)  (scale = 1/K folds the timestep mean, d_trak_grad.py:770) */
int gadm_pack_block(gadm_handle h, const void* src, int dtype, int64_t batch, int64_t numel, int64_t src_stride,
                    void* staged, int64_t ld, int64_t row0, int64_t col0, float scale, void* stream);

/* out[m, :] (+)= staged[m, :] * P[p_base : p_base + d_pad, 0:proj_dim]
 *   staged: bf16 [m_rows <= 512, d_pad] row-major with pitch ld (elements); d_pad % 64 == 0, ld % 8 == 0,
 *           16-byte aligned base; columns beyond the real gradient length must be zero
 *   p_base: canonical index (row of P) of staged column 0, multiple of 64
 *   proj_dim % 256 == 0; seed64 = seed + 10^4 * model_id (CudaProjector semantics)
 *   out: fp32 [m_rows, proj_dim] with pitch ld_out; accumulate != 0 adds to out (D-chunked projection)
 *   cta_group: 2 (CTA-pair UMMA, default) or 1 */
int gadm_project_staged(gadm_handle h, const void* staged, int64_t m_rows, int64_t d_pad, int64_t ld, int64_t p_base,
                        int64_t proj_dim, uint64_t seed64, int proj_type, float* out, int64_t ld_out, int accumulate,
                        void* workspace, int64_t workspace_bytes, int cta_group, void* stream);

/* out[r, j] = P[row0 + r, j] as fp32, r < nrows, j < proj_dim (oracle hook: the kernel's own matrix) */
int gadm_materialize_p(gadm_handle h, int64_t row0, int64_t nrows, int64_t proj_dim, uint64_t seed64, int proj_type,
                       float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GADM_H_ */
