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
 *   gadm_stage_rows / gadm_pack_block / gadm_project_staged / gadm_project / gadm_materialize_p
 *       trak.projectors.CudaProjector(...).project(grads, model_id)  -> fast_jl.project_*  (third-party,
 *       requirements.txt:10-11); call sites src/attributions/methods/d_trak_grad.py:504-511,776 and
 *       text_to_image/grad_text_to_image_lora.py:561-568,765,813; vectorize_and_ignore_buffers
 *       (d_trak_grad.py:188-226) is subsumed by the block table of gadm_stage_rows.
 *   gadm_accumulate_rows
 *       the timestep sum and mean of the featurisation loop: emb += grads, emb / K (d_trak_grad.py:764-770,
 *       grad_text_to_image_lora.py:808-812).
 *   gadm_gemm_tn / gadm_gram / gadm_cholesky / gadm_tri_inverse / gadm_solve_rows / gadm_score /
 *   gadm_transpose / gadm_row_norms / gadm_col_mean_scaled
 *       text_to_image/traks.py:141-186 (torch.matmul / torch.inverse / norms / mean) and
 *       src/attributions/methods/compute_gradient_score.py:75-79,108-130.
 *   gadm_group_reduce / gadm_stable_rank_desc
 *       text_to_image/traks.py:188-225, src/attributions/methods/attribution_utils.py:15-48,
 *       text_to_image/shapley_lds.py:294.
 *   gadm_shapley / gadm_banzhaf (= gadm_pack_masks, gadm_mask_gram, gadm_mask_xty, gadm_sym_pinv, gadm_shapley_rhs,
 *   gadm_dgemm_dk in sequence)
 *       src/attributions/methods/datashapley.py:8-48, src/attributions/methods/databanzhaf.py:5-26.
 *   gadm_center_columns / gadm_dgemm / gadm_sym_eig / gadm_ridge_gcv / gadm_ridge_select / gadm_ridge_intercept
 *       RidgeCV datamodel estimator, lds.py:411-421;  gadm_datamodel_ridge_systems: datamodel.py:8-37.
 *   gadm_lds (= gadm_mask_times_matrix, gadm_lds_spearman, gadm_lds_mean)
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
/* 16-bit formats of the staged gradients (the A operand of the projection GEMM):
 *   GADM_STAGE_BF16  bf16(v): 8 significant bits, no scaling (P generated as bf16).
 *   GADM_STAGE_F16G  fp16(v * 2^s), one power-of-two scale per (example row, group of GADM_STAGE_GROUP_COLS columns)
 *                    chosen so that the group's largest magnitude lands in [2^8, 65504) (guessed from a sample of the
 *                    group, exact when the guess misses; csrc/stage.cuh): 11 significant bits and no fp16 range
 *                    problem (P generated as fp16).  The inverse scales live in a caller-owned fp32 array
 *                    inv_scale[m_cap][gadm_stage_scale_count(d_pad)] written by gadm_stage_rows and read by
 *                    gadm_project_staged (multiplied in, exactly, when accumulation segments are promoted). */
enum { GADM_STAGE_BF16 = 0, GADM_STAGE_F16G = 1 };
#define GADM_STAGE_GROUP_COLS 32768
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
/* bound (ns of wall time) on any in-kernel barrier wait before the kernel traps; 0 disables (profilers) */
int gadm_set_watchdog_ns(gadm_handle h, uint64_t ns);

/* ---------------------------------------------------------------- JL projection */

/* Bytes of scratch gadm_project_staged needs for split-K partial tiles. */
int64_t gadm_project_workspace_bytes(gadm_handle h, int64_t m_rows, int64_t d_pad, int64_t proj_dim, int cta_group);

/* Staging buffer layout (16-bit elements): staged[kb][row][c] with kb = p / 64, c = p % 64, i.e. a contiguous
 * [d_pad / 64][m_cap][64] array, 128-byte aligned.  p is the position in the flattened gradient, row the example.
 * Tile-major so that the 128-row x 64-column tiles the kernel streams are contiguous in HBM. */

/* One parameter block of a batch of per-example gradients (what vmap(grad(f)) returns per parameter):
 *   ptr                 device pointer to [batch, numel_per_example] elements of `dtype`
 *   example_stride      elements between consecutive examples
 *   row_offset          position of the block's first element in the flattened gradient (= row of P) */
typedef struct {
  const void* ptr;
  int64_t numel_per_example;
  int64_t example_stride;
  int64_t row_offset;
} gadm_block;

/* scale groups per staged row = ceil(d_pad / GADM_STAGE_GROUP_COLS) */
int64_t gadm_stage_scale_count(int64_t d_pad);

/* Stage a batch of examples in ONE launch: rows row0 .. row0 + batch of the staging buffer receive
 * convert(src * scale) for every column; `blocks` (host array, sorted by row_offset, non-overlapping, at most 1024)
 * lists the parameter blocks, columns no block covers (gaps, the tail up to d_pad) are written as zeros.
 * scale = 1/K folds the timestep mean (d_trak_grad.py:770).  inv_scale: see GADM_STAGE_F16G (NULL for bf16).
 * coresident selects the CTA shape, results are identical: 0 = wide CTAs (the whole GPU when nothing else runs; while a
 * 4-CTA-cluster projection is running on another stream they are confined to the 16 SMs that grid cannot use and do
 * not disturb it), 1 = narrow CTAs sized to run beside a persistent projection CTA on the same SM (for the CTA-pair
 * projection kernels, whose grid covers every SM). */
int gadm_stage_rows(gadm_handle h, const gadm_block* blocks, int n_blocks, int dtype, int64_t batch, float scale,
                    void* staged, int stage_dtype, int64_t d_pad, int64_t m_cap, int64_t row0, float* inv_scale,
                    int coresident, void* stream);


/* Timestep accumulator: slab[row0 + b, p] = (accumulate ? slab[row0 + b, p] : 0) + scale * src  for an fp32 slab
 * [slab_rows][d_pad] (32-byte aligned).  Sum K timesteps with scale = 1/K, then stage the slab rows with
 * gadm_stage_rows (one fp32 block {slab, d_pad, d_pad, 0}).  Replaces emb += grads / emb / K of d_trak_grad.py:764-770
 * without the flatten / cat of vectorize_and_ignore_buffers. */
int gadm_accumulate_rows(gadm_handle h, const gadm_block* blocks, int n_blocks, int dtype, int64_t batch, float scale,
                         float* slab, int64_t d_pad, int64_t slab_rows, int64_t row0, int accumulate, void* stream);

/* Convert one gradient block to bf16 inside a GADM_STAGE_BF16 staging buffer (per-block form of gadm_stage_rows).
 *   src: [batch, numel] of `dtype`, consecutive examples `src_stride` elements apart
 *   block lands at rows row0..row0+batch, positions col0..col0+numel of the flattened gradient
 *   the value written is bf16(src * scale); positions no call writes must have been zeroed by the caller */
int gadm_pack_block(gadm_handle h, const void* src, int dtype, int64_t batch, int64_t numel, int64_t src_stride,
                    void* staged, int64_t d_pad, int64_t m_cap, int64_t row0, int64_t col0, float scale, void* stream);

/* out[m, :] (+)= G[m, :] * P[p_base : p_base + d_pad, 0:proj_dim]  for the first m_rows rows of the staging buffer
 *   staged: layout above, format stage_dtype (+ inv_scale for GADM_STAGE_F16G); d_pad % 64 == 0; positions beyond the
 *           real gradient length must be zero; rows >= m_rows may hold anything (they only feed output rows that are
 *           never written)
 *   m_rows <= 512 (256 for cta_group 1, 1024 for cta_group 4), m_cap >= m_rows
 *   p_base: canonical index (row of P) of position 0, multiple of 64
 *   proj_dim % 256 == 0; seed64 = seed + 10^4 * model_id (CudaProjector semantics)
 *   out: fp32 [m_rows, proj_dim] with pitch ld_out; accumulate != 0 adds to out (D-chunked projection)
 *   cta_group: 2 (CTA-pair UMMA, default), 1 (single CTA), or 4 = two CTA pairs per cluster that share the
 *              generated P tiles through DSMEM bulk copies (halves the generator work; up to 1024 rows)
 *   Launches in flight on different streams of one device are independent (each takes its own lockstep counter);
 *   the handle itself is still not thread-safe. */
int gadm_project_staged(gadm_handle h, const void* staged, int stage_dtype, const float* inv_scale, int64_t m_rows,
                        int64_t d_pad, int64_t m_cap, int64_t p_base, int64_t proj_dim, uint64_t seed64, int proj_type,
                        float* out, int64_t ld_out, int accumulate, void* workspace, int64_t workspace_bytes,
                        int cta_group, void* stream);

/* out[r, j] = P[row0 + r, j] as fp32, r < nrows, j < proj_dim (oracle hook: the kernel's own matrix).  P is
 * generated in the 16-bit format of the staged gradients (tcgen05 kind::f16 multiplies f16 x f16 or bf16 x bf16):
 * +-1 is the same matrix in both, the normal type is Box-Muller rounded to bf16 or to fp16 (stage_dtype). */
int gadm_materialize_p(gadm_handle h, int64_t row0, int64_t nrows, int64_t proj_dim, uint64_t seed64, int proj_type,
                       int stage_dtype, float* out, void* stream);

/* ---------------------------------------------------------------- TRAK scorer (fp32 data, 3xTF32 tensor-core GEMM) */

/* C[M, N] = alpha * A[M, K] * B[N, K]^T + beta * C, then C[i, i] += diag_add.
 *   A, B, C fp32 row-major (contraction index contiguous in A and B); lda, ldb multiples of 4; A, B 16-byte
 *   aligned.  lower_only is a flag word: bit 0 skips 128x128 tiles strictly above the diagonal (symmetric results);
 *   bit 1 (value 2) declares B lower-triangular (B[j, c] = 0 for c > j), bit 2 (value 4) upper-triangular
 *   (B[j, c] = 0 for c < j): the contraction of every output tile then stops / starts at its diagonal block, which
 *   halves the work of the two solve GEMMs against L^-1 and L^-T.
 *   Replaces torch.matmul at text_to_image/traks.py:141,149,152,156,171,176,181,184 and the numpy products at
 *   src/attributions/methods/compute_gradient_score.py:75,79,108,126 (fp32-grade accuracy via 3xTF32). */
int gadm_gemm_tn(gadm_handle h, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc,
                 int64_t m, int64_t n, int64_t k, float alpha, float beta, float diag_add, int lower_only,
                 void* stream);
/* The same product for `batch` independent problems whose operands / results lie stride_a / stride_b / stride_c
 * elements apart (blockIdx.z; 3-D tensor maps).  Used by gadm_tri_inverse: the merges of one recursion level are
 * one launch. */
int gadm_gemm_tn_batched(gadm_handle h, const float* a, int64_t lda, int64_t stride_a, const float* b, int64_t ldb,
                         int64_t stride_b, float* c, int64_t ldc, int64_t stride_c, int64_t m, int64_t n, int64_t k,
                         int64_t batch, float alpha, float beta, float diag_add, int lower_only, void* stream);
/* out[c, r] = in[r, c]  (in: [rows, cols] pitch ld_in; out: [cols, rows] pitch ld_out) */
int gadm_transpose(gadm_handle h, const float* in, int64_t rows, int64_t cols, int64_t ld_in, float* out,
                   int64_t ld_out, void* stream);
/* In-place lower Cholesky factor of the symmetric positive definite a [k, k] (pitch ld, ld % 4 == 0, lower
 * triangle read, strict upper of each diagonal block zeroed).  blocks: workspace of gadm_cholesky_workspace_bytes(k)
 * that receives the inverses of the 128x128 diagonal factor blocks (used by gadm_solve_rows).
 * info: device int, 0 on success else 1 + index of the first non-positive pivot.
 * Replaces torch.inverse / np.linalg.inv at traks.py:151,180 and compute_gradient_score.py:77,110 (the
 * inverse is never formed unless asked for: gadm_solve_rows applies it). */
int64_t gadm_cholesky_workspace_bytes(int64_t k);
int gadm_cholesky(gadm_handle h, float* a, int64_t ld, int64_t k, void* blocks, int64_t blocks_bytes, int* info,
                  void* stream);
/* x = (L L^T)^-1 b for ONE right-hand side b [k] through the factor of gadm_cholesky and its `blocks` workspace: forward
 * and backward substitution over the 128-blocks in one cooperative launch (k / 128 CTAs, point-to-point signalling).
 * What the mean-first TRAK score needs (traks.py:152-157: mean_t(gen_t) K^-1 Phi^T) without forming the triangular
 * inverse.  Requires k <= 8192, ld % 4 == 0, l 16-byte aligned (k need not be a multiple of 128); workspace of
 * gadm_cholesky_solve_vec_workspace_bytes(k), 16-byte aligned. */
int64_t gadm_cholesky_solve_vec_workspace_bytes(int64_t k);
int gadm_cholesky_solve_vec(gadm_handle h, const float* l, int64_t ld, const void* blocks, int64_t k, const float* b,
                            float* x, void* workspace, int64_t workspace_bytes, void* stream);
/* x = L^-1 (lower) and xt = L^-T (upper), both [k, k], of the factor left by gadm_cholesky (recursive doubling over
 * the 128-wide diagonal blocks whose inverses are in `blocks`).  rows @ K^-1 is then two GEMMs:
 * gadm_gemm_tn(rows, x) = rows L^-T, gadm_gemm_tn(., xt) = (.) L^-1 -- what torch.inverse + matmul do in
 * traks.py:151-154, without ever forming K^-1. */
int64_t gadm_tri_inverse_workspace_bytes(int64_t k);
int gadm_tri_inverse(gadm_handle h, const float* l, int64_t ldl, const void* blocks, int64_t k, float* x, int64_t ldx,
                     float* xt, int64_t ldxt, void* workspace, int64_t workspace_bytes, void* stream);
/* y[m, k] <- y * (L L^T)^-1 in place (every row of y is a right-hand side).  l: Cholesky factor from
 * gadm_cholesky, u: its transpose (gadm_transpose), blocks: the same workspace. */
int gadm_solve_rows(gadm_handle h, const float* l, int64_t ldl, const float* u, int64_t ldu, const void* blocks,
                    int64_t k, float* y, int64_t ldy, int64_t m, void* stream);
/* out[r] = col_scale[r] * sum_j x[r, j] * v[j]  (col_scale may be NULL).  The mean over generated images of
 * traks.py:157,162-168 is linear in the generated features: mean_t(gen_t K^-1 phi_n) = (mean_t gen_t) K^-1 phi_n,
 * i.e. one solved row and this product instead of the [T, N] score GEMM. */
int gadm_matvec_rows(gadm_handle h, const float* x, int64_t rows, int64_t cols, int64_t ld, const float* v,
                     const float* col_scale, float* out, void* stream);
/* out2[0] = min_i L[i, i], out2[1] = max_i L[i, i] of the factor left by gadm_cholesky ((max / min)^2 bounds cond(K)
 * from below: the host refuses fp32 results when it exceeds what fp32 can resolve; NaN pivots give NaN) */
int gadm_diag_minmax(gadm_handle h, const float* l, int64_t ld, int64_t k, float* out2, void* stream);
/* out[r] = ||x[r, :]||_2, or 1 / that when reciprocal != 0 (traks.py:143,162,166) */
int gadm_row_norms(gadm_handle h, const float* x, int64_t rows, int64_t cols, int64_t ld, int reciprocal, float* out,
                   void* stream);
/* out[n] = mean_t s[t, n] * row_scale[t] * col_scale[n]; either scale may be NULL (traks.py:146,157,162-168) */
int gadm_col_mean_scaled(gadm_handle h, const float* s, int64_t t, int64_t n, int64_t ld, const float* row_scale,
                         const float* col_scale, float* out, void* stream);
/* s[t, n] *= row_scale[t] * col_scale[n] in place (compute_gradient_score.py:114-126) */
int gadm_scale_rows_cols(gadm_handle h, float* s, int64_t t, int64_t n, int64_t ld, const float* row_scale,
                         const float* col_scale, void* stream);

/* ---------------------------------------------------------------- subset-mask aggregation (fp64) */

enum { GADM_GRAM_SHAPLEY = 0, GADM_GRAM_BANZHAF = 1 };
enum { GADM_GROUP_SUM = 0, GADM_GROUP_MEAN = 1, GADM_GROUP_MAX = 2 };
enum { GADM_DTYPE_F64 = 3 };

/* X: uint8 [n, d] (0/1 subset masks as built at shapley_lds.py:114-119) -> bit-packed
 * rowbits uint32 [n, ceil(d/32)] and colbits uint32 [d, ceil(n/32)] */
int gadm_pack_masks(gadm_handle h, const uint8_t* x, int64_t n, int64_t d, uint32_t* rowbits, uint32_t* colbits,
                    void* stream);
/* A [d, d] fp64: mode SHAPLEY  A = X^T X / n (datashapley.py:29);
 *                mode BANZHAF  A = (X - 1/2)^T (X - 1/2) (databanzhaf.py:19-22) -- exact from popcounts;
 *                mode 2        raw co-occurrence counts (pass the row bit planes with n and d swapped for X X^T) */
int gadm_mask_gram(gadm_handle h, const uint32_t* colbits, int64_t n, int64_t d, int mode, double* a, void* stream);
/* out[i, k] = (sum_r X[r,i] * (Y[r,k] - shift[k]) - half * sum_r (Y[r,k] - shift[k])) * scale ; Y [n, K], out [d, K]
 * Shapley b_hat: shift = v0, half = 0, scale = 1/n ; Banzhaf rhs: shift = NULL, half = 0.5, scale = 1 */
int gadm_mask_xty(gadm_handle h, const uint32_t* rowbits, const double* y, int64_t n, int64_t d, int64_t k,
                  const double* shift, double half, double scale, double* out, void* stream);
/* out[r, k] = sum_i X[r,i] * M[i,k]  (x_test @ attrs_all, shapley_lds.py:145) ; X given by its column bit planes
 * colbits [d, ceil(m/32)], M [d, K], out [m, K]; each row of M is streamed once */
int gadm_mask_times_matrix(gadm_handle h, const uint32_t* colbits, const double* mat, int64_t m, int64_t d, int64_t k,
                           double* out, void* stream);
/* Moore-Penrose inverse of a symmetric matrix by one-sided Jacobi SVD; singular values <= rcond * max are
 * dropped (numpy.linalg.pinv / lstsq cut-off semantics).  info (device int[2], may be NULL) = {sweeps, rank}. */
int64_t gadm_sym_pinv_workspace_bytes(int64_t d);
int gadm_sym_pinv(gadm_handle h, const double* a, int64_t d, double rcond, double* out, void* workspace,
                  int64_t workspace_bytes, int* info, void* stream);
/* C[i, k] = sum_j A[i, j] * B[j, k]; |C| < zero_below -> 0 (datashapley.py:45).  A [d, d], B and C [d, K]. */
int gadm_dgemm_dk(gadm_handle h, const double* a, const double* b, int64_t d, int64_t k, double zero_below, double* c,
                  void* stream);
/* ---- RidgeCV datamodel estimator: replaces `RidgeCV(alphas=np.linspace(0.01, 10, 100)).fit(masks, targets[:, i])`
 * per behaviour (lds.py:411-421; sklearn _RidgeGCV: efficient leave-one-out ridge with intercept), batched over all
 * behaviours and alphas.  Sequence: center_columns(X), center_columns(Y) -> dgemm C = Xc^T Xc -> sym_eig ->
 * dgemm Z = Xc V -> dgemm T = Z^T Yc -> ridge_gcv (scores[a, k]) -> ridge_select -> dgemm coef = V Ts ->
 * ridge_intercept.  All fp64, deterministic summation orders. */
/* xc[i, j] = x[i, j] - mean[j], mean[j] = mean_i x[i, j]; x, xc: [n, d] */
int gadm_center_columns(gadm_handle h, const double* x, int64_t n, int64_t d, double* xc, double* mean, void* stream);
/* c[m, n] = op(a) b; op(a)(i, j) = a[i * lda + j] (trans_a = 0) or a[j * lda + i] (trans_a = 1); b: [j, n] */
int gadm_dgemm(gadm_handle h, int trans_a, const double* a, int64_t lda, const double* b, int64_t ldb, int64_t m,
               int64_t j, int64_t n, double* c, int64_t ldc, void* stream);
/* a = v diag(evals) v^T for symmetric a [d, d] (one-sided Jacobi, fp64); column i of v = eigenvector i, unsorted.
 * info (device int[1], may be NULL) = sweeps. */
int64_t gadm_sym_eig_workspace_bytes(int64_t d);
int gadm_sym_eig(gadm_handle h, const double* a, int64_t d, double* evals, double* v, void* workspace,
                 int64_t workspace_bytes, int* info, void* stream);
/* score[a, k] = -mean_i looe(a)[i, k]^2 (sklearn _RidgeGCV with intercept).  z = Xc V [n, d], t = z^T yc [d, k],
 * yc [n, k] centred targets; workspace of gadm_ridge_gcv_workspace_bytes(n, d, k, n_alphas) bytes. */
int64_t gadm_ridge_gcv_workspace_bytes(int64_t n, int64_t d, int64_t k, int64_t n_alphas);
int gadm_ridge_gcv(gadm_handle h, const double* z, const double* t, const double* yc, const double* evals,
                   const double* alphas, int64_t n, int64_t d, int64_t k, int64_t n_alphas, void* workspace,
                   int64_t workspace_bytes, double* score, void* stream);
/* best[k] = first alpha index with the largest score (per behaviour, or of the mean over behaviours when
 * per_target = 0); t_scaled[j, k] = t[j, k] / (evals[j] + alphas[best[k]]) (coef = v t_scaled). */
int gadm_ridge_select(gadm_handle h, const double* score, int64_t n_alphas, int64_t k, int per_target,
                      const double* alphas, const double* evals, const double* t, int64_t d, int32_t* best,
                      double* best_score, double* t_scaled, void* stream);
/* intercept[k] = ymean[k] - xmean . coef[:, k] */
int gadm_ridge_intercept(gadm_handle h, const double* coef, const double* xmean, const double* ymean, int64_t d,
                         int64_t k, double* intercept, void* stream);
/* ---- bootstrapped datamodel: replaces `RidgeCV(cv=5, alphas=[0.1, 1.0, 10.0]).fit(x[idx], y[idx])` per resample
 * (src/attributions/methods/datamodel.py:8-37; sklearn GridSearchCV over Ridge(fit_intercept=True), KFold(5), R^2).
 * Every ridge fit runs in Gram space on g0 = X X^T of the original rows (gadm_mask_gram mode 2 on the row bit planes):
 * one CTA per system = (resample, held-out positions [f0, f1) or f0 == f1 for the refit, alpha).  Fold systems write
 * the held-out R^2 to scores[out]; refit systems write the centred dual weights w to wdual[out, :] so that
 * coef = X^T w (gadm_mask_xty).  idx: int32 [runs, n] resample indices.  workspace: any number of slots of
 * gadm_datamodel_slot_bytes(n); systems beyond the slot count are processed in turn. */
typedef struct gadm_ridge_system {
  int32_t run, f0, f1, out;
  double alpha;
} gadm_ridge_system;
int64_t gadm_datamodel_slot_bytes(int64_t n);
int gadm_datamodel_ridge_systems(gadm_handle h, const double* g0, const double* y, const int32_t* idx, int64_t n,
                                 const gadm_ridge_system* systems, int64_t n_systems, void* workspace,
                                 int64_t workspace_bytes, double* scores, double* wdual, void* stream);
/* Efficiency-constraint step of closed-form KernelSHAP (datashapley.py:38-43):
 * rhs[:, k] = b[:, k] - (1^T Ainv b[:, k] - v1[k] + v0[k]) / (1^T Ainv 1) ; colsum_work: d + 1 doubles */
int gadm_shapley_rhs(gadm_handle h, const double* ainv, const double* b, int64_t d, int64_t k, const double* v1,
                     const double* v0, double* colsum_work, double* rhs, void* stream);
/* rho[e, k] = Spearman(pred[idx[e, :], k], y[idx[e, :], k]) with average ranks (scipy.stats.spearmanr);
 * pred, y: [m, K]; idx: int32 [n_eval, rows_per_eval] or NULL (identity, n_eval = 1, rows_per_eval = m) */
int gadm_lds_spearman(gadm_handle h, const double* pred, const double* y, int64_t m, int64_t k, const int32_t* idx,
                      int64_t n_eval, int64_t rows_per_eval, double* rho, void* stream);
/* out[e] = 100 * mean_k rho[e, k] */
int gadm_lds_mean(gadm_handle h, const double* rho, int64_t n_eval, int64_t k, double* out, void* stream);
/* out[g] = sum | mean | max of values[i] with group[i] == g, accumulated in fp64 (traks.py:188-204) */
int gadm_group_reduce(gadm_handle h, const void* values, int dtype, const int32_t* group, int64_t n, int64_t n_groups,
                      int mode, double* out, void* stream);
/* rank = argsort(-x, kind="stable") as int64 (traks.py:218, shapley_lds.py:294) */
int gadm_stable_rank_desc(gadm_handle h, const double* x, int64_t n, int64_t* rank, void* stream);
/* out[i] = mean_k x[i, k] */
int gadm_row_mean(gadm_handle h, const double* x, int64_t n, int64_t k, double* out, void* stream);

/* ------------------------------------------------------------------ composite entry points
 * One call per reference function; each is the documented sequence of the calls above on
 * the same stream, with caller-owned workspace.  The Python host layer may call either level. */

/* CudaProjector.project on per-parameter blocks: stage every block (x scale) into `staged` rows 0..batch
 * (gadm_stage_rows), then project them (gadm_project_staged); arguments as for those two. */
int gadm_project(gadm_handle h, const gadm_block* blocks, int n_blocks, int dtype, int64_t batch, float scale,
                 void* staged, int stage_dtype, float* inv_scale, int64_t d_pad, int64_t m_cap, int64_t proj_dim,
                 uint64_t seed64, int proj_type, float* out, int64_t ld_out, int accumulate, void* workspace,
                 int64_t workspace_bytes, int cta_group, void* stream);
/* G (+)= Phi^T Phi (+ diag_add on the diagonal), lower tiles only (traks.py:149-150).  phi [n, k] pitch ld_phi;
 * phi_t_work: [k, ld_t] scratch with ld_t >= n, ld_t % 4 == 0; g [k, k] pitch ldg. */
int gadm_gram(gadm_handle h, const float* phi, int64_t n, int64_t k, int64_t ld_phi, float* phi_t_work, int64_t ld_t,
              float* g, int64_t ldg, float diag_add, int accumulate, void* stream);
/* S = gen K^-1 train^T [t, n_loc] with K^-1 = L^-T L^-1 from gadm_tri_inverse (traks.py:152-156), and optionally
 * mean_out[j] = mean_t S[t, j] * col_scale[j] (traks.py:157,162-168; col_scale may be NULL).
 * z_work: 2 * t * ldz floats (ldz >= k, ldz % 4 == 0). */
int gadm_score(gadm_handle h, const float* gen, int64_t t, int64_t ld_gen, const float* x, int64_t ldx, const float* xt,
               int64_t ldxt, int64_t k, const float* train, int64_t n_loc, int64_t ld_train, float* z_work, int64_t ldz,
               float* s, int64_t lds, const float* col_scale, float* mean_out, void* stream);
/* data_shapley for all behaviours (datashapley.py:8-48): phi [d, k].  workspace >= gadm_shapley_workspace_bytes(d, k). */
int64_t gadm_shapley_workspace_bytes(int64_t d, int64_t k);
int gadm_shapley(gadm_handle h, const uint32_t* rowbits, const uint32_t* colbits, int64_t n, int64_t d, const double* y,
                 int64_t k, const double* v1, const double* v0, void* workspace, int64_t workspace_bytes, double* phi,
                 void* stream);
/* data_banzhaf for all behaviours (databanzhaf.py:5-26): phi [d, k]; same workspace size. */
int gadm_banzhaf(gadm_handle h, const uint32_t* rowbits, const uint32_t* colbits, int64_t n, int64_t d, const double* y,
                 int64_t k, void* workspace, int64_t workspace_bytes, double* phi, void* stream);
/* LDS of one test set (evaluate_lds inner loop / lds.py:my_lds): lds_out[e] = 100 * mean_k Spearman(X_test[idx[e]] phi_k,
 * y_test[idx[e], k]).  test_colbits: column bit planes of the test masks [m, d]; idx NULL -> one identity evaluation.
 * workspace >= (m + n_eval) * k doubles. */
int gadm_lds(gadm_handle h, const uint32_t* test_colbits, int64_t m, int64_t d, const double* y_test, const double* phi,
             int64_t k, const int32_t* idx, int64_t n_eval, int64_t rows_per_eval, void* workspace,
             int64_t workspace_bytes, double* lds_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GADM_H_ */
