/* The drop-in boundary from plain C: project a batch of per-parameter gradient blocks with libgadm.so.
 *
 *   gcc -std=c99 -I include examples/c_abi_projection.c -o /tmp/c_abi_projection \
 *       -L group-attribution-for-diffusion-models_b200/csrc -lgadm -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,$PWD/group-attribution-for-diffusion-models_b200/csrc
 *
 * What a non-Python host does around CudaProjector.project (reference call sites
 * src/attributions/methods/d_trak_grad.py:504-511,776): allocate the 16-bit staging buffer (+ its group scales) and the
 * split-K workspace once, describe the gradient blocks of a batch, call gadm_project, read back [batch, proj_dim] floats.
 * Needs a B200 at run time (gadm_create refuses other devices); tests/test_abi_cpu.py only compiles the header. */
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "gadm.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc_ = (call);                                                      \
    if (rc_ != 0) {                                                        \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, gadm_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(void) {
  const int64_t batch = 8, proj_dim = 2048;
  const int64_t numel[3] = {1728, 64, 36864}; /* e.g. conv weight, bias, next conv weight */
  int64_t grad_dim = 0, d_pad, m_cap = 32, i;
  gadm_handle h;
  gadm_block blocks[3];
  void *staged = NULL, *workspace = NULL;
  float *out = NULL, *inv_scale = NULL, *grads[3];
  int64_t ws_bytes;

  CHECK(gadm_create(&h, 0));
  for (i = 0; i < 3; ++i) grad_dim += numel[i];
  d_pad = (grad_dim + 63) / 64 * 64;
  /* staging buffer [d_pad / 64][m_cap][64] fp16 and one inverse scale per (row, 32768-column group); gadm_stage_rows
   * writes every column of the rows it stages (padding included), so neither needs initialising */
  if (cudaMalloc(&staged, (size_t)(d_pad * m_cap * 2)) != cudaSuccess) return 1;
  if (cudaMalloc((void**)&inv_scale, (size_t)(m_cap * gadm_stage_scale_count(d_pad) * 4)) != cudaSuccess) return 1;
  ws_bytes = gadm_project_workspace_bytes(h, batch, d_pad, proj_dim, 2);
  if (ws_bytes < 0 || cudaMalloc(&workspace, (size_t)ws_bytes) != cudaSuccess) return 1;
  if (cudaMalloc((void**)&out, (size_t)(batch * proj_dim * 4)) != cudaSuccess) return 1;
  for (i = 0, grad_dim = 0; i < 3; ++i) {
    /* the per-example gradients of one parameter tensor: [batch, numel] fp32, as vmap(grad(f)) returns them */
    if (cudaMalloc((void**)&grads[i], (size_t)(batch * numel[i] * 4)) != cudaSuccess) return 1;
    cudaMemset(grads[i], 0, (size_t)(batch * numel[i] * 4));
    blocks[i].ptr = grads[i];
    blocks[i].numel_per_example = numel[i];
    blocks[i].example_stride = numel[i];
    blocks[i].row_offset = grad_dim;
    grad_dim += numel[i];
  }
  /* seed64 = seed + 10^4 * model_id (trak CudaProjector semantics); scale folds the 1/K timestep mean */
  CHECK(gadm_project(h, blocks, 3, GADM_DTYPE_F32, batch, 1.0f / 10.0f, staged, GADM_STAGE_F16G, inv_scale, d_pad, m_cap,
                     proj_dim, 42ull, GADM_PROJ_NORMAL, out, proj_dim, 0, workspace, ws_bytes, 2, NULL));
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  printf("projected %lld x %lld -> %lld x %lld; kernels launched: %lld\n", (long long)batch, (long long)grad_dim,
         (long long)batch, (long long)proj_dim, (long long)gadm_launch_count(h));
  CHECK(gadm_destroy(h));
  return 0;
}
