// Probe: how many clusters of size 2 / 4 / 8 with ~200 KB dynamic smem and 768 threads are co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { if (p) *p = 1; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s SMs=%d\n", prop.name, prop.multiProcessorCount);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    for (int threads : {512, 768}) {
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs); cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim = {(unsigned)cs, 1, 1};
      cfg.attrs = a; cfg.numAttrs = 1;
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("cluster %2d threads %d: max active clusters %d (%d SMs) %s\n", cs, threads, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
