// How many clusters of 576-thread, ~227 KB CTAs (K2's shape) can be co-resident on this GPU?
// nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(576, 1) probe(int* p) {
  extern __shared__ char s[];
  if (p) p[0] = s[0];
}
int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  printf("SMs %d\n", pr.multiProcessorCount);
  const int smem = 231000;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cl : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pr.multiProcessorCount / cl * cl);
    cfg.blockDim = dim3(576);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe, &cfg);
    printf("cluster %2d: max active clusters %d (= %d CTAs) %s\n", cl, n, n * cl,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
