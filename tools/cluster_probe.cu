// How many thread-block clusters of a 1-CTA-per-SM kernel (200 KB dynamic smem, 256 threads) can be co-resident on
// this GPU, per cluster size.  Decides whether a 4- or 8-CTA multicast cluster can still cover all 148 SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/cluster_probe tools/cluster_probe.cu && gpurun_out/cluster_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe_kernel(int* out) {
  extern __shared__ unsigned char smem[];
  if (out && threadIdx.x == 0) out[blockIdx.x] = smem[0];
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.multiProcessorCount / cs * cs);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = cs;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
