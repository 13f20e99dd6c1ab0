// Stand-alone probe for compute-sanitizer's synccheck on thread-block clusters (profiles/r02_sanitizer.txt): a cluster of 4
// CTAs, one __syncthreads behind a single-thread branch, two cluster barriers around a DSMEM read -- nothing else.
// nvcc -gencode arch=compute_100a,code=sm_100a -o variants/cluster_sync_probe profiles/cluster_sync_probe.cu
// compute-sanitizer --tool synccheck variants/cluster_sync_probe <clusters>      (148 SMs: try 64 and 2000)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 5) probe(unsigned *out, int spin)
{
  extern __shared__ unsigned dyn[];
  __shared__ unsigned s_val;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    unsigned acc = blockIdx.x;
    for (int i = 0; i < spin; ++i) acc = acc * 1664525u + 1013904223u;      // the single-thread branch takes a while
    s_val = acc | 1u;
    dyn[0] = acc;
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  unsigned a, v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"((unsigned)__cvta_generic_to_shared(&s_val)), "r"((rank + 1u) & 3u));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) out[blockIdx.x] = v;
}

int main(int argc, char **argv)
{
  const int clusters = argc > 1 ? atoi(argv[1]) : 64;
  unsigned *out = nullptr;
  cudaMalloc(&out, sizeof(unsigned) * clusters * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 44 * 1024);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * 4); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 44 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, probe, out, 2000);
  cudaError_t e2 = cudaDeviceSynchronize();
  printf("clusters %d: launch %s, sync %s\n", clusters, cudaGetErrorString(e), cudaGetErrorString(e2));
  return 0;
}
