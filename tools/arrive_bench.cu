// arrive_bench.cu -- cost of a remote (peer-CTA) mbarrier arrive issued by one thread, back to back:
// release.cluster vs relaxed.cluster, and of a local arrive, in SM cycles.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../articulated-object-nerf_b200/csrc/tc_ptx.cuh"
using namespace aon::ptx;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(long long* out, int iters) {
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ float buf[1024];
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), (1u << 20) - 1); fence_mbar_init(); }
  cluster_sync_all();
  const uint32_t remote = mapa(smem_u32(&bar[0]), rank ^ 1);
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) mbar_arrive_cluster(remote);
    long long t1 = clock64();
    for (int i = 0; i < iters; ++i)
      asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
    long long t2 = clock64();
    for (int i = 0; i < iters; ++i) mbar_arrive(smem_u32(&bar[1]));
    long long t3 = clock64();
    // release.cluster arrive with outstanding shared stores in front of it
    for (int i = 0; i < iters; ++i) {
      for (int j = 0; j < 8; ++j) buf[(i * 8 + j) & 1023] = (float)i;
      mbar_arrive_cluster(remote);
    }
    long long t4 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int j = 0; j < 8; ++j) buf[(i * 8 + j) & 1023] = (float)i;
      fence_proxy_async_smem();
      asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
    }
    long long t5 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; }
  }
  cluster_sync_all();
  if (threadIdx.x == 1 && buf[5] == 123.f) out[7] = 1;
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int iters = 256;
  probe<<<148, 128>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("%s  cycles per arrive: remote release.cluster %.1f | remote relaxed.cluster %.1f | local %.1f | 8 STS + remote release %.1f | 8 STS + fence.proxy.async + remote relaxed %.1f\n",
         cudaGetErrorString(e), h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3] / (double)iters, h[4] / (double)iters);
  return 0;
}
