// issue_bench.cu -- how many cycles does ONE thread need to issue a "stage" of the fused kernel?
// cta_group::2, M=256, N=256, K=16 MMAs on static operands; barriers are pre-completed so nothing ever blocks.
// floor = 128 cycles per MMA.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_bench issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../articulated-object-nerf_b200/csrc/tc_ptx.cuh"
using namespace aon::ptx;

__device__ __forceinline__ uint64_t mkd(uint32_t lo32) { return ((uint64_t)(8u | (1u << 14)) << 32) | lo32; }

template <int VAR, int NMMA>
__device__ long long run(uint32_t tm, uint32_t a16, uint32_t b16, uint32_t bar_ready, uint32_t bar_sink, uint32_t bar_done, int iters, uint32_t& parity) {
  const uint32_t idesc = idesc_f16(256, 256, 0);
  const uint32_t A_LBO = (2048u >> 4) << 16, B_LBO = 128u << 16;
  const long long t0 = clock64();
  bool ok_next = true;
  if (VAR == 2) ok_next = mbar_try_wait(bar_ready, 1);
  for (int i = 0; i < iters; ++i) {
    const uint32_t slot = i & 3;
    if (VAR == 1) {
      while (!mbar_try_wait(bar_ready, 1)) {}
      tc_fence_after();
    }
    if (VAR == 2) {
      while (!ok_next) ok_next = mbar_try_wait(bar_ready, 1);
      tc_fence_after();
      ok_next = mbar_try_wait(bar_ready, 1);      // next stage's wait, in flight while the MMAs are issued
    }
    if (VAR == 3) {   // two waits (operand chunk + weight stage), like the first K step of a chunk
      while (!mbar_try_wait(bar_ready, 1)) {}
      tc_fence_after();
      while (!mbar_try_wait(bar_ready + 8, 1)) {}
      tc_fence_after();
    }
    const uint32_t bd = (b16 + slot * 512u) | B_LBO;
#pragma unroll
    for (int m = 0; m < NMMA; ++m)
      mma2_f16_ss(tm, mkd((a16 + (m & 1) * 256u) | A_LBO), mkd(bd + (m >> 1) * 256u), idesc, (i | m) > 0);
    if (VAR >= 1) mma_commit2(bar_sink, 1);
  }
  mma_commit2(bar_done, 1);
  while (!mbar_try_wait(bar_done, parity)) {}
  parity ^= 1;
  return clock64() - t0;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  unsigned char* sm = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1);   // "ready" barriers: phase 0 left pending -> parity 1 is complete
    mbar_init(smem_u32(&bars[2]), 1u << 19);                               // sink for per-stage commits
    mbar_init(smem_u32(&bars[3]), 1);                                      // done
    fence_mbar_init();
  }
  if (threadIdx.x < 32) { tmem_alloc2(smem_u32(&tmem_slot), 512); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t a16 = smem_u32(sm) >> 4, b16 = (smem_u32(sm) + 32768) >> 4;
    uint32_t parity = 0;
    long long r[8];
    r[0] = run<0, 3>(tm, a16, b16, smem_u32(&bars[0]), smem_u32(&bars[2]), smem_u32(&bars[3]), iters, parity);
    r[1] = run<1, 3>(tm, a16, b16, smem_u32(&bars[0]), smem_u32(&bars[2]), smem_u32(&bars[3]), iters, parity);
    r[2] = run<2, 3>(tm, a16, b16, smem_u32(&bars[0]), smem_u32(&bars[2]), smem_u32(&bars[3]), iters, parity);
    r[3] = run<3, 3>(tm, a16, b16, smem_u32(&bars[0]), smem_u32(&bars[2]), smem_u32(&bars[3]), iters, parity);
    r[4] = run<0, 2>(tm, a16, b16, smem_u32(&bars[0]), smem_u32(&bars[2]), smem_u32(&bars[3]), iters, parity);
    r[5] = run<1, 2>(tm, a16, b16, smem_u32(&bars[0]), smem_u32(&bars[2]), smem_u32(&bars[3]), iters, parity);
    r[6] = run<2, 2>(tm, a16, b16, smem_u32(&bars[0]), smem_u32(&bars[2]), smem_u32(&bars[3]), iters, parity);
    r[7] = run<3, 2>(tm, a16, b16, smem_u32(&bars[0]), smem_u32(&bars[2]), smem_u32(&bars[3]), iters, parity);
    if (blockIdx.x == 0) for (int i = 0; i < 8; ++i) out[i] = r[i];
  }
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc2(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 256;
  probe<<<148, 128, 100 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  const char* names[4] = {"MMAs only", "wait+fence+MMAs+commit", "prefetched wait+fence+MMAs+commit", "2 waits+fence+MMAs+commit"};
  printf("%s\n", cudaGetErrorString(e));
  for (int i = 0; i < 8; ++i)
    printf("%d MMAs/stage, %-36s: %.1f cycles per stage (MMA floor %d)\n", i < 4 ? 3 : 2, names[i & 3], h[i] / (double)iters, (i < 4 ? 3 : 2) * 128);
  return 0;
}
