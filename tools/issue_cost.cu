// issue_cost.cu -- per-instruction issue cost (one thread, nothing else on the SM) of the pieces of an MMA stage.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../articulated-object-nerf_b200/csrc/tc_ptx.cuh"
using namespace aon::ptx;
__device__ __forceinline__ uint64_t mkd(uint32_t lo32) { return ((uint64_t)(8u | (1u << 14)) << 32) | lo32; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  unsigned char* sm = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1u << 19); mbar_init(smem_u32(&bars[2]), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc2(smem_u32(&tmem_slot), 512); tmem_relinquish2(); }
  fence_proxy_async_smem(); tc_fence_before(); cluster_sync_all(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t a16 = smem_u32(sm) >> 4, b16 = (smem_u32(sm) + 32768) >> 4;
    const uint32_t A_LBO = (2048u >> 4) << 16;
    const uint32_t idesc16 = idesc_f16(256, 16, 0);   // tiny MMA: execution ~8 cycles -> measures issue cost
    long long t[8]; uint32_t acc = 0;
    t[0] = clock64();
    for (int i = 0; i < iters; ++i) acc += mbar_try_wait(smem_u32(&bars[0]), 1);
    t[1] = clock64();
    for (int i = 0; i < iters; ++i) acc += mbar_test_wait(smem_u32(&bars[0]), 1);
    t[2] = clock64();
    for (int i = 0; i < iters; ++i) tc_fence_after();
    t[3] = clock64();
    for (int i = 0; i < iters; ++i) mma_commit2(smem_u32(&bars[1]), 1);
    t[4] = clock64();
    for (int i = 0; i < iters; ++i) mma2_f16_ss(tm, mkd((a16 + (i & 1) * 256u) | A_LBO), mkd((b16 + (i & 3) * 512u) | (8u << 16)), idesc16, i > 0);
    t[5] = clock64();
    for (int i = 0; i < iters; ++i) { mma2_f16_ss(tm, mkd((a16 + (i & 1) * 256u) | A_LBO), mkd((b16 + (i & 3) * 512u) | (8u << 16)), idesc16, 1); mma_commit2(smem_u32(&bars[1]), 1); }
    t[6] = clock64();
    for (int i = 0; i < iters; ++i) acc += ld_acquire_shared_u32(smem_u32(&tmem_slot));
    t[7] = clock64();
    mma_commit2(smem_u32(&bars[2]), 1);
    while (!mbar_try_wait(smem_u32(&bars[2]), 0)) {}
    if (blockIdx.x == 0) { for (int i = 0; i < 7; ++i) out[i] = t[i + 1] - t[i]; out[7] = acc; }
  }
  __syncthreads();
  if (rank == 0 && threadIdx.x < 32) {   // converged warp, elected lane issues
    const uint32_t a16 = smem_u32(sm) >> 4, b16 = (smem_u32(sm) + 32768) >> 4;
    const uint32_t A_LBO = (2048u >> 4) << 16;
    const uint32_t idesc16 = idesc_f16(256, 16, 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) mma2_f16_ss(tm, mkd((a16 + (i & 1) * 256u) | A_LBO), mkd((b16 + (i & 3) * 512u) | (8u << 16)), idesc16, 1);
      __syncwarp();
    }
    long long t1 = clock64();
    for (int i = 0; i < iters; ++i) {
      while (!mbar_try_wait(smem_u32(&bars[0]), 1)) {}
      tc_fence_after();
      if (elect_one()) {
        mma2_f16_ss(tm, mkd((a16 + (i & 1) * 256u) | A_LBO), mkd((b16 + (i & 3) * 512u) | (8u << 16)), idesc16, 1);
        mma2_f16_ss(tm, mkd((a16 + 256u) | A_LBO), mkd((b16 + (i & 3) * 512u) | (8u << 16)), idesc16, 1);
        mma2_f16_ss(tm, mkd((a16 + (i & 1) * 256u) | A_LBO), mkd((b16 + 256u + (i & 3) * 512u) | (8u << 16)), idesc16, 1);
        mma_commit2(smem_u32(&bars[1]), 1);
      }
      __syncwarp();
    }
    long long t2 = clock64();
    if (elect_one()) { mma_commit2(smem_u32(&bars[2]), 1); }
    while (!mbar_try_wait(smem_u32(&bars[2]), 1)) {}
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[8] = t1 - t0; out[9] = t2 - t1; }
  }
  tc_fence_before(); cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc2(tm, 512); }
}
int main() {
  long long* d; cudaMalloc(&d, 128); cudaMemset(d, 0, 128);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 512;
  probe<<<148, 128, 100 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  const char* n[7] = {"try_wait (complete)", "test_wait (complete)", "tcgen05.fence::after", "tcgen05.commit (multicast form, mask 1)", "tcgen05.mma issue (N=16)", "mma + commit", "ld.acquire.shared"};
  printf("%s\n", cudaGetErrorString(e));
  for (int i = 0; i < 7; ++i) printf("%-42s %.1f cycles\n", n[i], h[i] / (double)iters);
  printf("converged warp + elect: mma issue %.1f cycles; stage (wait+fence+3 mma+commit) %.1f cycles\n", h[8] / (double)iters, h[9] / (double)iters);
  return 0;
}
