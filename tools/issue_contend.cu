// issue_contend.cu -- is the single-thread tcgen05 issue loop slowed down by busy warps on the SAME scheduler
// sub-partition (warp % 4), by busy warps elsewhere, or by shared-memory barrier traffic?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../articulated-object-nerf_b200/csrc/tc_ptx.cuh"
using namespace aon::ptx;
__device__ __forceinline__ uint64_t mkd(uint32_t lo32) { return ((uint64_t)(8u | (1u << 14)) << 32) | lo32; }

// mode 0: no background; 1: ALU-busy warps on the issuer's sub-partition (warps 4, 8); 2: ALU-busy warps on other
// sub-partitions (warps 5, 6, 9, 10); 3: warps everywhere doing mbarrier arrives on local barriers; 4: STS + fence.proxy.async
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) probe(int iters, int mode, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[20];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  unsigned char* sm = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 20; ++i) mbar_init(smem_u32(&bars[i]), i < 3 ? 1 : (1u << 19)); stop = 0; fence_mbar_init(); }
  if (warp == 1) { tmem_alloc2(smem_u32(&tmem_slot), 512); tmem_relinquish2(); }
  fence_proxy_async_smem(); tc_fence_before(); cluster_sync_all(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 0) {
    if (rank == 0 && lane == 0) {
      const uint32_t a16 = smem_u32(sm) >> 4, b16 = (smem_u32(sm) + 32768) >> 4;
      const uint32_t A_LBO = (2048u >> 4) << 16;
      const uint32_t idesc16 = idesc_f16(256, 16, 0);
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        while (!mbar_try_wait(smem_u32(&bars[0]), 1)) {}
        tc_fence_after();
        mma2_f16_ss(tm, mkd((a16 + (i & 1) * 256u) | A_LBO), mkd((b16 + (i & 3) * 512u) | (8u << 16)), idesc16, 1);
        mma2_f16_ss(tm, mkd((a16 + 256u) | A_LBO), mkd((b16 + (i & 3) * 512u) | (8u << 16)), idesc16, 1);
        mma2_f16_ss(tm, mkd((a16 + (i & 1) * 256u) | A_LBO), mkd((b16 + 256u + (i & 3) * 512u) | (8u << 16)), idesc16, 1);
        mma_commit2(smem_u32(&bars[3]), 1);
      }
      const long long t1 = clock64();
      mma_commit2(smem_u32(&bars[1]), 3);
      while (!mbar_try_wait(smem_u32(&bars[1]), 0)) {}
      if (blockIdx.x == 0) out[0] = t1 - t0;
      stop = 1;
    } else if (rank == 1 && lane == 0) {
      while (!mbar_try_wait(smem_u32(&bars[1]), 0)) {}
      stop = 1;
    }
  } else if (warp >= 4) {
    const bool same = (warp & 3) == 0;
    float x = lane * 0.001f, y = 1.0001f;
    if ((mode == 1 && same && warp <= 8) || (mode == 2 && !same && warp <= 10 && (warp & 3) != 3)) {
      while (!stop) {
#pragma unroll
        for (int k = 0; k < 64; ++k) { x = fmaf(x, y, 0.5f); y = fmaf(y, 0.999f, 0.001f); }
      }
    } else if (mode == 3 && warp < 12) {
      while (!stop) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&bars[4 + (warp & 7)])); }
      }
    } else if (mode == 4 && warp < 12) {
      unsigned char* base = sm + 98304 + (warp - 4) * 8192 + lane * 16;
      while (!stop) {
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(smem_u32(base + k * 512)), "r"(k) : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[4 + (warp & 7)]));
      }
    }
    if (x == 123.456f) out[5] = (long long)y;
  }
  tc_fence_before(); cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc2(tm, 512); }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 512;
  const char* names[5] = {"no background", "2 ALU-busy warps on the issuer's sub-partition", "6 ALU-busy warps on the other sub-partitions",
                          "8 warps doing mbarrier arrives", "8 warps: 8x STS.128 + fence.proxy.async + arrive"};
  for (int mode = 0; mode < 5; ++mode) {
    cudaMemset(d, 0, 64);
    probe<<<148, 512, 200 * 1024>>>(iters, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%s  %-52s: %.1f cycles per stage (wait + fence + 3 tiny MMAs + commit)\n", cudaGetErrorString(e), names[mode], h[0] / (double)iters);
  }
  return 0;
}
