// contend_bench.cu -- does concurrent shared-memory traffic (epilogue-style STS.128, bulk copies) slow down
// tcgen05.mma.cta_group::2 (M=256,N=256,K=16, SS operands)?  Thread 0 of the leader issues a stream of MMAs;
// NBG background warps per CTA stream 16-byte stores into their own shared-memory region meanwhile.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../articulated-object-nerf_b200/csrc/tc_ptx.cuh"
using namespace aon::ptx;
__device__ __forceinline__ uint64_t mkd(uint32_t lo32) { return ((uint64_t)(8u | (1u << 14)) << 32) | lo32; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(416, 1) probe(int iters, int nbg, int mode, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  unsigned char* sm = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); stop = 0; fence_mbar_init(); }
  if (warp == 0) { tmem_alloc2(smem_u32(&tmem_slot), 512); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 0) {
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = idesc_f16(256, 256, 0);
      const uint32_t a16 = smem_u32(sm) >> 4, b16 = (smem_u32(sm) + 32768) >> 4;
      const uint32_t A_LBO = (2048u >> 4) << 16, B_LBO = 128u << 16;
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t bd = (b16 + (i & 3) * 512u) | B_LBO;
        mma2_f16_ss(tm, mkd((a16 + (i & 1) * 256u) | A_LBO), mkd(bd), idesc, i > 0);
      }
      mma_commit2(smem_u32(&bars[0]), 3);
      while (!mbar_try_wait(smem_u32(&bars[0]), 0)) {}
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      stop = 1;
    } else if (rank == 1 && lane == 0) {
      while (!mbar_try_wait(smem_u32(&bars[0]), 0)) {}
      stop = 1;
    }
  } else if (warp <= nbg) {
    // background: each warp streams 512 contiguous bytes per instruction (conflict free), like the epilogue's
    // operand stores, into a private 8 KB window above 96 KB
    unsigned char* base = sm + 98304 + (warp - 1) * 8192 + lane * 16;
    uint4 v = make_uint4(warp, lane, 0, 0);
    long n = 0;
    while (!stop) {
      if (mode == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(base + k * 512)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) { uint4 t; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(smem_u32(base + k * 512)) : "memory"); v.x ^= t.x; }
      }
      ++n;
    }
    if (n == 123456789 && v.x == 42) out[7] = n;
    if (blockIdx.x == 0 && warp == 1 && lane == 0) out[1] = n;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2048;
  for (int mode = 0; mode < 2; ++mode)
    for (int nbg : {0, 1, 2, 4, 8, 12}) {
      cudaMemset(d, 0, 64);
      probe<<<148, 416, 200 * 1024>>>(iters, nbg, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      const double cyc = h[0] / (double)iters;
      printf("%s  background %s warps/CTA %2d : %.1f cycles per MMA (floor 128); background rate %.1f B/cycle/SM\n", cudaGetErrorString(e),
             mode ? "LDS.128" : "STS.128", nbg, cyc, h[0] ? (double)h[1] * 16 * 512 * nbg / h[0] : 0.0);
    }
  return 0;
}
