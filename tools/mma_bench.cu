// mma_bench.cu -- micro-benchmark: cycles per tcgen05.mma (M=128, kind::f16, cta_group::1) for different
// shared-memory operand layouts.  One CTA per SM, one thread issues a stream of MMAs on static operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu && ./mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../articulated-object-nerf_b200/csrc/tc_ptx.cuh"
using namespace aon::ptx;

struct Cfg { int N; int layout; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; uint32_t a_kstep, b_kstep; int a_tmem; };

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc),
               "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(Cfg c, int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  unsigned char* sm = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm) + 65536;
    const uint32_t idesc = idesc_f16(128, c.N, 0);
    uint32_t parity = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int ks = i & 3;
        if (c.a_tmem) mma_ts(tm, tm + 256 + ks * 8, desc(b0 + ks * c.b_kstep, c.b_lbo, c.b_sbo, c.layout), idesc, i > 0);
        else mma_f16_ss(tm, desc(a0 + ks * c.a_kstep, c.a_lbo, c.a_sbo, c.layout), desc(b0 + ks * c.b_kstep, c.b_lbo, c.b_sbo, c.layout), idesc, i > 0);
      }
      mma_commit(smem_u32(&bar));
      while (!mbar_try_wait(smem_u32(&bar), parity)) {}
      parity ^= 1;
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[rep] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 512;
  struct Named { const char* name; Cfg c; } cfgs[] = {
    // name, {N, layout, a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep}
    {"none  N=128 (k-group slabs: LBO=rows*16, SBO=128)", {128, 0, 2048, 128, 2048, 128, 4096, 4096, 0}},
    {"none  N=256 (k-group slabs)                      ", {256, 0, 2048, 128, 4096, 128, 4096, 8192, 0}},
    {"none  N=256 (8-row groups of 2 core mats: LBO=128, SBO=256)", {256, 0, 128, 256, 128, 256, 4096, 8192, 0}},
    {"sw32  N=256 (SBO=256)", {256, 6, 16, 256, 16, 256, 4096, 8192, 0}},
    {"sw64  N=256 (SBO=512, k-step +32B)", {256, 4, 16, 512, 16, 512, 32, 32, 0}},
    {"sw128 N=256 (SBO=1024, k-step +32B)", {256, 2, 16, 1024, 16, 1024, 32, 32, 0}},
    {"sw128 N=128 (SBO=1024, k-step +32B)", {128, 2, 16, 1024, 16, 1024, 32, 32, 0}},
    {"sw128 N=64  (SBO=1024, k-step +32B)", {64, 2, 16, 1024, 16, 1024, 32, 32, 0}},
    {"A in TMEM, B none  N=256 (k-group slabs)", {256, 0, 0, 0, 4096, 128, 0, 8192, 1}},
    {"A in TMEM, B sw128 N=256", {256, 2, 0, 0, 16, 1024, 0, 32, 1}},
    {"A in TMEM, B sw128 N=128", {128, 2, 0, 0, 16, 1024, 0, 32, 1}},
  };
  for (auto& n : cfgs) {
    bench<<<148, 128, 200 * 1024>>>(n.c, iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3] = {0, 0, 0};
    cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost);
    printf("%-62s : %s  cycles/MMA %.1f %.1f %.1f  (ideal %d)\n", n.name, cudaGetErrorString(e), h[0] / (double)iters,
           h[1] / (double)iters, h[2] / (double)iters, n.c.N / 2);
  }
  return 0;
}
