// mma2_bench.cu -- correctness + timing probe for tcgen05.mma.cta_group::2 (CTA pair, M=256) with the
// no-swizzle K-major operand layout the render kernel uses, operands written by ordinary threads of BOTH
// CTAs (generic proxy -> fence.proxy.async -> remote mbarrier arrive on the leader).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma2_bench mma2_bench.cu && ./mma2_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../articulated-object-nerf_b200/csrc/tc_ptx.cuh"
using namespace aon::ptx;

constexpr int K = 64;            // 4 K-steps of 16
constexpr int OFF_B = 32768;     // A: 128 x 64 fp16 = 16 KB at 0; B half: NH x 64 fp16 at 32 KB

__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// A_g: [2 ctas][128][K] fp16 ; B_g: [N][K] fp16 (row n belongs to cta n / (N/2)) ; D_g: [2][128][N] fp32
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __half* A_g, const __half* B_g, float* D_g, int N, int iters, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_ready, bar_done;
  __shared__ uint32_t tmem_slot;
  unsigned char* sm = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const uint32_t rank = cta_rank();
  const int pair = blockIdx.x >> 1;
  const int NH = N / 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar_ready), 8); mbar_init(smem_u32(&bar_done), 1); fence_mbar_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  // operands: generic-proxy stores. element (row, k) at (k/8)*rows*16 + row*16 + (k%8)*2
  if (pair == 0 || true) {
    const __half* Ap = A_g + (size_t)rank * 128 * K;
    for (int i = tid; i < 128 * K; i += 128) {
      const int r = i / K, k = i % K;
      *reinterpret_cast<__half*>(sm + (k >> 3) * 2048 + r * 16 + (k & 7) * 2) = Ap[i];
    }
    const __half* Bp = B_g + (size_t)rank * NH * K;
    for (int i = tid; i < NH * K; i += 128) {
      const int r = i / K, k = i % K;
      *reinterpret_cast<__half*>(sm + OFF_B + (k >> 3) * NH * 16 + r * 16 + (k & 7) * 2) = Bp[i];
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if ((tid & 31) == 0) remote_arrive(mapa(smem_u32(&bar_ready), 0));   // 4 warps x 2 CTAs -> leader's barrier
  if (rank == 0 && tid == 0) {
    while (!try_wait_cluster(smem_u32(&bar_ready), 0)) {}
    tc_fence_after();
    const uint32_t idesc = idesc_f16(256, N, 0);
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm) + OFF_B;
    for (int ks = 0; ks < K / 16; ++ks)
      mma2(tm, desc(a0 + ks * 4096, 2048, 128), desc(b0 + ks * NH * 32, NH * 16, 128), idesc, ks > 0);
    commit2(smem_u32(&bar_done), 3);
  }
  // every CTA waits for the multicast commit, then dumps its 128 x N accumulator
  while (!mbar_try_wait(smem_u32(&bar_done), 0)) {}
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    if (pair == 0)
      for (int i = 0; i < 32; ++i) D_g[((size_t)rank * 128 + tid) * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  // timing: a stream of MMAs on the static operands
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = idesc_f16(256, N, 0);
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm) + OFF_B;
    uint32_t parity = 1;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int ks = i & 3;
        mma2(tm + 256, desc(a0 + ks * 4096, 2048, 128), desc(b0 + ks * NH * 32, NH * 16, 128), idesc, i > 0);
      }
      commit2(smem_u32(&bar_done), 1);
      while (!mbar_try_wait(smem_u32(&bar_done), parity)) {}
      parity ^= 1;
      if (blockIdx.x == 0) cyc[rep] = clock64() - t0;
    }
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
  }
}

int main() {
  for (int N : {256, 128}) {
    std::vector<__half> A(2 * 128 * K), B(N * K);
    std::vector<float> Af(A.size()), Bf(B.size());
    srand(1);
    for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)(rand() % 7 - 3); A[i] = __float2half(Af[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (float)(rand() % 5 - 2); B[i] = __float2half(Bf[i]); }
    __half *dA, *dB; float* dD; long long* dC;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 2 * 128 * N * 4); cudaMalloc(&dC, 64);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 2 * 128 * N * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 512;
    probe<<<148, 128, 100 * 1024>>>(dA, dB, dD, N, iters, dC);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(2 * 128 * N);
    long long h[3] = {0, 0, 0};
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h, dC, 24, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int c = 0; c < 2; ++c)
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int k = 0; k < K; ++k) ref += Af[(c * 128 + r) * K + k] * Bf[n * K + k];
          const float got = D[(c * 128 + r) * N + n];
          const double err = fabs((double)got - ref);
          if (!(err <= 1e-3)) { if (bad < 5) printf("  mismatch cta %d row %d col %d: got %g want %g\n", c, r, n, got, ref); ++bad; }
          if (err > maxerr) maxerr = err;
        }
    printf("cta_group::2 M=256 N=%d: %s  mismatches %d  maxerr %g  cycles/MMA %.1f %.1f %.1f (floor %d)\n", N,
           cudaGetErrorString(e), bad, maxerr, h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, N / 2);
  }
  return 0;
}
