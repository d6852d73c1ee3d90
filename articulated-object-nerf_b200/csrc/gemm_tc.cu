// gemm_tc.cu -- tcgen05 GEMMs of the TRAINING path (stage 2): forward, dgrad and wgrad of the MLP's
// nn.Linear layers (model.py:99-118) on the 5th-gen tensor cores, fp16 hi+lo split operands (3 MMAs per K step, fp32
// accumulation in TMEM) like the parity mode of the fused render kernel.
//
// Data layout.  Every activation / gradient matrix X[rows, feat] lives in HBM ONCE, as two 16-bit planes (hi, lo) in the
// "k-group packed" layout  PK(rows, feat) = [rows/128][feat/8][128][8]:  a 128-row x 8-feature slab is 2 KB contiguous.
//   * as the K-major A operand of forward / dgrad (rows = samples, contraction = features) a (row tile, K chunk) is one
//     contiguous block that lands in shared memory in the UMMA canonical no-swizzle layout with LBO = 2048, SBO = 128;
//   * as an MN-major operand of wgrad (dW = dY^T X: contraction = samples) THE SAME bytes are the canonical MN-major
//     layout ((8,m),(8,k)):((1,SBO),(8,LBO)) with SBO = slab stride, LBO = 128 -- no transposed copy is ever written.
// Weights are packed per step into PW(rows, contraction) = [contraction/8][rows][8] (K-major B operand), once as W
// (forward) and once as W^T (dgrad).
//
// One CTA = one 128 x N output tile (N <= 256): warp 0 = producer (TMA global -> 2-stage shared-memory ring, mbarrier
// complete_tx: the contiguous A block of forward / dgrad as one 1-D bulk copy, everything strided -- the weight rows of a K
// chunk, the 32-sample slabs of both wgrad operands -- as ONE tensor-map box per plane), warp 1 = MMA issuer (tcgen05.mma
// cta_group::1 kind::f16, tcgen05.commit releases ring slots / publishes the accumulator); both run as converged warps with
// one elected lane executing the TMA / MMA instructions (operands stay in uniform registers), warps 2-5 = epilogue (tcgen05.ld, bias / ReLU / ReLU-mask, pack to
// hi+lo, coalesced 16-byte stores).  96 KB of shared memory and <= 256 TMEM columns per CTA -> two CTAs per SM, so one
// CTA's epilogue overlaps the other's MMAs.  wgrad CTAs loop over their share of the sample tiles (split-K) and write fp32
// partial tiles that wgrad_reduce_kernel sums in a fixed order (deterministic).
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "aon_common.cuh"
#include "tc_ptx.cuh"

namespace aon {
using namespace ptx;

constexpr int GT_THREADS = 192;
constexpr int GT_KC = 32;                       // contraction elements per pipeline stage
constexpr int GT_NS = 2;                        // stages
constexpr uint32_t GT_A_PLANE = 128 * GT_KC * 2;   // 8 KB: 128 rows x 32 k x 2 B
constexpr uint32_t GT_B_PLANE = 256 * GT_KC * 2;   // 16 KB: up to 256 rows
constexpr uint32_t GT_STAGE = 2 * GT_A_PLANE + 2 * GT_B_PLANE;   // 48 KB
constexpr uint32_t GT_SMEM = GT_NS * GT_STAGE + 1024;             // + alignment slack

// Tensor maps of the wgrad (TN) operands (CUtensorMap, 64-byte aligned, read from kernel parameter space): tm[plane] =
// operand A, tm[2 + plane] = operand B.  A PK tensor [tiles][feat/8][128][8] x 16 bit is described in 8-byte elements as
// {256, feat/8, tiles} (the 128 x 8 slab of a feature group is 2 KB contiguous); a stage's box {GT_KC * 2, groups, 1} takes
// GT_KC samples (512 contiguous bytes) of `groups` feature groups and lands as [group][sample][8] -- one instruction instead
// of 16 + N/8 separate 512-byte copies per plane.  (An inner box of 8 x 16-bit = 16 bytes moves the same bytes as hundreds of
// 16-byte rows and is slower than the separate copies were: measured.)  Forward / dgrad (NT) operands are contiguous per K
// chunk and stay 1-D bulk copies.
struct alignas(64) GtParams {
  CUtensorMap tm[4];
  AonGemm g;
};

__device__ __forceinline__ void gt_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// fp16 operand planes carry power-of-two scaled values (activations x8, weights x64, gradients x2^k): a value beyond the
// fp16 range would become inf and poison every accumulator it meets as NaN.  Saturate instead (finite, visibly clipped);
// lit.Trainer checks the gradient buffer for non-finite values at its logging interval and raises.
__device__ __forceinline__ float sat_h(float v) { return fminf(fmaxf(v, -65504.0f), 65504.0f); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  return (uint32_t)__half_as_ushort(__float2half_rn(sat_h(a))) | ((uint32_t)__half_as_ushort(__float2half_rn(sat_h(b))) << 16);
}

__global__ void __launch_bounds__(GT_THREADS, 2) gemm_tc_kernel(const __grid_constant__ GtParams P) {
  extern __shared__ unsigned char gt_smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2 * GT_NS + 1];
  __shared__ uint32_t s_tmem;
  __shared__ float s_colsum[4][256];     // per epilogue warp: column sums of its 32 rows (bias gradients)
  __shared__ float s_bias[256];
  const AonGemm& g = P.g;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler: the control warps stay converged
  const uint32_t sm0 = (smem_u32(gt_smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(s_bar);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (GT_NS + s); };
  const uint32_t accum = bar0 + 8u * (2 * GT_NS);
  const int N = g.N;
  const uint32_t tmem_cols = N <= 32 ? 32u : (N <= 64 ? 64u : (N <= 128 ? 128u : 256u));
  const bool x3 = g.x3 != 0;

  if (tid == 0) {
    for (int s = 0; s < GT_NS; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(accum, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(&s_tmem), tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, s_tmem, 0);

  // ---- work decomposition -----------------------------------------------------------------------------------------
  // NT: blockIdx.x = row tile; stages = for each segment, K chunks of GT_KC.
  // TN: blockIdx.x = 128-feature tile of operand A (output rows), blockIdx.y = split; stages = row tiles x 4 quarter tiles.
  const bool tn = g.mode == AON_GEMM_TN;
  int t_begin = 0, t_end = 0;
  if (tn) {
    t_begin = (int)blockIdx.y * g.tiles_per_split;
    t_end = min(g.m_tiles, t_begin + g.tiles_per_split);
  }
  long n_stage_total = 0;
  if (tn) n_stage_total = (long)max(0, t_end - t_begin) * (128 / GT_KC);
  else for (int sgi = 0; sgi < g.nseg; ++sgi) n_stage_total += (g.kext[sgi] + GT_KC - 1) / GT_KC;

  if (warp == 0) {
    // ================= producer (converged warp; one elected lane issues the copies) =================
    long it = 0;
    const int planes = x3 ? 2 : 1;
    if (!tn) {
      const long tile = blockIdx.x;
      for (int sgi = 0; sgi < g.nseg; ++sgi) {
        const char* a_pl[2] = {(const char*)g.a_hi[sgi], (const char*)g.a_lo[sgi]};
        const char* b_pl[2] = {(const char*)g.b_hi[sgi], (const char*)g.b_lo[sgi]};
        const long a_tile = tile * (long)(g.a_feat[sgi] / 8) * 2048;
        for (int k0 = 0; k0 < g.kext[sgi]; k0 += GT_KC, ++it) {
          const int s = (int)(it % GT_NS);
          const int kc = min(GT_KC, g.kext[sgi] - k0), nkg = kc / 8;
          gt_wait(empty(s), (uint32_t)(((it / GT_NS) & 1) ^ 1));
          if (elect_one()) {
            const uint32_t a_bytes = (uint32_t)nkg * 2048u, b_piece = (uint32_t)N * 16u;
            mbar_arrive_expect_tx(full(s), (uint32_t)planes * (a_bytes + (uint32_t)nkg * b_piece));
            const uint32_t st = sm0 + (uint32_t)s * GT_STAGE;
            const bool whole_rows = g.b_row0[sgi] == 0 && N == g.b_feat[sgi];   // the k-groups of the chunk are contiguous
            for (int pl = 0; pl < planes; ++pl) {
              bulk_g2s(st + (uint32_t)pl * GT_A_PLANE, a_pl[pl] + a_tile + (long)((g.a_off[sgi] + k0) / 8) * 2048, a_bytes, full(s));
              const char* bsrc = b_pl[pl] + ((long)((g.b_off[sgi] + k0) / 8) * g.b_feat[sgi] + g.b_row0[sgi]) * 16;
              const uint32_t bdst = st + 2 * GT_A_PLANE + (uint32_t)pl * GT_B_PLANE;
              if (whole_rows) {
                bulk_g2s(bdst, bsrc, (uint32_t)nkg * b_piece, full(s));
              } else {
                for (int kg = 0; kg < nkg; ++kg) bulk_g2s(bdst + (uint32_t)kg * b_piece, bsrc + (long)kg * g.b_feat[sgi] * 16, b_piece, full(s));
              }
            }
          }
        }
      }
    } else {
      const int a_ng0 = g.a_off[0] / 8 + (int)blockIdx.x * 16, b_ng0 = g.b_off[0] / 8;
      const uint32_t stage_bytes = (uint32_t)planes * (uint32_t)(16 + N / 8) * (GT_KC * 16u);
      for (int t = t_begin; t < t_end; ++t) {
        for (int q4 = 0; q4 < 128 / GT_KC; ++q4, ++it) {
          const int s = (int)(it % GT_NS);
          gt_wait(empty(s), (uint32_t)(((it / GT_NS) & 1) ^ 1));
          if (elect_one()) {
            mbar_arrive_expect_tx(full(s), stage_bytes);
            const uint32_t st = sm0 + (uint32_t)s * GT_STAGE;
            for (int pl = 0; pl < planes; ++pl) {
              // 32 samples x 128 features of A and x N features of B: one box each, landing as [feature group][sample][8]
              tma_load_3d(st + (uint32_t)pl * GT_A_PLANE, &P.tm[pl], q4 * (GT_KC * 2), a_ng0, t, full(s));
              tma_load_3d(st + 2 * GT_A_PLANE + (uint32_t)pl * GT_B_PLANE, &P.tm[2 + pl], q4 * (GT_KC * 2), b_ng0, t, full(s));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp; one elected lane issues) =================
    {
      const uint32_t idesc = idesc_f16(128, N, 0) | (tn ? ((1u << 15) | (1u << 16)) : 0u);
      uint32_t accumulate = 0;
      long it = 0;
      auto issue_stage = [&](int s, int kc) {
        const uint32_t st = sm0 + (uint32_t)s * GT_STAGE;
        const uint32_t a_hi = st, a_lo = st + GT_A_PLANE, b_hi = st + 2 * GT_A_PLANE, b_lo = b_hi + GT_B_PLANE;
        for (int kk = 0; kk < kc / 16; ++kk) {
          uint32_t a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo;
          if (!tn) {   // K-major: [kg][rows][8]
            a_off = (uint32_t)kk * 4096u; a_lbo = 2048u; a_sbo = 128u;
            b_off = (uint32_t)kk * 2u * (uint32_t)N * 16u; b_lbo = (uint32_t)N * 16u; b_sbo = 128u;
          } else {     // MN-major: [ng][GT_KC rows][8]: MN-group stride (SBO) = GT_KC*16, K-group stride (LBO) = 128
            a_off = (uint32_t)kk * 256u; a_lbo = 128u; a_sbo = GT_KC * 16u;
            b_off = a_off; b_lbo = 128u; b_sbo = GT_KC * 16u;
          }
          mma_f16_ss(tmem_base, smem_desc(a_hi + a_off, a_lbo, a_sbo), smem_desc(b_hi + b_off, b_lbo, b_sbo), idesc, accumulate);
          accumulate = 1;
          if (x3) {
            mma_f16_ss(tmem_base, smem_desc(a_lo + a_off, a_lbo, a_sbo), smem_desc(b_hi + b_off, b_lbo, b_sbo), idesc, 1);
            mma_f16_ss(tmem_base, smem_desc(a_hi + a_off, a_lbo, a_sbo), smem_desc(b_lo + b_off, b_lbo, b_sbo), idesc, 1);
          }
        }
      };
      if (!tn) {
        for (int sgi = 0; sgi < g.nseg; ++sgi)
          for (int k0 = 0; k0 < g.kext[sgi]; k0 += GT_KC, ++it) {
            const int s = (int)(it % GT_NS);
            gt_wait(full(s), (uint32_t)((it / GT_NS) & 1));
            tc_fence_after();
            if (elect_one()) {
              issue_stage(s, min(GT_KC, g.kext[sgi] - k0));
              mma_commit(empty(s));
            }
            accumulate = 1;
          }
      } else {
        for (; it < n_stage_total; ++it) {
          const int s = (int)(it % GT_NS);
          gt_wait(full(s), (uint32_t)((it / GT_NS) & 1));
          tc_fence_after();
          if (elect_one()) {
            issue_stage(s, GT_KC);
            mma_commit(empty(s));
          }
          accumulate = 1;
        }
      }
      if (elect_one()) mma_commit(accum);
    }
  } else {
    // ================= epilogue (warps 2..5: TMEM lane quadrant = warp % 4) =================
    const int quad = warp & 3, row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const long tile = blockIdx.x;
    // Everything the epilogue needs from global memory is fetched while the main loop runs: the bias vector goes to shared
    // memory and this thread's row of the ReLU mask is compressed to one bit per feature (8 registers).
    uint32_t mbits[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    if (g.epi == AON_GEMM_EPI_LINEAR) {
      for (int c = tid - 64; c < N; c += 128) s_bias[c] = g.bias ? __ldg(g.bias + c) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    } else if (g.epi == AON_GEMM_EPI_MASK && g.mask_bits) {
      const uint32_t* bp = g.mask_bits + (tile * 128 + row) * (long)(N / 32);
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) if (cc * 32 < N) mbits[cc] = __ldg(bp + cc);
    } else if (g.epi == AON_GEMM_EPI_MASK && g.mask_hi) {
      const char* mp = (const char*)g.mask_hi + ((tile * (g.mask_feat / 8) + g.mask_off / 8) * 128 + row) * 16;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        if (cc * 32 < N) {
          uint32_t bits = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(mp + (long)(cc * 4 + q) * 2048));
            const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t lo16 = w[j] & 0xffffu, hi16 = w[j] >> 16;       // fp16 > 0: magnitude bits set, sign clear
              bits |= (uint32_t)((lo16 & 0x7fffu) != 0 && (lo16 & 0x8000u) == 0) << (q * 8 + 2 * j);
              bits |= (uint32_t)((hi16 & 0x7fffu) != 0 && (hi16 & 0x8000u) == 0) << (q * 8 + 2 * j + 1);
            }
          }
          mbits[cc] = bits;
        }
      }
    }
    if (n_stage_total > 0) {
      gt_wait(accum, 0);
      tc_fence_after();
    }
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      const int nc = min(32, N - c0);          // 16 or 32
      if (n_stage_total > 0) {
        tmem_ld32(lane_base + (uint32_t)c0, r);  // N = 16: the upper 16 columns are allocated (32 columns) but unused
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (g.reserved & 1) continue;              // debug: drain-only epilogue (timing experiments)
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * g.inv_scale;
      if (g.epi == AON_GEMM_EPI_PARTIAL) {
        float* dst = g.partial + (((long)blockIdx.y * gridDim.x + blockIdx.x) * 128 + row) * N + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (i < nc) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        continue;
      }
      if (g.epi == AON_GEMM_EPI_LINEAR) {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (i < nc) v[i] += s_bias[c0 + i];
        if (g.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          if (g.relu_bits_out) {
            uint32_t bits = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) bits |= (uint32_t)(v[i] > 0.f) << i;
            g.relu_bits_out[(tile * 128 + row) * (long)(N / 32) + (c0 >> 5)] = bits;
          }
        }
      } else {   // AON_GEMM_EPI_MASK: pass where the forward activation (hi plane of PK(rows, mask_feat)) is > 0
        uint32_t bits = 0xffffffffu;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) if (cc == (c0 >> 5)) bits = mbits[cc];      // register select (no local-memory indexing)
#pragma unroll
        for (int i = 0; i < 32; ++i) if (!((bits >> i) & 1u)) v[i] = 0.f;
      }
      if (g.colsum) {
        // 32 columns x 32 lanes -> lane l holds the sum of column c0 + l over this warp's rows: a transpose-reduce butterfly,
        // 31 shuffles (each step halves the columns a lane carries and doubles the rows they cover)
        float t[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = (i < nc) ? v[i] : 0.f;
#pragma unroll
        for (int k = 16; k >= 1; k >>= 1) {
          const bool up = (lane & k) != 0;
#pragma unroll
          for (int i = 0; i < k; ++i) {
            const float send = up ? t[i] : t[i + k];
            const float keep = up ? t[i + k] : t[i];
            t[i] = keep + __shfl_xor_sync(0xffffffffu, send, k);
          }
        }
        s_colsum[quad][c0 + lane] = t[0];
      }
      if (g.out_f32) {
        float* dst = g.out_f32 + (tile * 128 + row) * g.ldc;
#pragma unroll
        for (int i = 0; i < 32; ++i) if (c0 + i < g.n_valid) dst[c0 + i] = v[i];
      }
      if (g.out_hi) {
        const long base = ((tile * (g.out_feat / 8) + (g.out_off + c0) / 8) * 128 + row) * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q * 8 < nc) {
            float s[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = v[q * 8 + j] * g.out_scale;
            uint4 h;
            h.x = pack_h2(s[0], s[1]); h.y = pack_h2(s[2], s[3]); h.z = pack_h2(s[4], s[5]); h.w = pack_h2(s[6], s[7]);
            *reinterpret_cast<uint4*>((char*)g.out_hi + base + (long)q * 2048) = h;
            if (g.out_lo) {
              const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
              float l[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                l[2 * j] = s[2 * j] - __half2float(__ushort_as_half((unsigned short)(hw[j] & 0xffffu)));
                l[2 * j + 1] = s[2 * j + 1] - __half2float(__ushort_as_half((unsigned short)(hw[j] >> 16)));
              }
              uint4 lo;
              lo.x = pack_h2(l[0], l[1]); lo.y = pack_h2(l[2], l[3]); lo.z = pack_h2(l[4], l[5]); lo.w = pack_h2(l[6], l[7]);
              *reinterpret_cast<uint4*>((char*)g.out_lo + base + (long)q * 2048) = lo;
            }
          }
        }
      }
    }
    if (g.colsum) {
      asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps
      const int e = tid - 64;                             // 0..127
      for (int c = e; c < N; c += 128)
        g.colsum[tile * N + c] = (s_colsum[0][c] + s_colsum[1][c]) + (s_colsum[2][c] + s_colsum[3][c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ---- persistent NT kernel (the large forward / dgrad GEMMs: N = 128 or 256, at least one row tile per SM) -------------------
// One CTA per SM walks row tiles blockIdx.x, + gridDim.x, ...; the 4-stage ring runs continuously across tiles and the
// accumulator is double buffered in TMEM (2 x N columns), so the epilogue of tile i overlaps the loads and MMAs of tile i + 1.
// EIGHT epilogue warps (two per TMEM lane quadrant, even / odd 32-column chunks) keep the epilogue shorter than the main loop --
// with four (and global-memory reads on its critical path) the first persistent version was epilogue-bound and slower than the
// two-CTAs-per-SM kernel above (tools/experiments/README.md).
constexpr int GP_NS = 4;
constexpr int GP_EPI_WARPS = 8;
constexpr int GP_THREADS = 64 + 32 * GP_EPI_WARPS;                 // 320
constexpr int GP_EPI_THREADS = 32 * GP_EPI_WARPS;                  // 256
constexpr uint32_t GP_SMEM = GP_NS * GT_STAGE + 1024;

__global__ void __launch_bounds__(GP_THREADS, 1) gemm_tc_nt_persistent_kernel(const __grid_constant__ GtParams P) {
  extern __shared__ unsigned char gt_smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2 * GP_NS + 4];
  __shared__ uint32_t s_tmem;
  __shared__ float s_colsum[2][4][256];  // [tile parity][lane quadrant]: column sums of the quadrant's 32 rows (bias gradients)
  __shared__ float s_bias[256];
  const AonGemm& g = P.g;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t sm0 = (smem_u32(gt_smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(s_bar);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (GP_NS + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (2 * GP_NS + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (2 * GP_NS + 2 + b); };
  const int N = g.N;                                   // 128 or 256
  const uint32_t buf_cols = (uint32_t)N, tmem_cols = 2u * buf_cols;
  const bool x3 = g.x3 != 0;

  if (tid == 0) {
    for (int s = 0; s < GP_NS; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), GP_EPI_THREADS); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(&s_tmem), tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, s_tmem, 0);
  const long tile0 = blockIdx.x, tile_step = gridDim.x, tile_end = g.m_tiles;

  if (warp == 0) {
    // ================= producer (converged warp; one elected lane issues the copies) =================
    long it = 0;
    const int planes = x3 ? 2 : 1;
    for (long tile = tile0; tile < tile_end; tile += tile_step) {
      for (int sgi = 0; sgi < g.nseg; ++sgi) {
        const char* a_pl[2] = {(const char*)g.a_hi[sgi], (const char*)g.a_lo[sgi]};
        const char* b_pl[2] = {(const char*)g.b_hi[sgi], (const char*)g.b_lo[sgi]};
        const long a_tile = tile * (long)(g.a_feat[sgi] / 8) * 2048;
        for (int k0 = 0; k0 < g.kext[sgi]; k0 += GT_KC, ++it) {
          const int s = (int)(it % GP_NS);
          const int kc = min(GT_KC, g.kext[sgi] - k0), nkg = kc / 8;
          gt_wait(empty(s), (uint32_t)(((it / GP_NS) & 1) ^ 1));
          if (elect_one()) {
            const uint32_t a_bytes = (uint32_t)nkg * 2048u, b_piece = (uint32_t)N * 16u;
            mbar_arrive_expect_tx(full(s), (uint32_t)planes * (a_bytes + (uint32_t)nkg * b_piece));
            const uint32_t st = sm0 + (uint32_t)s * GT_STAGE;
            const bool whole_rows = g.b_row0[sgi] == 0 && N == g.b_feat[sgi];   // the k-groups of the chunk are contiguous
            for (int pl = 0; pl < planes; ++pl) {
              bulk_g2s(st + (uint32_t)pl * GT_A_PLANE, a_pl[pl] + a_tile + (long)((g.a_off[sgi] + k0) / 8) * 2048, a_bytes, full(s));
              const char* bsrc = b_pl[pl] + ((long)((g.b_off[sgi] + k0) / 8) * g.b_feat[sgi] + g.b_row0[sgi]) * 16;
              const uint32_t bdst = st + 2 * GT_A_PLANE + (uint32_t)pl * GT_B_PLANE;
              if (whole_rows) {
                bulk_g2s(bdst, bsrc, (uint32_t)nkg * b_piece, full(s));
              } else {
                for (int kg = 0; kg < nkg; ++kg) bulk_g2s(bdst + (uint32_t)kg * b_piece, bsrc + (long)kg * g.b_feat[sgi] * 16, b_piece, full(s));
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp; one elected lane issues) =================
    {
      const uint32_t idesc = idesc_f16(128, N, 0);
      long it = 0, lt = 0;
      for (long tile = tile0; tile < tile_end; tile += tile_step, ++lt) {
        const int b = (int)(lt & 1);
        gt_wait(acc_empty(b), (uint32_t)(((lt >> 1) & 1) ^ 1));     // the epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)b * buf_cols;
        uint32_t accumulate = 0;
        for (int sgi = 0; sgi < g.nseg; ++sgi)
          for (int k0 = 0; k0 < g.kext[sgi]; k0 += GT_KC, ++it) {
            const int s = (int)(it % GP_NS);
            gt_wait(full(s), (uint32_t)((it / GP_NS) & 1));
            tc_fence_after();
            const uint32_t st = sm0 + (uint32_t)s * GT_STAGE;
            const uint32_t a_hi = st, a_lo = st + GT_A_PLANE, b_hi = st + 2 * GT_A_PLANE, b_lo = b_hi + GT_B_PLANE;
            const int kc = min(GT_KC, g.kext[sgi] - k0);
            if (elect_one()) {
              uint32_t acc = accumulate;
              for (int kk = 0; kk < kc / 16; ++kk) {
                const uint32_t a_off = (uint32_t)kk * 4096u, b_off = (uint32_t)kk * 2u * (uint32_t)N * 16u, b_lbo = (uint32_t)N * 16u;
                mma_f16_ss(d_tmem, smem_desc(a_hi + a_off, 2048u, 128u), smem_desc(b_hi + b_off, b_lbo, 128u), idesc, acc);
                acc = 1;
                if (x3) {
                  mma_f16_ss(d_tmem, smem_desc(a_lo + a_off, 2048u, 128u), smem_desc(b_hi + b_off, b_lbo, 128u), idesc, 1);
                  mma_f16_ss(d_tmem, smem_desc(a_hi + a_off, 2048u, 128u), smem_desc(b_lo + b_off, b_lbo, 128u), idesc, 1);
                }
              }
              mma_commit(empty(s));
            }
            accumulate = 1;
          }
        if (elect_one()) mma_commit(acc_full(b));
      }
    }
  } else {
    // ================= epilogue: warps 2..9, TMEM lane quadrant = warp % 4, chunk parity = (warp - 2) / 4 =================
    const int quad = warp & 3, half = (warp - 2) >> 2, row = quad * 32 + lane, e = tid - 64;
    if (g.epi == AON_GEMM_EPI_LINEAR) {
      for (int c = e; c < N; c += GP_EPI_THREADS) s_bias[c] = g.bias ? __ldg(g.bias + c) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const int n_chunks = N / 32;
    long lt = 0;
    float cs_acc = 0.f;
    static_assert(GP_EPI_THREADS >= 256, "one epilogue thread per output column");
    for (long tile = tile0; tile < tile_end; tile += tile_step, ++lt) {
      const int ab = (int)(lt & 1);
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)ab * buf_cols;
      // the ReLU mask of this tile's row, one bit per feature, fetched before the accumulator is ready
      uint32_t mbits[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (g.epi == AON_GEMM_EPI_MASK && g.mask_bits) {
        const uint32_t* bp = g.mask_bits + (tile * 128 + row) * (long)(N / 32);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) if (cc < n_chunks) mbits[cc] = __ldg(bp + cc);
      } else if (g.epi == AON_GEMM_EPI_MASK && g.mask_hi) {
        const char* mp = (const char*)g.mask_hi + ((tile * (g.mask_feat / 8) + g.mask_off / 8) * 128 + row) * 16;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          if (cc < n_chunks && (cc & 1) == half) {
            uint32_t bits = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 m = __ldg(reinterpret_cast<const uint4*>(mp + (long)(cc * 4 + q) * 2048));
              const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t lo16 = w[j] & 0xffffu, hi16 = w[j] >> 16;
                bits |= (uint32_t)((lo16 & 0x7fffu) != 0 && (lo16 & 0x8000u) == 0) << (q * 8 + 2 * j);
                bits |= (uint32_t)((hi16 & 0x7fffu) != 0 && (hi16 & 0x8000u) == 0) << (q * 8 + 2 * j + 1);
              }
            }
            mbits[cc] = bits;
          }
        }
      }
      gt_wait(acc_full(ab), (uint32_t)((lt >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int cj = 0; cj < 4; ++cj) {
        const int ci = 2 * cj + half;
        if (ci < n_chunks) {
          const int c0 = ci * 32;
          uint32_t r[32];
          tmem_ld32(lane_base + (uint32_t)c0, r);
          tmem_ld_wait();
          if (ci + 2 >= n_chunks) {                  // this warp's last read of the buffer: hand it back to the MMA issuer
            tc_fence_before();
            mbar_arrive(acc_empty(ab));
          }
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * g.inv_scale;
          if (g.epi == AON_GEMM_EPI_LINEAR) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += s_bias[c0 + i];
            if (g.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              if (g.relu_bits_out) {
                uint32_t bits = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) bits |= (uint32_t)(v[i] > 0.f) << i;
                g.relu_bits_out[(tile * 128 + row) * (long)(N / 32) + ci] = bits;
              }
            }
          } else {
            const uint32_t bits = half ? mbits[2 * cj + 1] : mbits[2 * cj];      // cj is unrolled: a register select
#pragma unroll
            for (int i = 0; i < 32; ++i) if (!((bits >> i) & 1u)) v[i] = 0.f;
          }
          if (g.colsum) {
            float t[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = v[i];
#pragma unroll
            for (int k = 16; k >= 1; k >>= 1) {
              const bool up = (lane & k) != 0;
#pragma unroll
              for (int i = 0; i < k; ++i) {
                const float send = up ? t[i] : t[i + k];
                const float keep = up ? t[i + k] : t[i];
                t[i] = keep + __shfl_xor_sync(0xffffffffu, send, k);
              }
            }
            s_colsum[ab][quad][c0 + lane] = t[0];
          }
          if (g.out_f32) {
            float* dst = g.out_f32 + (tile * 128 + row) * g.ldc;
#pragma unroll
            for (int i = 0; i < 32; ++i) if (c0 + i < g.n_valid) dst[c0 + i] = v[i];
          }
          if (g.out_hi) {
            const long base = ((tile * (g.out_feat / 8) + (g.out_off + c0) / 8) * 128 + row) * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float s[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) s[j] = v[q * 8 + j] * g.out_scale;
              uint4 h;
              h.x = pack_h2(s[0], s[1]); h.y = pack_h2(s[2], s[3]); h.z = pack_h2(s[4], s[5]); h.w = pack_h2(s[6], s[7]);
              *reinterpret_cast<uint4*>((char*)g.out_hi + base + (long)q * 2048) = h;
              if (g.out_lo) {
                const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
                float l[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  l[2 * j] = s[2 * j] - __half2float(__ushort_as_half((unsigned short)(hw[j] & 0xffffu)));
                  l[2 * j + 1] = s[2 * j + 1] - __half2float(__ushort_as_half((unsigned short)(hw[j] >> 16)));
                }
                uint4 lo;
                lo.x = pack_h2(l[0], l[1]); lo.y = pack_h2(l[2], l[3]); lo.z = pack_h2(l[4], l[5]); lo.w = pack_h2(l[6], l[7]);
                *reinterpret_cast<uint4*>((char*)g.out_lo + base + (long)q * 2048) = lo;
              }
            }
          }
        }
      }
      if (g.colsum) {
        // the eight epilogue warps meet once per tile; s_colsum is double buffered by tile parity, so a warp that runs ahead
        // into the next tile writes the other half
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // one partial row per CTA (not per tile): the column sums of this CTA's tiles accumulate in a register, in tile order
        if (e < N) cs_acc += (s_colsum[ab][0][e] + s_colsum[ab][1][e]) + (s_colsum[ab][2][e] + s_colsum[ab][3][e]);
      }
    }
    if (g.colsum && e < N) g.colsum[(long)blockIdx.x * N + e] = cs_acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ---- packing ------------------------------------------------------------------------------------------------------------
// fp32 [rows_in, C] (row stride ld; source row of packed row m is m / row_div) -> PK(m_tiles*128, c_pad) hi (+ lo) * scale;
// rows >= M and columns >= C are zero.
__global__ void pack_rows_kernel(const float* __restrict__ src, long ld, int C, long M, int row_div, int m_tiles, int c_pad,
                                 float scale, uint4* __restrict__ hi, uint4* __restrict__ lo) {
  const long total = (long)m_tiles * (c_pad / 8) * 128;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int row = (int)(idx & 127);
    const long rest = idx >> 7;
    const int kg = (int)(rest % (c_pad / 8));
    const long tile = rest / (c_pad / 8);
    const long m = tile * 128 + row;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = kg * 8 + j;
      s[j] = (m < M && c < C) ? src[(m / row_div) * ld + c] * scale : 0.f;
    }
    uint4 h;
    h.x = pack_h2(s[0], s[1]); h.y = pack_h2(s[2], s[3]); h.z = pack_h2(s[4], s[5]); h.w = pack_h2(s[6], s[7]);
    hi[idx] = h;
    if (lo) {
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
      float l[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        l[2 * j] = s[2 * j] - __half2float(__ushort_as_half((unsigned short)(hw[j] & 0xffffu)));
        l[2 * j + 1] = s[2 * j + 1] - __half2float(__ushort_as_half((unsigned short)(hw[j] >> 16)));
      }
      uint4 q;
      q.x = pack_h2(l[0], l[1]); q.y = pack_h2(l[2], l[3]); q.z = pack_h2(l[4], l[5]); q.w = pack_h2(l[6], l[7]);
      lo[idx] = q;
    }
  }
}

// the same into the tile order of the fused training forward (render_tc.cu, TrainDump): packed row (tile rt * S + s, r) reads
// source row (rt*128 + r) * S + s, or rt*128 + r when the source is per ray; rows of rays >= R are zero
__global__ void pack_rows_tiled_kernel(const float* __restrict__ src, long ld, int C, int R, int S, int per_ray, long m_tiles, int c_pad,
                                       float scale, uint4* __restrict__ hi, uint4* __restrict__ lo) {
  const long total = m_tiles * (c_pad / 8) * 128;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int row = (int)(idx & 127);
    const long rest = idx >> 7;
    const int kg = (int)(rest % (c_pad / 8));
    const long tile = rest / (c_pad / 8);
    const long ray = (tile / S) * 128 + row;
    const long m = per_ray ? ray : ray * S + (tile % S);
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = kg * 8 + j;
      s[j] = (ray < R && c < C) ? src[m * ld + c] * scale : 0.f;
    }
    uint4 h;
    h.x = pack_h2(s[0], s[1]); h.y = pack_h2(s[2], s[3]); h.z = pack_h2(s[4], s[5]); h.w = pack_h2(s[6], s[7]);
    hi[idx] = h;
    if (lo) {
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
      float l[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        l[2 * j] = s[2 * j] - __half2float(__ushort_as_half((unsigned short)(hw[j] & 0xffffu)));
        l[2 * j + 1] = s[2 * j + 1] - __half2float(__ushort_as_half((unsigned short)(hw[j] >> 16)));
      }
      uint4 q;
      q.x = pack_h2(l[0], l[1]); q.y = pack_h2(l[2], l[3]); q.z = pack_h2(l[4], l[5]); q.w = pack_h2(l[6], l[7]);
      lo[idx] = q;
    }
  }
}

__global__ void unpack_rows_tiled_kernel(const float* __restrict__ src, long ld, int C, int R, int S, float* __restrict__ dst) {
  const long total = (long)R * S * C;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long m = idx / C;
    const long ray = m / S, s = m % S;
    dst[idx] = src[(((ray >> 7) * S + s) * 128 + (ray & 127)) * ld + c];
  }
}

// nn.Linear weight W [out, in] fp32 -> PW(r_pad, k_pad) = [k_pad/8][r_pad][8] hi (+ lo) * scale;
// transpose = 0: rows = out features, contraction = in features (forward); 1: rows = in, contraction = out (dgrad).
__global__ void pack_linear_kernel(const float* __restrict__ W, int out_f, int in_f, int transpose, int r_pad, int k_pad, float scale,
                                   uint4* __restrict__ hi, uint4* __restrict__ lo) {
  const int total = (k_pad / 8) * r_pad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx % r_pad, kg = idx / r_pad;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kg * 8 + j;
      const int o = transpose ? k : r, i = transpose ? r : k;
      s[j] = (o < out_f && i < in_f) ? W[(long)o * in_f + i] * scale : 0.f;
    }
    uint4 h;
    h.x = pack_h2(s[0], s[1]); h.y = pack_h2(s[2], s[3]); h.z = pack_h2(s[4], s[5]); h.w = pack_h2(s[6], s[7]);
    hi[idx] = h;
    if (lo) {
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
      float l[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        l[2 * j] = s[2 * j] - __half2float(__ushort_as_half((unsigned short)(hw[j] & 0xffffu)));
        l[2 * j + 1] = s[2 * j + 1] - __half2float(__ushort_as_half((unsigned short)(hw[j] >> 16)));
      }
      uint4 q;
      q.x = pack_h2(l[0], l[1]); q.y = pack_h2(l[2], l[3]); q.z = pack_h2(l[4], l[5]); q.w = pack_h2(l[6], l[7]);
      lo[idx] = q;
    }
  }
}

// dst[r, col_off + c] (or dst[c, col_off + r] if transpose) = scale * sum_split partial[split][r][c], fixed summation order.
// A block owns 64 consecutive outputs; its four 64-thread groups each sum a contiguous quarter of the splits (four independent
// accumulators per thread keep 4 loads in flight), and the quarters are combined in a fixed order through shared memory: four
// times the memory-level parallelism of one thread per output (148 splits x 256 KB = 39 MB per 256 x 256 layer), same result on
// every run.
constexpr int WR_OUT = 64, WR_GROUPS = 4;
__global__ void __launch_bounds__(WR_OUT * WR_GROUPS)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int rows_pad, int N, float scale,
                    float* __restrict__ dst, long ld, int col_off, int rows_valid, int cols_valid, int flags) {
  __shared__ float s_part[WR_GROUPS][WR_OUT];
  const int total = rows_pad * N;
  const bool transpose = flags & 1, accumulate = flags & 2;   // accumulate: dst += (row-tile sub-batches of one backward pass)
  const int tx = threadIdx.x & (WR_OUT - 1), grp = threadIdx.x / WR_OUT;
  const int idx = blockIdx.x * WR_OUT + tx;
  const int per = (splits + WR_GROUPS - 1) / WR_GROUPS;
  const int k0 = grp * per, k1 = min(splits, k0 + per);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (idx < total) {
    const float* p = partial + idx;
    int k = k0;
    for (; k + 3 < k1; k += 4) {
      a0 += p[(long)k * total]; a1 += p[(long)(k + 1) * total]; a2 += p[(long)(k + 2) * total]; a3 += p[(long)(k + 3) * total];
    }
    for (; k < k1; ++k) a0 += p[(long)k * total];
  }
  s_part[grp][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (grp != 0 || idx >= total) return;
  const int r = idx / N, c = idx % N;
  if (r >= rows_valid || c >= cols_valid) return;
  const float s = (s_part[0][tx] + s_part[1][tx]) + (s_part[2][tx] + s_part[3][tx]);
  float* q = transpose ? dst + (long)c * ld + col_off + r : dst + (long)r * ld + col_off + c;
  *q = accumulate ? *q + s * scale : s * scale;
}

// bias gradient from the partial column sums of a dgrad GEMM: dst[c] = scale * sum_r partial[r][c], rows summed in order
// (one thread per column, four accumulators: the rows are few -- one per CTA of the persistent kernel -- and coalesced)
__global__ void colsum_finish_kernel(const float* __restrict__ partial, int rows, int N, float scale, float* __restrict__ dst) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int r = 0;
  for (; r + 3 < rows; r += 4) {
    a0 += partial[(long)r * N + c]; a1 += partial[(long)(r + 1) * N + c]; a2 += partial[(long)(r + 2) * N + c]; a3 += partial[(long)(r + 3) * N + c];
  }
  for (; r < rows; ++r) a0 += partial[(long)r * N + c];
  dst[c] = ((a0 + a1) + (a2 + a3)) * scale;
}

// Column sums of a PK(rows, feat) tensor (hi + lo planes): partial[split][feat] = sum over the split's row tiles.
__global__ void __launch_bounds__(128) colsum_packed_kernel(const uint4* __restrict__ hi, const uint4* __restrict__ lo, int feat,
                                                             int m_tiles, int tiles_per_split, float* __restrict__ partial) {
  __shared__ float red[4][8];
  const int kg = blockIdx.x, split = blockIdx.y, row = threadIdx.x;
  const int t0 = split * tiles_per_split, t1 = min(m_tiles, t0 + tiles_per_split);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int t = t0; t < t1; ++t) {
    const long idx = ((long)t * (feat / 8) + kg) * 128 + row;
    const uint4 h = hi[idx];
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[2 * j] += __half2float(__ushort_as_half((unsigned short)(hw[j] & 0xffffu)));
      s[2 * j + 1] += __half2float(__ushort_as_half((unsigned short)(hw[j] >> 16)));
    }
    if (lo) {
      const uint4 l = lo[idx];
      const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[2 * j] += __half2float(__ushort_as_half((unsigned short)(lw[j] & 0xffffu)));
        s[2 * j + 1] += __half2float(__ushort_as_half((unsigned short)(lw[j] >> 16)));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
  if ((row & 31) == 0)
    for (int j = 0; j < 8; ++j) red[row >> 5][j] = s[j];
  __syncthreads();
  if (row < 8) partial[(long)split * feat + kg * 8 + row] = (red[0][row] + red[1][row]) + (red[2][row] + red[3][row]);
}

}  // namespace aon

using namespace aon;

// ---- tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency) ------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (TmapEncodeFn)p;
  });
  return fn;
}
// Encoded maps are pure functions of (pointer, shape, box): cached per thread ("lazily-created CUtensorMaps cached
// per (ptr, shape)"), so a training step that re-uses its planes pays the ~1 us encode once.
typedef std::tuple<const void*, int, int, int, int, int> TmapKey;
static int tmap_get(const TmapKey& key, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                    CUtensorMap* out) {
  thread_local std::map<TmapKey, CUtensorMap> cache;
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return AON_OK; }
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return AON_E_UNSUPPORTED; }
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank, const_cast<void*>(std::get<0>(key)), dims, strides, box,
                         ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return AON_E_CUDA; }
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return AON_OK;
}
// PK(rows, feat) = [tiles][feat/8][128][8] x 16 bit, in 8-byte elements: box = GT_KC samples x `groups` feature groups of one tile
static int tmap_packed(const void* base, int feat, int tiles, int groups, CUtensorMap* out) {
  const cuuint64_t dims[3] = {256, (cuuint64_t)(feat / 8), (cuuint64_t)tiles};
  const cuuint64_t strides[2] = {2048, (cuuint64_t)(feat / 8) * 2048};
  const cuuint32_t box[3] = {GT_KC * 2, (cuuint32_t)groups, 1};
  return tmap_get(TmapKey(base, 3, feat, tiles, groups, 0), 3, dims, strides, box, out);
}

extern "C" size_t aon_gemm_struct_size(void) { return sizeof(AonGemm); }

static bool nt_persistent(const AonGemm& g, int sms) {
  return g.mode == AON_GEMM_NT && (g.N == 128 || g.N == 256) && g.m_tiles >= 2 * sms && !(g.reserved & 16);
}

extern "C" int aon_gemm_tc(const AonGemm* gp, aon_stream_t stream) {
  AON_REQUIRE(gp != nullptr, "aon_gemm_tc: null descriptor");
  const AonGemm& g = *gp;
  AON_REQUIRE(g.mode == AON_GEMM_NT || g.mode == AON_GEMM_TN, "aon_gemm_tc: bad mode %d", g.mode);
  AON_REQUIRE(g.N >= 16 && g.N <= 256 && g.N % 16 == 0, "aon_gemm_tc: N = %d must be a multiple of 16 in [16, 256]", g.N);
  AON_REQUIRE(g.m_tiles >= 0, "aon_gemm_tc: bad m_tiles");
  AON_REQUIRE(g.nseg >= 1 && g.nseg <= AON_GEMM_MAX_SEG, "aon_gemm_tc: bad nseg %d", g.nseg);
  for (int s = 0; s < g.nseg; ++s) {
    AON_REQUIRE(g.a_hi[s] && g.b_hi[s] && (!g.x3 || (g.a_lo[s] && g.b_lo[s])), "aon_gemm_tc: null operand plane (segment %d)", s);
    AON_REQUIRE(g.a_feat[s] % 8 == 0 && g.b_feat[s] % 8 == 0 && g.a_off[s] % 8 == 0 && g.b_off[s] % 8 == 0,
                "aon_gemm_tc: feature counts / offsets must be multiples of 8 (segment %d)", s);
    if (g.mode == AON_GEMM_NT) AON_REQUIRE(g.kext[s] > 0 && g.kext[s] % 16 == 0, "aon_gemm_tc: kext must be a positive multiple of 16");
  }
  if (g.m_tiles == 0) return AON_OK;
  int dev = 0, major = 0;
  AON_CUDA_CHECK(cudaGetDevice(&dev));
  AON_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("aon_gemm_tc needs an sm_100 device (found compute capability %d.x)", major);
    return AON_E_UNSUPPORTED;
  }
  dim3 grid;
  if (g.mode == AON_GEMM_NT) {
    AON_REQUIRE(g.epi == AON_GEMM_EPI_LINEAR || g.epi == AON_GEMM_EPI_MASK, "aon_gemm_tc: NT mode takes a LINEAR or MASK epilogue");
    AON_REQUIRE(g.colsum == nullptr || g.N % 32 == 0, "aon_gemm_tc: colsum needs N to be a multiple of 32");
    AON_REQUIRE((g.relu_bits_out == nullptr && g.mask_bits == nullptr) || g.N % 32 == 0, "aon_gemm_tc: bit-plane masks need N % 32 == 0");
    grid = dim3((unsigned)g.m_tiles);
  } else {
    AON_REQUIRE(g.epi == AON_GEMM_EPI_PARTIAL && g.partial != nullptr && g.nseg == 1, "aon_gemm_tc: TN mode writes partial tiles");
    AON_REQUIRE(g.a_tiles >= 1 && g.splits >= 1 && g.tiles_per_split >= 1 && (long)g.splits * g.tiles_per_split >= g.m_tiles,
                "aon_gemm_tc: bad TN decomposition");
    grid = dim3((unsigned)g.a_tiles, (unsigned)g.splits);
  }
  GtParams P;
  memset(&P, 0, sizeof(P));
  P.g = g;
  int rc;
  if (g.mode == AON_GEMM_TN) {
    AON_REQUIRE(g.a_off[0] + g.a_tiles * 128 <= g.a_feat[0] && g.b_off[0] + g.N <= g.b_feat[0], "aon_gemm_tc: TN operands out of range");
    if ((rc = tmap_packed(g.a_hi[0], g.a_feat[0], g.m_tiles, 16, &P.tm[0])) != AON_OK) return rc;
    if ((rc = tmap_packed(g.b_hi[0], g.b_feat[0], g.m_tiles, g.N / 8, &P.tm[2])) != AON_OK) return rc;
    if (g.x3) {
      if ((rc = tmap_packed(g.a_lo[0], g.a_feat[0], g.m_tiles, 16, &P.tm[1])) != AON_OK) return rc;
      if ((rc = tmap_packed(g.b_lo[0], g.b_feat[0], g.m_tiles, g.N / 8, &P.tm[3])) != AON_OK) return rc;
    }
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (nt_persistent(g, sms)) {
    // the large forward / dgrad GEMMs: persistent kernel, one CTA per SM
    AON_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_nt_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GP_SMEM));
    gemm_tc_nt_persistent_kernel<<<sms, GP_THREADS, GP_SMEM, (cudaStream_t)stream>>>(P);
    AON_LAUNCH_CHECK();
    return AON_OK;
  }
  AON_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
  gemm_tc_kernel<<<grid, GT_THREADS, GT_SMEM, (cudaStream_t)stream>>>(P);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

// rows of the colsum buffer an NT launch writes: one per row tile, or one per CTA when the persistent kernel takes the GEMM
extern "C" int aon_gemm_colsum_rows(const AonGemm* gp) {
  if (!gp) return 0;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return nt_persistent(*gp, sms) ? sms : gp->m_tiles;
}

extern "C" int aon_pack_rows(const float* src, long ld, int C, long M, int row_div, int m_tiles, int c_pad, float scale, void* hi,
                             void* lo, aon_stream_t stream) {
  AON_REQUIRE(src && hi, "aon_pack_rows: null pointer");
  AON_REQUIRE(C >= 1 && c_pad % 8 == 0 && c_pad >= C && row_div >= 1 && m_tiles >= 0 && M <= (long)m_tiles * 128,
              "aon_pack_rows: bad sizes C=%d c_pad=%d M=%ld m_tiles=%d", C, c_pad, M, m_tiles);
  if (m_tiles == 0) return AON_OK;
  const long total = (long)m_tiles * (c_pad / 8) * 128;
  const long blocks = (total + 255) / 256;
  pack_rows_kernel<<<(unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks), 256, 0, (cudaStream_t)stream>>>(src, ld, C, M, row_div, m_tiles, c_pad, scale,
                                                                                                   (uint4*)hi, (uint4*)lo);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_pack_rows_tiled(const float* src, long ld, int C, int R, int S, int per_ray, int c_pad, float scale, void* hi, void* lo,
                                   aon_stream_t stream) {
  AON_REQUIRE(src && hi && C >= 1 && R >= 1 && S >= 1 && c_pad >= C && c_pad % 8 == 0 && ld >= C, "aon_pack_rows_tiled: bad arguments");
  const long m_tiles = (long)((R + 255) / 256) * 2 * S;
  const long total = m_tiles * (c_pad / 8) * 128;
  const long blocks = (total + 255) / 256;
  pack_rows_tiled_kernel<<<(unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks), 256, 0, (cudaStream_t)stream>>>(src, ld, C, R, S, per_ray, m_tiles, c_pad,
                                                                                                         scale, (uint4*)hi, (uint4*)lo);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_unpack_rows_tiled(const float* src, long ld, int C, int R, int S, float* dst, aon_stream_t stream) {
  AON_REQUIRE(src && dst && C >= 1 && R >= 1 && S >= 1 && ld >= C, "aon_unpack_rows_tiled: bad arguments");
  const long total = (long)R * S * C;
  const long blocks = (total + 255) / 256;
  unpack_rows_tiled_kernel<<<(unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks), 256, 0, (cudaStream_t)stream>>>(src, ld, C, R, S, dst);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_pack_linear(const float* W, int out_features, int in_features, int transpose, int r_pad, int k_pad, float scale,
                               void* hi, void* lo, aon_stream_t stream) {
  AON_REQUIRE(W && hi, "aon_pack_linear: null pointer");
  const int rows = transpose ? in_features : out_features, kk = transpose ? out_features : in_features;
  AON_REQUIRE(r_pad >= rows && k_pad >= kk && k_pad % 8 == 0 && r_pad % 8 == 0, "aon_pack_linear: bad padding r_pad=%d k_pad=%d", r_pad, k_pad);
  const int total = (k_pad / 8) * r_pad;
  pack_linear_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, out_features, in_features, transpose, r_pad, k_pad, scale,
                                                                            (uint4*)hi, (uint4*)lo);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_wgrad_reduce(const float* partial, int splits, int rows_pad, int N, float scale, float* dst, long ld, int col_off,
                                int rows_valid, int cols_valid, int transpose, aon_stream_t stream) {
  AON_REQUIRE(partial && dst && splits >= 1 && rows_pad >= 1 && N >= 1, "aon_wgrad_reduce: bad arguments");
  const int total = rows_pad * N;
  wgrad_reduce_kernel<<<(total + WR_OUT - 1) / WR_OUT, WR_OUT * WR_GROUPS, 0, (cudaStream_t)stream>>>(partial, splits, rows_pad, N, scale, dst, ld, col_off,
                                                                             rows_valid, cols_valid, transpose);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_colsum_finish(const float* partial, int rows, int N, float scale, float* dst, aon_stream_t stream) {
  AON_REQUIRE(partial && dst && rows >= 1 && N >= 1, "aon_colsum_finish: bad arguments");
  colsum_finish_kernel<<<(N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(partial, rows, N, scale, dst);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_colsum_packed(const void* hi, const void* lo, int feat, int m_tiles, int splits, float* partial, aon_stream_t stream) {
  AON_REQUIRE(hi && partial && feat % 8 == 0 && feat >= 8 && splits >= 1 && m_tiles >= 0, "aon_colsum_packed: bad arguments");
  const int tps = (m_tiles + splits - 1) / splits;
  colsum_packed_kernel<<<dim3(feat / 8, splits), 128, 0, (cudaStream_t)stream>>>((const uint4*)hi, (const uint4*)lo, feat, m_tiles,
                                                                                 tps > 0 ? tps : 1, partial);
  AON_LAUNCH_CHECK();
  return AON_OK;
}
