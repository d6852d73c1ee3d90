// render_tc.cu -- tensor-core (tcgen05 / TMEM) version of the fused per-level render kernel.
//
// One CTA renders 128 rays.  MMA row m = ray m of the tile, and the CTA walks the S samples of its
// rays one 128-row "sample plane" at a time, so a ray's transmittance / colour / depth accumulators
// are running scalars in one epilogue thread and no [rays x samples x features] tensor ever reaches
// HBM (reference: helper.py:25-26,136-140,157-195; model.py:95-120,174-195;
// model_autodecoder.py:171-239,306-331).
//
// Work decomposition.  Every nn.Linear with >= 128 outputs is cut into "units": 128 output features
// x all K.  K is cut into 32-wide "chunks"; a chunk is either 32 hidden features written by the
// epilogue of the previous layer (ids 0..7), one half of the 64-wide positional encoding (8, 9), the
// 32-wide view-direction encoding (10) or the raw sample position of the deformation MLP (11).
//   warp 0   producer: streams the pre-packed weight stream (8 KB stages, already in the UMMA
//            canonical K-major layout) from L2 into a shared-memory ring with 1-D bulk copies
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M=128, N=128, K=16) into one of four
//            128-column TMEM accumulators; a K chunk is issued as soon as the epilogue has published
//            it, so layer l+1 starts while the second half of layer l is still being drained
//   warps 4-7 epilogue: tcgen05.ld the accumulators (one TMEM lane = one ray = one thread), add
//            bias, ReLU, convert to the 16-bit operand format and store the next layer's A operand
//            chunk to shared memory; the 1-/3-wide heads (density, rgb, deformation) are fp32 FMAs
//            on the fp32 accumulators; then activations + alpha compositing in registers.
// Precision modes: AON_PREC_TC_F16 / _BF16 = one MMA per K step; AON_PREC_TC_F16X3 = operands split
// into fp16 hi + fp16 lo (about 22 significand bits), three MMAs per K step
// (hi*hi + lo*hi + hi*lo) with fp32 accumulation -- the mode that meets the 1e-4 parity bar.
#include <string.h>

#include "aon_common.cuh"
#include "tc_ptx.cuh"

namespace aon {

void layout_tail(int kind, PackedLayout& L, int64_t off);  // aon_api.cu

// ---- program (unit schedule), built on the host, passed to the kernels by value ------------------
constexpr int MAX_UNITS = 28;
constexpr int CH_E0 = 8, CH_V = 10, CH_P = 11, NUM_CHUNK_IDS = 12;
constexpr int STAGE_BYTES = 8192;
enum Epi : int { EPI_STORE = 0, EPI_STORE_SIGMA = 1, EPI_RGB = 2, EPI_DEFORM = 3 };

struct Unit {
  uint8_t gemm, half, n_chunks, epi;
  uint8_t relu, in_buf, out_buf, wait_next;
  uint8_t chunk[12];
  uint8_t gen[12];
  uint16_t bias_off;
  uint16_t stage0;  // first weight stage of this unit within a sample
};

struct Program {
  int n_units, n_stages, n_bias, x3;
  int n_gemm;
  int wps[NUM_CHUNK_IDS];  // writes per sample of each chunk id
  uint16_t layer_n[MAX_GEMM];  // out features of each GEMM layer (bias vector lengths)
  Unit u[MAX_UNITS];
};

static int stages_of_chunk(int id, int x3) { return x3 ? (id == CH_P ? 1 : 2) : 1; }

static Program build_program(int kind, int precision) {
  Program P;
  memset(&P, 0, sizeof(P));
  const int x3 = precision == AON_PREC_TC_F16X3;
  P.x3 = x3;
  const GemmLayer* g = gemm_layers(kind);
  const int ng = num_gemm(kind);
  int count[NUM_CHUNK_IDS] = {0};
  if (kind == AON_KIND_VANILLA) count[CH_E0] = count[CH_E0 + 1] = 1;
  else count[CH_P] = 1;
  count[CH_V] = 1;
  int bias_pref = 0, stage = 0, nu = 0;
  for (int gi = 0; gi < ng; ++gi) {
    const int halves = g[gi].N / 128;
    int epi = EPI_STORE;
    if (kind == AON_KIND_VANILLA) {
      if (gi == 7) epi = EPI_STORE_SIGMA;
      if (gi == 9) epi = EPI_RGB;
    } else {
      if (gi == 3) epi = EPI_DEFORM;
      if (gi == 11) epi = EPI_STORE_SIGMA;
      if (gi == 16) epi = EPI_RGB;
    }
    for (int h = 0; h < halves; ++h) {
      Unit& u = P.u[nu++];
      u.gemm = gi; u.half = h; u.epi = epi; u.relu = g[gi].relu;
      u.out_buf = x3 ? 0 : (gi & 1);
      u.in_buf = x3 ? 0 : ((gi + 1) & 1);
      u.wait_next = (x3 && halves == 2 && h == 0 && epi <= EPI_STORE_SIGMA) ? 1 : 0;
      u.bias_off = (uint16_t)(bias_pref + h * 128);
      u.stage0 = (uint16_t)stage;
      int nc = 0;
      for (int j = 0; j < g[gi].K1 / 32; ++j) u.chunk[nc++] = j;
      if (g[gi].aux == AUX_E) { u.chunk[nc++] = CH_E0; u.chunk[nc++] = CH_E0 + 1; }
      if (g[gi].aux == AUX_V) u.chunk[nc++] = CH_V;
      if (g[gi].aux == AUX_P) u.chunk[nc++] = CH_P;
      u.n_chunks = nc;
      for (int c = 0; c < nc; ++c) {
        u.gen[c] = (uint8_t)(count[u.chunk[c]] - 1);
        stage += stages_of_chunk(u.chunk[c], x3);
      }
    }
    // after both halves of the layer: its outputs become new generations of the A chunks
    if (epi <= EPI_STORE_SIGMA)
      for (int j = 0; j < halves * 4; ++j) count[j]++;
    if (epi == EPI_DEFORM) { count[CH_E0]++; count[CH_E0 + 1]++; }
    bias_pref += g[gi].N;
  }
  // generations consumed by a unit's first half must not see the layer's own writes: the loop above
  // bumps the counts only after both halves, which is exactly that.
  P.n_units = nu;
  P.n_gemm = ng;
  for (int gi = 0; gi < ng; ++gi) P.layer_n[gi] = (uint16_t)g[gi].N;
  P.n_stages = stage;
  P.n_bias = bias_pref;
  for (int i = 0; i < NUM_CHUNK_IDS; ++i) P.wps[i] = count[i];
  P.wps[CH_V] = 0;  // written once per CTA
  return P;
}

PackedLayout layout_tc(int kind, int precision) {
  PackedLayout L;
  memset(&L, 0, sizeof(L));
  const Program P = build_program(kind, precision);
  for (int i = 0; i < num_gemm(kind); ++i) L.w[i] = 0;  // one contiguous stream, see Program::stage0
  layout_tail(kind, L, (int64_t)P.n_stages * STAGE_BYTES);
  return L;
}

// ---- weight stream packing ----------------------------------------------------------------------------
// Stage layout (8 KB, exactly what tcgen05.mma reads as its B operand, K-major, no swizzle):
//   one pass : [4 k-groups][128 n][8 k] 16-bit                    (K = 32: two K=16 steps)
//   x3       : hi [2 k-groups][128 n][8 k] fp16, then lo likewise (K = 16: one step, hi and lo parts)
struct PackSrc {
  const float* w[20];
  GemmLayer g[MAX_GEMM];
  int in_features[MAX_GEMM];
};

template <bool X3, bool BF16>
__global__ void pack_stream_kernel(Program P, PackSrc src, int kind, uint16_t* __restrict__ out) {
  const long total = (long)P.n_stages * (STAGE_BYTES / 2);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int stage = (int)(idx / (STAGE_BYTES / 2));
    const int e = (int)(idx % (STAGE_BYTES / 2));
    int ui = 0;
    while (ui + 1 < P.n_units && P.u[ui + 1].stage0 <= stage) ++ui;
    const Unit& u = P.u[ui];
    // locate chunk and sub-stage
    int rel = stage - u.stage0, c = 0, sub = 0;
    for (;; ++c) {
      const int ns = X3 ? (u.chunk[c] == CH_P ? 1 : 2) : 1;
      if (rel < ns) { sub = rel; break; }
      rel -= ns;
    }
    const int id = u.chunk[c];
    int part = 0, kg, n, kk, kin;
    if (X3) {
      part = e / 2048;
      const int e2 = e % 2048;
      kg = e2 / 1024; n = (e2 % 1024) / 8; kk = e2 % 8;
      kin = sub * 16 + kg * 8 + kk;
    } else {
      kg = e / 1024; n = (e % 1024) / 8; kk = e % 8;
      kin = kg * 8 + kk;
    }
    const GemmLayer& g = src.g[u.gemm];
    const int in_features = src.in_features[u.gemm];
    int col = -1;
    if (id < 8) col = id * 32 + kin;
    else {
      const int a = (id == CH_E0 + 1 ? 32 : 0) + kin;
      if (a < g.aux_cnt) col = g.aux_col0 + a;
    }
    float v = 0.f;
    if (col >= 0) v = src.w[g.src][(size_t)(u.half * 128 + n) * in_features + col];
    uint16_t bits;
    if (BF16) {
      bits = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    } else {
      const __half hi = __float2half_rn(v);
      if (X3 && part == 1) bits = __half_as_ushort(__float2half_rn(v - __half2float(hi)));
      else bits = __half_as_ushort(hi);
    }
    out[idx] = bits;
  }
}

// ---- render kernel -----------------------------------------------------------------------------------------
constexpr int TC_THREADS = 256;
constexpr int SMEM_MAX = 232448;

template <int KIND, bool X3>
struct SmemPlan {
  static constexpr int A = 0;                          // 1-pass: two 64 KB buffers; x3: hi | lo
  static constexpr int A_LO = 65536;
  static constexpr int E = 131072;                     // 64-wide encoding (two chunks)
  static constexpr int E_LO = 16384;
  static constexpr int V = E + (X3 ? 32768 : 16384);   // 32-wide view encoding
  static constexpr int V_LO = 8192;
  static constexpr int P = V + (X3 ? 16384 : 8192);    // raw position (auto-decoder only)
  static constexpr int P_LO = 4096;
  static constexpr int P_BYTES = KIND == AON_KIND_AUTODECODER ? 8192 : 0;
  static constexpr int PARAMS = P + P_BYTES;           // fp32 biases + head weights
  static constexpr int PARAM_FLOATS = KIND == AON_KIND_AUTODECODER ? 4400 : 3200;
  static constexpr int BARS = PARAMS + PARAM_FLOATS * 4;
  static constexpr int BAR_BYTES = 512;
  static constexpr int RING = (BARS + BAR_BYTES + 127) / 128 * 128;
  static constexpr int NSTAGE_RAW = (SMEM_MAX - 1024 - RING) / STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
  static constexpr int TOTAL = RING + NSTAGE * STAGE_BYTES + 1024;
  static_assert(NSTAGE >= 3, "weight ring too small");
};

struct TcParams {
  Program prog;
  PackedLayout L;
  const char* packed;
  const float* folded;
  const float* rays_o;
  const float* rays_d;
  const float* viewdirs;
  const float* t_vals;
  long t_stride;
  int R, S, white_bkgd;
  float* comp_rgb;
  float* acc;
  float* depth;
  float* weights;
  float* dbg;        // optional [n_units][128][128] pre-activation dump of tile 0 / sample 0
  int* err_flag;     // optional: set to a non-zero code when a barrier wait times out
};

// barrier slots (8 bytes each) inside the BARS region
constexpr int BAR_FULL = 0, BAR_EMPTY = 8, BAR_DFULL = 16, BAR_DEMPTY = 20, BAR_CHUNK = 24, BAR_TMEM = 40;

__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s: a schedule bug, not a slow kernel
      if (err_flag) atomicExch(err_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}

template <int L>
__device__ __forceinline__ float enc_val(int k, float x, float y, float z) {
  // helper.py:136-140: [x y z | sin(2^f v_d), f-major | sin(2^f v_d + pi/2), f-major], then zero padding
  if (k < 3) return k == 0 ? x : (k == 1 ? y : z);
  if (k >= 3 + 6 * L) return 0.f;
  const bool shifted = k >= 3 + 3 * L;
  const int j = k - 3 - (shifted ? 3 * L : 0);
  const int f = j / 3, d = j - 3 * f;
  const float v = d == 0 ? x : (d == 1 ? y : z);
  const float xb = v * (float)(1 << f);  // exact (power of two)
  return sinf(shifted ? __fadd_rn(xb, AON_HALF_PI_F) : xb);
}

template <bool X3, bool BF16>
__device__ __forceinline__ void pack8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (BF16) {
      __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      h[i] = *reinterpret_cast<uint32_t*>(&b);
    } else {
      __half2 a = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      h[i] = *reinterpret_cast<uint32_t*>(&a);
      if (X3) {
        const float2 f = __half22float2(a);
        __half2 r = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
        l[i] = *reinterpret_cast<uint32_t*>(&r);
      }
    }
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// writes NK (multiple of 8) encoded values of one row into an operand region: element k of row r at
// base + (k/8)*2048 + r*16 + (k%8)*2
template <int L, int NK, bool X3, bool BF16>
__device__ __forceinline__ void encode_store(unsigned char* base, int lo_delta, int row, float x, float y, float z) {
#pragma unroll 1
  for (int kg = 0; kg < NK / 8; ++kg) {
    float v[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) v[kk] = enc_val<L>(kg * 8 + kk, x, y, z);
    uint4 hi, lo;
    pack8<X3, BF16>(v, hi, lo);
    *reinterpret_cast<uint4*>(base + kg * 2048 + row * 16) = hi;
    if (X3) *reinterpret_cast<uint4*>(base + lo_delta + kg * 2048 + row * 16) = lo;
  }
}

__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
  // tcgen05.wait::ld with the destination registers as in/out operands so that no use of them can be
  // scheduled above the wait
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

template <int KIND, bool X3, bool BF16>
__global__ void __launch_bounds__(TC_THREADS, 1) render_tc_kernel(const __grid_constant__ TcParams p) {
  using SP = SmemPlan<KIND, X3>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sm_u32 = ptx::smem_u32(sm);
  float* s_par = reinterpret_cast<float*>(sm + SP::PARAMS);
  const uint32_t bars = sm_u32 + SP::BARS;
  auto bar = [&](int slot) { return bars + 8u * slot; };
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(sm + SP::BARS + 8 * BAR_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Program& P = p.prog;
  const int S = p.S;
  constexpr int NSTAGE = SP::NSTAGE;

  // ---- one-time setup ---------------------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { ptx::mbar_init(bar(BAR_FULL + i), 1); ptx::mbar_init(bar(BAR_EMPTY + i), 1); }
    for (int i = 0; i < 4; ++i) { ptx::mbar_init(bar(BAR_DFULL + i), 1); ptx::mbar_init(bar(BAR_DEMPTY + i), 4); }
    for (int i = 0; i < NUM_CHUNK_IDS; ++i) ptx::mbar_init(bar(BAR_CHUNK + i), 4);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(sm_u32 + SP::BARS + 8 * BAR_TMEM, 512);
    ptx::tmem_relinquish();
  }
  // biases (folded ones for the latent-conditioned layers) and head weights -> shared memory
  {
    int off = 0;
    for (int gi = 0; gi < P.n_gemm; ++gi) {
      const int N = P.layer_n[gi];
      const float* src = (KIND == AON_KIND_AUTODECODER && p.L.fold[gi] >= 0)
                             ? p.folded + p.L.fold[gi]
                             : reinterpret_cast<const float*>(p.packed + p.L.bias[gi]);
      for (int i = tid; i < N; i += TC_THREADS) s_par[off + i] = src[i];
      off += N;
    }
    // heads: [deform 3x128 + 4,] density 1x256 + 4, rgb 3x128 + 4   (same order as PackedLayout::head_*)
    constexpr int NH = KIND == AON_KIND_VANILLA ? 2 : 3;
    for (int h = 0; h < NH; ++h) {
      const int n = (KIND == AON_KIND_AUTODECODER ? (h == 1 ? 256 : 384) : (h == 0 ? 256 : 384));
      const float* w = reinterpret_cast<const float*>(p.packed + p.L.head_w[h]);
      const float* b = reinterpret_cast<const float*>(p.packed + p.L.head_b[h]);
      for (int i = tid; i < n; i += TC_THREADS) s_par[off + i] = w[i];
      if (tid < 4) s_par[off + n + tid] = b[tid];
      off += n + 4;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ================================ weight producer ================================
    if (lane == 0) {
      const long total = (long)S * P.n_stages;
      int srcs = 0;
      for (long i = 0; i < total; ++i) {
        const int slot = (int)(i % NSTAGE);
        if (i >= NSTAGE) wait_bar(bar(BAR_EMPTY + slot), (uint32_t)((i / NSTAGE - 1) & 1), p.err_flag, 1);
        ptx::mbar_arrive_expect_tx(bar(BAR_FULL + slot), STAGE_BYTES);
        ptx::bulk_g2s(sm_u32 + SP::RING + slot * STAGE_BYTES, p.packed + (size_t)srcs * STAGE_BYTES, STAGE_BYTES,
                      bar(BAR_FULL + slot));
        if (++srcs == P.n_stages) srcs = 0;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t IDESC = ptx::idesc_f16(128, 128, BF16 ? 1 : 0);
      long g = 0, stage_it = 0;
      // generations of each operand chunk already known to be complete: a barrier must not be
      // waited on again for a generation it has moved two phases past (the parity would alias)
      int seen[NUM_CHUNK_IDS];
      for (int i = 0; i < NUM_CHUNK_IDS; ++i) seen[i] = 0;
      for (int s = 0; s < S; ++s) {
        for (int ui = 0; ui < P.n_units; ++ui, ++g) {
          const Unit& u = P.u[ui];
          const int b = (int)(g & 3);
          if (g >= 4) wait_bar(bar(BAR_DEMPTY + b), (uint32_t)(((g >> 2) - 1) & 1), p.err_flag, 2);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)b * 128u;
          uint32_t accum = 0;
          for (int c = 0; c < u.n_chunks; ++c) {
            const int id = u.chunk[c];
            const int need = s * P.wps[id] + u.gen[c];
            if (need >= seen[id]) {
              wait_bar(bar(BAR_CHUNK + id), (uint32_t)(need & 1), p.err_flag, 3);
              seen[id] = need + 1;
            }
            ptx::tc_fence_after();
            uint32_t a_addr, lo_delta;
            if (id < 8) { a_addr = sm_u32 + SP::A + (X3 ? 0 : u.in_buf * 65536) + id * 8192; lo_delta = SP::A_LO; }
            else if (id < CH_V) { a_addr = sm_u32 + SP::E + (id - CH_E0) * 8192; lo_delta = SP::E_LO; }
            else if (id == CH_V) { a_addr = sm_u32 + SP::V; lo_delta = SP::V_LO; }
            else { a_addr = sm_u32 + SP::P; lo_delta = SP::P_LO; }
            const int nst = X3 ? (id == CH_P ? 1 : 2) : 1;
            for (int st = 0; st < nst; ++st, ++stage_it) {
              const int slot = (int)(stage_it % NSTAGE);
              wait_bar(bar(BAR_FULL + slot), (uint32_t)((stage_it / NSTAGE) & 1), p.err_flag, 4);
              ptx::tc_fence_after();
              const uint32_t b_addr = sm_u32 + SP::RING + slot * STAGE_BYTES;
              if (X3) {
                const uint64_t a_hi = ptx::smem_desc(a_addr + st * 4096, 2048, 128);
                const uint64_t a_lo = ptx::smem_desc(a_addr + lo_delta + st * 4096, 2048, 128);
                const uint64_t b_hi = ptx::smem_desc(b_addr, 2048, 128);
                const uint64_t b_lo = ptx::smem_desc(b_addr + 4096, 2048, 128);
                ptx::mma_f16_ss(d_tmem, a_hi, b_hi, IDESC, accum);
                ptx::mma_f16_ss(d_tmem, a_lo, b_hi, IDESC, 1);
                ptx::mma_f16_ss(d_tmem, a_hi, b_lo, IDESC, 1);
              } else {
                ptx::mma_f16_ss(d_tmem, ptx::smem_desc(a_addr, 2048, 128), ptx::smem_desc(b_addr, 2048, 128), IDESC, accum);
                ptx::mma_f16_ss(d_tmem, ptx::smem_desc(a_addr + 4096, 2048, 128), ptx::smem_desc(b_addr + 4096, 2048, 128),
                                IDESC, 1);
              }
              accum = 1;
              ptx::mma_commit(bar(BAR_EMPTY + slot));
            }
          }
          ptx::mma_commit(bar(BAR_DFULL + b));
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue / per-ray state ================================
    const int row = tid - 128;            // ray within the tile == TMEM lane
    const int quad = warp & 3;            // TMEM lane quadrant of this warp
    const long ray = (long)blockIdx.x * 128 + row;
    const bool valid = ray < p.R;
    const long rl = valid ? ray : (long)p.R - 1;
    const float ox = p.rays_o[3 * rl + 0], oy = p.rays_o[3 * rl + 1], oz = p.rays_o[3 * rl + 2];
    const float dx = p.rays_d[3 * rl + 0], dy = p.rays_d[3 * rl + 1], dz = p.rays_d[3 * rl + 2];
    const float dnorm = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    const float* tv = p.t_vals + (p.t_stride ? rl * p.t_stride : 0);
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);

    auto publish = [&](int id) {  // make this warp's operand stores visible to the MMA (async proxy)
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(BAR_CHUNK + id));
    };
    auto encode_point = [&](float x, float y, float z) {
      encode_store<10, 64, X3, BF16>(sm + SP::E, SP::E_LO, row, x, y, z);
      publish(CH_E0);
      publish(CH_E0 + 1);
    };

    // head weights / biases in shared memory
    constexpr int NBIAS = KIND == AON_KIND_VANILLA ? 2432 : 3328;
    const float* hw_def = s_par + NBIAS;                                             // auto-decoder only
    const float* hw_sig = s_par + NBIAS + (KIND == AON_KIND_AUTODECODER ? 388 : 0);
    const float* hw_rgb = hw_sig + 260;

    {  // view-direction encoding, once per tile (model.py:174: pos_enc(viewdirs, 0, 4))
      const float vx = p.viewdirs[3 * rl + 0], vy = p.viewdirs[3 * rl + 1], vz = p.viewdirs[3 * rl + 2];
      encode_store<4, 32, X3, BF16>(sm + SP::V, SP::V_LO, row, vx, vy, vz);
      publish(CH_V);
    }

    float trans = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, cdepth = 0.f, cacc = 0.f;
    float t_cur = tv[0];
    float px, py, pz;
    auto sample_point = [&](float t) {  // cast_rays (helper.py:25-26)
      px = __fadd_rn(ox, __fmul_rn(t, dx));
      py = __fadd_rn(oy, __fmul_rn(t, dy));
      pz = __fadd_rn(oz, __fmul_rn(t, dz));
    };
    sample_point(t_cur);
    if (KIND == AON_KIND_VANILLA) encode_point(px, py, pz);

    long g = 0;
    for (int s = 0; s < S; ++s) {
      const float t_next = (s + 1 < S) ? tv[s + 1] : 0.f;
      if (KIND == AON_KIND_AUTODECODER) {
        // raw position -> operand chunk P (deformation MLP input; model_autodecoder.py:196-198)
        float v[8] = {px, py, pz, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint4 hi, lo;
        pack8<X3, BF16>(v, hi, lo);
        const uint4 z4 = make_uint4(0, 0, 0, 0);
        constexpr int PG = X3 ? 2 : 4;  // k-groups of the P chunk
#pragma unroll
        for (int kg = 0; kg < PG; ++kg) {
          *reinterpret_cast<uint4*>(sm + SP::P + kg * 2048 + row * 16) = kg == 0 ? hi : z4;
          if (X3) *reinterpret_cast<uint4*>(sm + SP::P + SP::P_LO + kg * 2048 + row * 16) = kg == 0 ? lo : z4;
        }
        publish(CH_P);
      }
      float sig = 0.f, h0 = 0.f, h1 = 0.f, h2 = 0.f;  // head accumulators (density; rgb / deformation)

      for (int ui = 0; ui < P.n_units; ++ui, ++g) {
        const Unit& u = P.u[ui];
        const int b = (int)(g & 3);
        wait_bar(bar(BAR_DFULL + b), (uint32_t)((g >> 2) & 1), p.err_flag, 5);
        if (u.wait_next) wait_bar(bar(BAR_DFULL + ((b + 1) & 3)), (uint32_t)(((g + 1) >> 2) & 1), p.err_flag, 6);
        ptx::tc_fence_after();
        const float* bias = s_par + u.bias_off;
        const int epi = u.epi;
        const bool relu = u.relu != 0;
        if (epi == EPI_RGB || epi == EPI_DEFORM) { h0 = h1 = h2 = 0.f; }
        const float* hw = epi == EPI_STORE_SIGMA ? hw_sig + u.half * 128 : (epi == EPI_RGB ? hw_rgb : hw_def);
        unsigned char* out_base = sm + SP::A + (X3 ? 0 : u.out_buf * 65536) + (u.half * 4) * 8192 + row * 16;

        uint32_t r[2][32];
        ptx::tmem_ld32(lane_base + (uint32_t)(b * 128), r[0]);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          tmem_ld_wait_dep(r[cc & 1]);
          if (cc + 1 < 4) ptx::tmem_ld32(lane_base + (uint32_t)(b * 128 + (cc + 1) * 32), r[(cc + 1) & 1]);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias + cc * 32 + i);
            v[i + 0] = __uint_as_float(r[cc & 1][i + 0]) + b4.x;
            v[i + 1] = __uint_as_float(r[cc & 1][i + 1]) + b4.y;
            v[i + 2] = __uint_as_float(r[cc & 1][i + 2]) + b4.z;
            v[i + 3] = __uint_as_float(r[cc & 1][i + 3]) + b4.w;
          }
          if (p.dbg != nullptr && blockIdx.x == 0 && s == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) p.dbg[((size_t)ui * 128 + row) * 128 + cc * 32 + i] = v[i];
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (epi == EPI_STORE_SIGMA) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(hw + cc * 32 + i);
              sig = fmaf(w4.x, v[i], sig); sig = fmaf(w4.y, v[i + 1], sig);
              sig = fmaf(w4.z, v[i + 2], sig); sig = fmaf(w4.w, v[i + 3], sig);
            }
          }
          if (epi == EPI_RGB || epi == EPI_DEFORM) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 w0 = *reinterpret_cast<const float4*>(hw + cc * 32 + i);
              const float4 w1 = *reinterpret_cast<const float4*>(hw + 128 + cc * 32 + i);
              const float4 w2 = *reinterpret_cast<const float4*>(hw + 256 + cc * 32 + i);
              h0 = fmaf(w0.x, v[i], h0); h0 = fmaf(w0.y, v[i + 1], h0); h0 = fmaf(w0.z, v[i + 2], h0); h0 = fmaf(w0.w, v[i + 3], h0);
              h1 = fmaf(w1.x, v[i], h1); h1 = fmaf(w1.y, v[i + 1], h1); h1 = fmaf(w1.z, v[i + 2], h1); h1 = fmaf(w1.w, v[i + 3], h1);
              h2 = fmaf(w2.x, v[i], h2); h2 = fmaf(w2.y, v[i + 1], h2); h2 = fmaf(w2.z, v[i + 2], h2); h2 = fmaf(w2.w, v[i + 3], h2);
            }
          } else {
            // next layer's A operand: 32 hidden features of this row -> chunk (half*4 + cc)
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
              float w8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) w8[i] = v[kg * 8 + i];
              uint4 hi, lo;
              pack8<X3, BF16>(w8, hi, lo);
              *reinterpret_cast<uint4*>(out_base + cc * 8192 + kg * 2048) = hi;
              if (X3) *reinterpret_cast<uint4*>(out_base + SP::A_LO + cc * 8192 + kg * 2048) = lo;
            }
            publish(u.half * 4 + cc);
          }
        }
        // accumulator drained: hand the TMEM buffer back to the MMA issuer
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar(BAR_DEMPTY + b));

        if (KIND == AON_KIND_AUTODECODER && epi == EPI_DEFORM) {
          // model_autodecoder.py:203: x' = deformation_layer(h) + pos, then pos_enc(x', 0, 10)
          const float* hb = hw_def + 384;
          const float ex = __fadd_rn(h0 + hb[0], px), ey = __fadd_rn(h1 + hb[1], py), ez = __fadd_rn(h2 + hb[2], pz);
          encode_point(ex, ey, ez);
        }
      }

      // ---- activations + alpha compositing of sample s (helper.py:157-195) ----
      {
        const float raw_sigma = sig + hw_sig[256];
        const float* hb = hw_rgb + 384;
        float rr = h0 + hb[0], gg = h1 + hb[1], bb = h2 + hb[2];
        float sigma;
        if (KIND == AON_KIND_VANILLA) {
          rr = sigmoidf_ref(rr); gg = sigmoidf_ref(gg); bb = sigmoidf_ref(bb);   // model.py:186
          sigma = fmaxf(raw_sigma, 0.f);                                           // model.py:187
        } else {                                                                   // model_autodecoder.py:321-323
          rr = __fsub_rn(__fmul_rn(sigmoidf_ref(rr), 1.002f), 0.001f);
          gg = __fsub_rn(__fmul_rn(sigmoidf_ref(gg), 1.002f), 0.001f);
          bb = __fsub_rn(__fmul_rn(sigmoidf_ref(bb), 1.002f), 0.001f);
          sigma = softplusf_ref(__fadd_rn(raw_sigma, -1.0f));
        }
        const float delta = (s + 1 < S) ? __fsub_rn(t_next, t_cur) : 1e10f;
        const float dist = __fmul_rn(delta, dnorm);
        const float alpha = __fsub_rn(1.0f, expf(__fmul_rn(-sigma, dist)));
        const float w = __fmul_rn(alpha, trans);
        cr = fmaf(w, rr, cr); cg = fmaf(w, gg, cg); cb = fmaf(w, bb, cb);
        cdepth = fmaf(w, t_cur, cdepth);
        cacc += w;
        trans = __fmul_rn(trans, __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f));
        if (p.weights && valid) p.weights[ray * S + s] = w;
      }
      t_cur = t_next;
      if (s + 1 < S) {
        sample_point(t_cur);
        if (KIND == AON_KIND_VANILLA) encode_point(px, py, pz);
      }
    }

    if (valid) {
      if (isnan(cdepth)) cdepth = INFINITY;  // helper.py:179 nan_to_num(depth, nan=inf)
      else if (isinf(cdepth)) cdepth = cdepth > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
      if (p.white_bkgd) {
        const float bg = __fsub_rn(1.0f, cacc);
        cr += bg; cg += bg; cb += bg;
      }
      p.comp_rgb[3 * ray + 0] = cr; p.comp_rgb[3 * ray + 1] = cg; p.comp_rgb[3 * ray + 2] = cb;
      p.acc[ray] = cacc;
      p.depth[ray] = cdepth;
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side -------------------------------------------------------------------------------------------
static float* g_dbg = nullptr;
static int* g_err = nullptr;

template <int KIND, bool X3, bool BF16>
static int launch(const TcParams& p, int grid, cudaStream_t st) {
  using SP = SmemPlan<KIND, X3>;
  auto kern = render_tc_kernel<KIND, X3, BF16>;
  AON_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SP::TOTAL));
  kern<<<grid, TC_THREADS, SP::TOTAL, st>>>(p);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

int render_level_tc(int kind, int precision, const void* packed, const float* folded, const float* rays_o,
                    const float* rays_d, const float* viewdirs, const float* t_vals, long t_stride, int R, int S,
                    int white_bkgd, float* comp_rgb, float* acc, float* depth, float* weights, cudaStream_t st) {
  int dev = 0, major = 0;
  AON_CUDA_CHECK(cudaGetDevice(&dev));
  AON_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("tensor-core precision modes need an sm_100 device (found compute capability %d.x)", major);
    return AON_E_UNSUPPORTED;
  }
  TcParams p;
  p.prog = build_program(kind, precision);
  p.L = layout_tc(kind, precision);
  p.packed = (const char*)packed;
  p.folded = folded;
  p.rays_o = rays_o; p.rays_d = rays_d; p.viewdirs = viewdirs;
  p.t_vals = t_vals; p.t_stride = t_stride;
  p.R = R; p.S = S; p.white_bkgd = white_bkgd;
  p.comp_rgb = comp_rgb; p.acc = acc; p.depth = depth; p.weights = weights;
  p.dbg = g_dbg;
  p.err_flag = g_err;
  const int grid = (R + 127) / 128;
  const bool van = kind == AON_KIND_VANILLA;
  switch (precision) {
    case AON_PREC_TC_F16X3:
      return van ? launch<AON_KIND_VANILLA, true, false>(p, grid, st) : launch<AON_KIND_AUTODECODER, true, false>(p, grid, st);
    case AON_PREC_TC_F16:
      return van ? launch<AON_KIND_VANILLA, false, false>(p, grid, st) : launch<AON_KIND_AUTODECODER, false, false>(p, grid, st);
    case AON_PREC_TC_BF16:
      return van ? launch<AON_KIND_VANILLA, false, true>(p, grid, st) : launch<AON_KIND_AUTODECODER, false, true>(p, grid, st);
  }
  set_error("bad precision %d", precision);
  return AON_E_ARG;
}

int pack_tail(int kind, const PackedLayout& L, const float* const* w, const float* const* b, char* packed,
              cudaStream_t st);  // aon_api.cu

}  // namespace aon

using namespace aon;

extern "C" int aon_pack_weights_tc(int kind, int precision, const float* const* w, const float* const* b, void* packed,
                                   size_t packed_bytes, cudaStream_t st) {
  (void)packed_bytes;
  const Program P = build_program(kind, precision);
  const PackedLayout L = layout_tc(kind, precision);
  PackSrc src;
  memset(&src, 0, sizeof(src));
  for (int i = 0; i < num_layers(kind); ++i) src.w[i] = w[i];
  for (int i = 0; i < num_gemm(kind); ++i) {
    src.g[i] = gemm_layers(kind)[i];
    src.in_features[i] = layer_shapes(kind)[src.g[i].src][1];
  }
  uint16_t* out = (uint16_t*)packed;
  const int blocks = 592;
  if (precision == AON_PREC_TC_F16X3) pack_stream_kernel<true, false><<<blocks, 256, 0, st>>>(P, src, kind, out);
  else if (precision == AON_PREC_TC_F16) pack_stream_kernel<false, false><<<blocks, 256, 0, st>>>(P, src, kind, out);
  else pack_stream_kernel<false, true><<<blocks, 256, 0, st>>>(P, src, kind, out);
  AON_LAUNCH_CHECK();
  return pack_tail(kind, L, w, b, (char*)packed, st);
}

// Debug hooks (not part of the public ABI in include/aon.h): a device buffer that receives the
// pre-activation outputs of every unit for tile 0 / sample 0, and a device int that receives a code
// if a barrier wait ever times out.
extern "C" void aon_debug_set_buffers(float* dbg_dev, int* err_dev) {
  g_dbg = dbg_dev;
  g_err = err_dev;
}
extern "C" int aon_debug_program_info(int kind, int precision, int* n_units, int* n_stages, int* smem_bytes) {
  const Program P = build_program(kind, precision);
  if (n_units) *n_units = P.n_units;
  if (n_stages) *n_stages = P.n_stages;
  if (smem_bytes) {
    const bool x3 = precision == AON_PREC_TC_F16X3;
    *smem_bytes = kind == AON_KIND_VANILLA ? (x3 ? SmemPlan<0, true>::TOTAL : SmemPlan<0, false>::TOTAL)
                                           : (x3 ? SmemPlan<1, true>::TOTAL : SmemPlan<1, false>::TOTAL);
  }
  return 0;
}
