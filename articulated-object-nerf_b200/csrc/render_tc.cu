// render_tc.cu -- tensor-core (tcgen05 / TMEM) version of the fused per-level render kernel.
//
// A CTA PAIR (cluster of 2, tcgen05 cta_group::2) renders 256 rays, 128 per CTA.  MMA row m = ray m of
// the CTA's tile, and each CTA walks the S samples of its rays one 128-row "sample plane" at a time,
// so a ray's transmittance / colour / depth accumulators are running scalars in one epilogue thread and
// no [rays x samples x features] tensor ever reaches HBM (reference: helper.py:25-26,136-140,157-195;
// model.py:95-120,174-195; model_autodecoder.py:171-239,306-331).
//
// Work decomposition.  Every nn.Linear with >= 128 outputs is one "unit" (all N outputs x all K).  K is
// cut into "chunks": 32 hidden features written by the epilogue of the previous layer (ids 0..7), one
// half of the 64-wide positional encoding (8, 9), the 32-wide view-direction encoding (10), the raw
// sample position of the deformation MLP (11), or the constant "ones" chunk (12) whose weight rows carry
// the layer's bias split into three 16-bit pieces -- the bias add happens inside the fp32 accumulation
// of the tensor core, not in the epilogue.
//   warp 12    producer: streams THIS CTA's half (N/2 output features) of the pre-packed weight stream
//              (8 KB stages, already in the UMMA canonical K-major layout) from L2 into a shared-memory
//              ring with 1-D bulk copies: every SM ingests, stores and feeds to its tensor core only half
//              of every B tile (cta_group::2 exchanges the halves in hardware), which lifts the
//              M=128,N=256 MMA from 171 cycles (shared-memory operand bound, measured) to the 128-cycle floor
//   warps 13,14 leader CTA: MMA issuers -- two threads (alternate weight stages) issue tcgen05.mma.cta_group::2 (M=256 over the pair,
//              N=256/128, K=16) into one of two 256-column TMEM accumulators of BOTH CTAs; a K chunk is
//              issued as soon as the epilogues of both CTAs have published it, so layer l+1 starts while
//              layer l is still being drained.  peer CTA: relay -- forwards "my half of stage s landed" to
//              the leader's full barrier (mbarrier.try_wait only works on the local CTA)
//   warps 8-11 encoder: cast_rays + pos_enc of the next sample (one ray per thread), off the critical path
//   warps 0-7  epilogue, two warps per TMEM lane quadrant (even / odd 32-column chunks): tcgen05.ld the
//              accumulators (one TMEM lane = one ray = one thread), ReLU, convert to the 16-bit operand
//              format and store the next layer's A operand chunk to shared memory; the 1-/3-wide heads
//              (density, rgb, deformation) are fp32 FMAs on the fp32 accumulators (partial sums of the odd
//              warp handed to the even warp through shared memory); the even warps own the per-ray state:
//              activations + alpha compositing in registers.
// Precision modes: AON_PREC_TC_F16 / _BF16 = one MMA per K step; AON_PREC_TC_F16X3 = operands split
// into fp16 hi + fp16 lo (about 22 significand bits), three MMAs per K step
// (hi*hi + lo*hi + hi*lo) with fp32 accumulation -- the mode that meets the 1e-4 parity bar.
#include <stdlib.h>
#include <string.h>

#include "aon_common.cuh"
#include "sampling.cuh"
#include "tc_ptx.cuh"

namespace aon {

void layout_tail(int kind, PackedLayout& L, int64_t off);  // aon_api.cu

// ---- program (unit schedule), built on the host, passed to the kernels by value ------------------
constexpr int MAX_UNITS = 28;
constexpr int MAX_CHUNKS = 11;
constexpr int CH_E0 = 8, CH_V = 10, CH_P = 11, CH_ONE = 12, NUM_CHUNK_IDS = 13;
enum Epi : int { EPI_STORE = 0, EPI_STORE_SIGMA = 1, EPI_RGB = 2, EPI_DEFORM = 3 };

// shared-memory offsets of the operand regions (bytes from the 128-byte aligned base); host and device
constexpr int OFF_A = 0;
__host__ __device__ constexpr int off_E(bool x3) { return x3 ? 131072 : 65536; }
__host__ __device__ constexpr int off_V(bool x3) { return off_E(x3) + (x3 ? 32768 : 16384); }
__host__ __device__ constexpr int off_ONE(bool x3) { return off_V(x3) + (x3 ? 16384 : 8192); }
__host__ __device__ constexpr int off_P(bool x3) { return off_ONE(x3) + 4096; }
constexpr int LO_A = 65536, LO_E = 16384, LO_V = 8192, LO_P = 4096;  // hi -> lo part (x3 only)

// Exact power-of-two operand scaling.  fp16 has only 5 exponent bits: the lo part of a hi+lo split of a
// typical weight (|w| ~ 0.05 -> lo ~ 1e-5) or activation falls into the fp16 subnormal range and loses
// bits.  Weights are stored as SW*w, activations (every A operand chunk) as SA*a, bias rows as SW*SA*b, so the
// accumulator holds SW*SA*(W a + b); the epilogue multiplies by 1/SW (exact) and gets SA * pre-activation,
// i.e. the next layer's scaled operand; the fp32 head weights are pre-divided by SA.
constexpr float SCALE_W = 64.0f, SCALE_A = 8.0f;

// chunk word: [0,16) operand offset >> 4 | [16,20) chunk id | 20 fresh (first use of a generation: wait)
//             | 21 writes-per-sample odd | 22 generation parity within a sample | [23,26) K steps of 16 (1, 2, 4)
//             | [26,31) (lo part offset) >> 12 | 31 pair: covers chunk ids id and id + 1 (one-pass modes merge
//             adjacent 32-wide chunks into one K=64 weight stage = 4 MMAs per issue step)
struct Unit {
  uint8_t gemm, n_chunks, epi, relu;
  uint8_t n128;        // N / 128 (1 or 2)
  uint8_t last_e_use;  // the encoding operand may be overwritten once this unit's MMAs have completed
  uint8_t n_hidden;    // number of 32-wide hidden input chunks (K1 / 32)
  uint8_t last_hidden; // chunk list entry that holds the last hidden input chunk (0 if none)
  int16_t fold;        // >= 0: the bias stage comes from the per-call folded buffer, at fold * 16 bytes (per-rank part)
  uint32_t ch[MAX_CHUNKS];
};

struct Program {
  int n_units, n_gemm, x3;
  long stream_bytes;           // weight stream bytes per sample PER CTA of the pair (the blob holds two such halves)
  long fold_half_bytes;        // per-rank size of the folded bias stages (auto-decoder) or 0
  Unit u[MAX_UNITS];
};

// Per-CTA weight stages (each CTA of the pair streams its own N/2 rows of B):
//   one pass : one stage per chunk entry (K = 64 for merged pairs, 32 otherwise) [K/8 k-groups][N/2][8] = N/2 * 2K bytes
//   x3       : one stage per K=16 step         hi [2 k-groups][N/2][8], then lo likewise = N/2 * 64 bytes
//   bias     : one stage per unit (chunk CH_ONE, K=16) [2 k-groups][N/2][8]              = N/2 * 32 bytes
__host__ __device__ inline int chunk_id(uint32_t w) { return (int)((w >> 16) & 15u); }
__host__ __device__ inline int chunk_ksteps(uint32_t w) { return (int)((w >> 23) & 7u); }
__host__ __device__ inline int stage_bytes(const Unit& u, uint32_t w, int x3) {
  if (x3) return (u.n128 * 4096) >> (chunk_id(w) == CH_ONE ? 1 : 0);
  return u.n128 * 2048 * chunk_ksteps(w);
}
__host__ __device__ inline int chunk_stages(uint32_t w, int x3) { return (x3 && chunk_id(w) != CH_ONE) ? chunk_ksteps(w) : 1; }
__host__ __device__ inline long unit_bytes(const Unit& u, int x3) {
  long b = 0;
  for (int c = 0; c < u.n_chunks; ++c) b += (long)chunk_stages(u.ch[c], x3) * stage_bytes(u, u.ch[c], x3);
  return b;
}

static Program build_program(int kind, int precision) {
  Program P;
  memset(&P, 0, sizeof(P));
  const int x3 = precision == AON_PREC_TC_F16X3;
  P.x3 = x3;
  const GemmLayer* g = gemm_layers(kind);
  const int ng = num_gemm(kind);
  int count[NUM_CHUNK_IDS] = {0}, seen_gen[NUM_CHUNK_IDS];
  for (int i = 0; i < NUM_CHUNK_IDS; ++i) seen_gen[i] = -1;
  if (kind == AON_KIND_VANILLA) count[CH_E0] = count[CH_E0 + 1] = 1;
  else count[CH_P] = 1;
  count[CH_V] = 1;
  struct Use { int ui, c, id, gen, pair; };
  Use uses[MAX_UNITS * MAX_CHUNKS];
  int n_uses = 0, last_e = -1;
  long fold_off = 0;
  for (int gi = 0; gi < ng; ++gi) {
    Unit& u = P.u[gi];
    int epi = EPI_STORE;
    if (kind == AON_KIND_VANILLA) {
      if (gi == 7) epi = EPI_STORE_SIGMA;
      if (gi == 9) epi = EPI_RGB;
    } else {
      if (gi == 3) epi = EPI_DEFORM;
      if (gi == 11) epi = EPI_STORE_SIGMA;
      if (gi == 16) epi = EPI_RGB;
    }
    u.gemm = gi; u.epi = epi; u.relu = g[gi].relu; u.n128 = g[gi].N / 128;
    u.n_hidden = (uint8_t)(g[gi].K1 / 32);
    u.fold = -1;
    if (g[gi].lat_col0 >= 0) {   // latent-conditioned layer: per-call bias stage (aon_fold_latents)
      u.fold = (int16_t)(fold_off / 16);
      fold_off += g[gi].N / 2 * 32;
    }
    int ids[MAX_CHUNKS], pairs[MAX_CHUNKS] = {0}, nc = 0;
    ids[nc++] = CH_ONE;          // bias first: needs no activation, so the unit can start at once
    const int step = x3 ? 1 : 2; // one-pass modes: adjacent chunks (2j, 2j+1) share one K=64 weight stage
    for (int j = 0; j < g[gi].K1 / 32; j += step) { pairs[nc] = step == 2; ids[nc++] = j; }
    u.last_hidden = (uint8_t)(g[gi].K1 ? nc - 1 : 0);
    if (g[gi].aux == AUX_E) {
      pairs[nc] = step == 2; ids[nc++] = CH_E0;
      if (step == 1) ids[nc++] = CH_E0 + 1;
      last_e = gi;
    }
    if (g[gi].aux == AUX_V) ids[nc++] = CH_V;
    if (g[gi].aux == AUX_P) ids[nc++] = CH_P;
    u.n_chunks = nc;
    for (int c = 0; c < nc; ++c) uses[n_uses++] = {gi, c, ids[c], count[ids[c]] - 1, pairs[c]};
    if (epi <= EPI_STORE_SIGMA)
      for (int j = 0; j < u.n128 * 4; ++j) count[j]++;
    if (epi == EPI_DEFORM) { count[CH_E0]++; count[CH_E0 + 1]++; }
  }
  P.fold_half_bytes = fold_off;
  count[CH_V] = 0;  // written once per CTA: its generation never advances
  for (int i = 0; i < n_uses; ++i) {
    const Use& x = uses[i];
    int off, lo;
    if (x.id < 8) { off = OFF_A + x.id * 8192; lo = LO_A; }
    else if (x.id < CH_V) { off = off_E(x3) + (x.id - CH_E0) * 8192; lo = LO_E; }
    else if (x.id == CH_V) { off = off_V(x3); lo = LO_V; }
    else if (x.id == CH_P) { off = off_P(x3); lo = LO_P; }
    else { off = off_ONE(x3); lo = 0; }
    const int ksteps = (x.id == CH_ONE || (x3 && x.id == CH_P)) ? 1 : (x.pair ? 4 : 2);
    const int fresh = x.id != CH_ONE && (x.gen != seen_gen[x.id] || x.id == CH_V);
    seen_gen[x.id] = x.gen;
    uint32_t w = (uint32_t)(off >> 4) | ((uint32_t)x.id << 16) | ((uint32_t)fresh << 20) |
                 ((uint32_t)(count[x.id] & 1) << 21) | ((uint32_t)(x.gen & 1) << 22) | ((uint32_t)ksteps << 23) |
                 ((uint32_t)(lo >> 12) << 26) | ((uint32_t)x.pair << 31);
    P.u[x.ui].ch[x.c] = w;
  }
  long bytes = 0;
  for (int gi = 0; gi < ng; ++gi) bytes += unit_bytes(P.u[gi], x3);
  if (last_e >= 0) P.u[last_e].last_e_use = 1;
  P.n_units = ng;
  P.n_gemm = ng;
  P.stream_bytes = bytes;
  return P;
}

PackedLayout layout_tc(int kind, int precision) {
  PackedLayout L;
  memset(&L, 0, sizeof(L));
  const Program P = build_program(kind, precision);
  for (int i = 0; i < num_gemm(kind); ++i) L.w[i] = 0;  // one contiguous stream in unit order
  layout_tail(kind, L, (2 * P.stream_bytes + 255) / 256 * 256);
  return L;
}

// folded buffer (auto-decoder): [A_FOLDED_FLOATS fp32 biases | A_LATENT_FLOATS staging | bias stages rank 0 | rank 1]
constexpr int FOLD_STAGE_FLOAT0 = A_FOLDED_FLOATS + A_LATENT_FLOATS;
size_t folded_floats_tc(int kind) {
  if (kind != AON_KIND_AUTODECODER) return 0;
  const Program P = build_program(kind, AON_PREC_TC_F16);
  return (size_t)FOLD_STAGE_FLOAT0 + (size_t)(2 * P.fold_half_bytes) / 4;
}

// ---- weight stream packing ----------------------------------------------------------------------------
// The blob holds two streams, one per CTA rank of the pair; stream r is the B operand rows
// n in [r*N/2, (r+1)*N/2) of every MMA of one sample plane, in issue order, already in the layout
// tcgen05.mma reads (K-major, no swizzle: [k-group][n][8 k] 16-bit, 8-row x 16-byte core matrices).
struct PackSrc {
  const float* w[20];
  const float* b[20];
  GemmLayer g[MAX_GEMM];
  int in_features[MAX_GEMM];
  long unit_byte0[MAX_UNITS + 1];   // within one half stream
};

// value -> piece k (0 hi, 1 mid, 2 lo) of its 3-way 16-bit split
template <bool BF16>
__device__ __forceinline__ uint16_t split3(float v, int k) {
  float r = v;
  uint16_t bits = 0;
  for (int i = 0; i <= k; ++i) {
    if (BF16) {
      const __nv_bfloat16 h = __float2bfloat16_rn(r);
      bits = __bfloat16_as_ushort(h);
      r -= __bfloat162float(h);
    } else {
      const __half h = __float2half_rn(r);
      bits = __half_as_ushort(h);
      r -= __half2float(h);
    }
  }
  return bits;
}

template <bool X3, bool BF16>
__global__ void pack_stream_kernel(Program P, PackSrc src, uint16_t* __restrict__ out) {
  const long half = P.stream_bytes / 2;   // 16-bit elements per half stream
  const long total = 2 * half;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int rank = (int)(idx / half);
    const long hidx = idx % half;
    int ui = 0;
    while (ui + 1 < P.n_units && src.unit_byte0[ui + 1] <= hidx * 2) ++ui;
    const Unit& u = P.u[ui];
    const int NH = u.n128 * 64;
    const GemmLayer& g = src.g[u.gemm];
    long rel = hidx - src.unit_byte0[ui] / 2;   // element within the unit
    int c = 0, stage = 0;
    for (;; ++c) {
      const long se = stage_bytes(u, u.ch[c], X3) / 2;
      const long ce = (long)chunk_stages(u.ch[c], X3) * se;
      if (rel < ce) { stage = (int)(rel / se); rel -= (long)stage * se; break; }
      rel -= ce;
    }
    const int e = (int)rel;
    const int id = chunk_id(u.ch[c]);
    uint16_t bits;
    if (id == CH_ONE) {
      // bias rows: k = 0,1,2 carry the three pieces of bias[n]; the A side has ones in those columns
      const int k = (e / (NH * 8)) * 8 + (e & 7);
      const int n = ((e % (NH * 8)) >> 3) + rank * NH;
      bits = k < 3 ? split3<BF16>(src.b[g.src][n] * (SCALE_W * SCALE_A), k) : (uint16_t)0;
    } else {
      int part = 0, kin, n;
      if (X3) {
        part = e / (NH * 16);
        const int r = e % (NH * 16);
        kin = stage * 16 + (r / (NH * 8)) * 8 + (r & 7);
        n = (r % (NH * 8)) >> 3;
      } else {
        kin = (e / (NH * 8)) * 8 + (e & 7);
        n = (e % (NH * 8)) >> 3;
      }
      n += rank * NH;
      int col = -1;
      if (id < 8) col = id * 32 + kin;
      else {
        const int a = (id == CH_E0 + 1 ? 32 : 0) + kin;
        if (a < g.aux_cnt) col = g.aux_col0 + a;
      }
      float v = 0.f;
      if (col >= 0) v = src.w[g.src][(size_t)n * src.in_features[u.gemm] + col] * SCALE_W;
      if (BF16) {
        bits = __bfloat16_as_ushort(__float2bfloat16_rn(v));
      } else {
        const __half hi = __float2half_rn(v);
        if (X3 && part == 1) bits = __half_as_ushort(__float2half_rn(v - __half2float(hi)));
        else bits = __half_as_ushort(hi);
      }
    }
    out[idx] = bits;
  }
}

// per-call bias stages of the latent-conditioned layers: folded fp32 biases -> [rank][layer][2 k-groups][N/2][8]
template <bool BF16>
__global__ void fold_stage_kernel(Program P, const float* __restrict__ folded, PackedLayout L, uint16_t* __restrict__ out) {
  for (int ui = 0; ui < P.n_units; ++ui) {
    const Unit& u = P.u[ui];
    if (u.fold < 0) continue;
    const int NH = u.n128 * 64;
    const float* bias = folded + L.fold[u.gemm];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * NH * 16; idx += gridDim.x * blockDim.x) {
      const int rank = idx / (NH * 16), e = idx % (NH * 16);
      const int k = (e / (NH * 8)) * 8 + (e & 7);
      const int n = ((e % (NH * 8)) >> 3) + rank * NH;
      out[((size_t)rank * P.fold_half_bytes + (size_t)u.fold * 16) / 2 + e] = k < 3 ? split3<BF16>(bias[n] * (SCALE_W * SCALE_A), k) : (uint16_t)0;
    }
  }
}

int fold_stages_tc(int kind, int precision, const PackedLayout& L, float* folded, cudaStream_t st) {
  const Program P = build_program(kind, precision);
  uint16_t* out = reinterpret_cast<uint16_t*>(folded + FOLD_STAGE_FLOAT0);
  if (precision == AON_PREC_TC_BF16) fold_stage_kernel<true><<<8, 256, 0, st>>>(P, folded, L, out);
  else fold_stage_kernel<false><<<8, 256, 0, st>>>(P, folded, L, out);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

// ---- render kernel -----------------------------------------------------------------------------------------
constexpr int TC_THREADS = 512;  // 16 warps = 4 warpgroups: 0-7 epilogue | 8-11 encoder | 12 producer, 13 MMA issuer B, 14 MMA issuer A / relay, 15 idle.
// 512 threads start with 128 registers each; setmaxnreg moves registers from the light warpgroups to the epilogue:
// 256 x 184 + 128 x 96 + 128 x 48 = 65536.
constexpr int REGS_EPILOGUE = 176, REGS_ENCODER = 96, REGS_CONTROL = 56;
// The warp scheduler favours the highest warp id of a sub-partition (measured, B300_MICROARCH.md): the two
// single-thread latency-critical roles sit above the ALU-heavy epilogue warps that share their sub-partitions.
constexpr int SMEM_MAX = 232448;
constexpr int MAX_STAGES = 12;

template <int KIND, bool X3>
struct SmemPlan {
  static constexpr int P = off_P(X3);
  static constexpr int P_BYTES = KIND == AON_KIND_AUTODECODER ? 8192 : 0;
  static constexpr int PARAMS = P + P_BYTES;           // fp32 head weights + head biases
  static constexpr int PARAM_FLOATS = KIND == AON_KIND_AUTODECODER ? 1040 : 648;
  static constexpr int XCH = PARAMS + PARAM_FLOATS * 4;  // [4][128] fp32 head partial sums (odd -> even epilogue warp)
  static constexpr int BARS = XCH + 2048;
  static constexpr int BAR_BYTES = 384;
  static constexpr int RING = (BARS + BAR_BYTES + 127) / 128 * 128;
  static constexpr int STAGE = X3 ? 8192 : 16384;      // ring slot size (stages of 128-wide layers / bias stages use less)
  static constexpr int NSTAGE_RAW = (SMEM_MAX - 128 - RING) / STAGE;   // 128 B of slack for aligning the dynamic smem base
  // even: the two MMA issuers take alternate stages, so with an even ring each thread only ever re-uses slots whose
  // previous round it waited for itself (a parity wait on a barrier that is still a round behind passes falsely)
  static constexpr int NSTAGE = (NSTAGE_RAW > MAX_STAGES ? MAX_STAGES : NSTAGE_RAW) & ~1;
  static constexpr int TOTAL = RING + NSTAGE * STAGE + 128;
  static_assert(NSTAGE >= 3, "weight ring too small");
};

// Training forward (TRAIN instantiations): every unit's output -- the activation the backward GEMMs of csrc/gemm_tc.cu need as
// wgrad operand and ReLU mask -- leaves the epilogue ONCE, as 16-byte rows of the packed plane layout of gemm_tc.cu
// ([tile][feature / 8][128 rows][8] 16-bit, hi and lo plane), which is this kernel's shared-memory operand layout; the next
// layer reads it from shared memory, never from HBM.  Row tile of (ray tile rt, sample s) = rt * S + s, row = ray % 128.
struct TrainDump {
  uint4* hi[MAX_UNITS];
  uint4* lo[MAX_UNITS];
  uint32_t* bits[MAX_UNITS];   // [tiles * 128][N / 32] ReLU masks (bit i of word c: feature 32 c + i > 0) or nullptr
  uint4* e_hi;                 // encoding operand [tiles][8][128]
  uint4* e_lo;
  float4* raw;                 // [R * S] in ray-major order: (r, g, b) before the sigmoid, raw density
  float* warped;               // auto-decoder: [tiles * 128][3] warped positions, tile order
  int dbg_flags;               // profiling experiments only (AON_TRAIN_DEBUG): 1 = no plane stores, 2 = no mask words, 4 = no chunk barrier
};

struct TcParams {
  Program prog;
  TrainDump dump;
  PackedLayout L;
  // per level: the fused image kernel walks level 0 (coarse) then level 1 (fine); a classic launch has level 0 only
  const char* packed[2];
  const float* folded[2];
  int S_level[2];
  int n_levels;
  // ray source: arrays [R,3], or a pinhole camera (cam_on; pixel index = ray0 + ray; sampling.cuh ray_from_camera)
  const float* rays_o;
  const float* rays_d;
  const float* viewdirs;
  int cam_on, img_H, img_W;
  float focal;
  Cam cam;
  long ray0;
  // level-0 sample positions: shared table [S] (t_stride 0) or [R,S]
  const float* t_vals;
  long t_stride;
  // fused kernel only: inverse-cdf draws [R,nf] or nullptr (deterministic table), and the per-CTA scratch slots
  // ([n_slots][128 x (S0 + S1)] floats: coarse weights, then fine t) handed out through a bit mask
  const float* u;
  long u_stride;
  float* slots;
  unsigned* slot_mask;
  int n_slots;
  int R, white_bkgd;
  int seg_len, n_seg;  // classic launch: samples per CTA (blockIdx.y = segment) and number of segments
  float4* seg_samples; // n_seg > 1: [R][S] (alpha, r, g, b) per sample, composited in order by composite_samples_kernel
  // outputs of the last level: planes (comp_rgb [R,3], acc [R], depth [R]) or interleaved out5 [R,5]
  float* comp_rgb;
  float* acc;
  float* depth;
  float* out5;
  float* weights;      // classic launch, n_seg == 1: [R,S] compositing weights (optional)
  float* coarse_out5;  // fused kernel: [R,5] of level 0 (optional)
  float* dbg;        // optional [n_units][128][256] pre-activation dump of tile 0 / sample 0
  int* err_flag;     // optional: set to a non-zero code when a barrier wait times out
  long long* tl;     // optional timeline [3 roles][4 samples][MAX_UNITS][4 events] of SM clocks, CTA 0 only
};

// barrier slots (8 bytes each) inside the BARS region; the last two words hold the MMA issue ticket and the scratch slot
constexpr int BAR_FULL = 0, BAR_EMPTY = MAX_STAGES, BAR_DFULL = 2 * MAX_STAGES, BAR_DEMPTY = BAR_DFULL + 2,
              BAR_CHUNK = BAR_DEMPTY + 2, BAR_EFREE = BAR_CHUNK + NUM_CHUNK_IDS, BAR_XW = BAR_EFREE + 1,
              BAR_TMEM = BAR_XW + 2, WORD_TICKET = 2 * (BAR_TMEM + 1), WORD_SLOT = WORD_TICKET + 1;
static_assert((WORD_SLOT + 1) * 4 <= 384, "barrier region too small");

__device__ __forceinline__ void wait_slow(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  const long long t0 = clock64();
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s: a schedule bug, not a slow kernel
      if (err_flag) atomicExch(err_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (!ptx::mbar_try_wait(bar, parity)) wait_slow(bar, parity, err_flag, code);
}

// Lean wait for the single-thread control roles (MMA issuers, producer, relay): no time-out code.  The inlined
// time-out path of wait_bar sits between the barrier test and the instructions that follow it, so every test cost
// a taken branch over ~25 cold instructions and the issuers' hot loop was spread over a dozen instruction-cache
// lines (ncu: stall_branch_resolving + stall_no_inst dominated those warps).  A hang still traps: the epilogue and
// encoder warps keep their timed waits.
__device__ __forceinline__ void spin_bar(uint32_t bar, uint32_t parity) {
  while (!ptx::mbar_try_wait(bar, parity)) {
  }
}

// Issue ticket of the two MMA issuer threads: the number of weight stages whose MMAs have been ISSUED so far.  The thread
// that owns stage n does all its waiting first (accumulator drained, operand chunk published, weights landed), then
// spins until the ticket equals n, issues the stage's MMAs and passes the ticket on -- after issue, not after completion.
// tcgen05.mma instructions execute in the order they enter the tensor core's queue, so every accumulator sees its K
// steps in program order: results are bit-reproducible from run to run and independent of how rays are grouped into
// tiles.  The hand-over costs a shared-memory store + load (tens of cycles) per stage against ~385 cycles of MMA
// execution per stage.
__device__ __forceinline__ void ticket_wait(uint32_t addr, uint32_t want) {
  uint32_t v;
  do {
    asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  } while (v != want);
}
__device__ __forceinline__ void ticket_set(uint32_t addr, uint32_t v) {
  asm volatile("st.volatile.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <bool X3, bool BF16>
__device__ __forceinline__ void split16(float v, uint16_t& hi, uint16_t& lo) {
  if (BF16) {
    hi = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    lo = 0;
  } else {
    const __half h = __float2half_rn(v);
    hi = __half_as_ushort(h);
    lo = X3 ? __half_as_ushort(__float2half_rn(v - __half2float(h))) : (uint16_t)0;
  }
}

// element k of row `row` of an operand region lives at base + (k/8)*2048 + row*16 + (k%8)*2
template <bool X3, bool BF16>
__device__ __forceinline__ void put16(unsigned char* base, int lo_delta, int row, int k, float v) {
  uint16_t hi, lo;
  split16<X3, BF16>(v * SCALE_A, hi, lo);
  unsigned char* p = base + (k >> 3) * 2048 + row * 16 + (k & 7) * 2;
  *reinterpret_cast<uint16_t*>(p) = hi;
  if (X3) *reinterpret_cast<uint16_t*>(p + lo_delta) = lo;
}

// pos_enc (helper.py:136-140) of one point (row `row`) into an operand region of NK columns:
// [x y z | sin(2^f v_d) f-major | sin(2^f v_d + pi/2) f-major | zero padding]
template <int L, int NK, bool X3, bool BF16>
__device__ __forceinline__ void encode_store(unsigned char* base, int lo_delta, int row, const float (&x)[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) put16<X3, BF16>(base, lo_delta, row, d, x[d]);
#pragma unroll 1
  for (int f = 0; f < L; ++f) {
    const float sc = (float)(1 << f);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = x[d] * sc;  // exact (power of two)
      const float sn = sinf(v);
      const float cs = sinf(__fadd_rn(v, AON_HALF_PI_F));
      put16<X3, BF16>(base, lo_delta, row, 3 + 3 * f + d, sn);
      put16<X3, BF16>(base, lo_delta, row, 3 + 3 * L + 3 * f + d, cs);
    }
  }
#pragma unroll
  for (int k = 3 + 6 * L; k < NK; ++k) put16<X3, BF16>(base, lo_delta, row, k, 0.f);
}

template <bool X3, bool BF16>
__device__ __forceinline__ void pack8(const float* v, uint4& hi, uint4& lo) {
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

__device__ __forceinline__ uint64_t mk_desc(uint32_t lo32) {
  // upper word: SBO = 128 B (>>4 = 8) at bits [32,46), descriptor version 1 at bits [46,48)
  return ((uint64_t)(8u | (1u << 14)) << 32) | lo32;
}

// named barrier shared by the two epilogue warps of a TMEM lane quadrant (ids 1..4)
// (bar.sync / bar.arrive are warp-aligned instructions: __syncwarp() first, so that lanes that took different branches just
// before -- e.g. the padding rays of a ragged tile -- are converged again when the warp reaches the barrier)
__device__ __forceinline__ void pair_bar_sync(int id) { __syncwarp(); asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair_bar_arrive(int id) { __syncwarp(); asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
// named barrier of the epilogue + encoder warps (0-11) around the in-kernel hierarchical sampling of the fused kernel
// (not inlined: the epilogue and the encoder warps then wait at the SAME barrier instruction, which is what synccheck expects)
__device__ __noinline__ void level_bar_sync() { __syncwarp(); asm volatile("bar.sync 5, 384;" ::: "memory"); }

// ray `row` of this CTA's tile: origin, direction, view direction (clamped to the last ray for the padding rows)
__device__ __forceinline__ void load_ray(const TcParams& p, int row, float (&o)[3], float (&d)[3], float (&v)[3]) {
  const long ray = (long)blockIdx.x * 128 + row;
  const long rl = ray < p.R ? ray : (long)p.R - 1;
  if (p.cam_on) {
    ray_from_camera(p.ray0 + rl, p.img_H, p.img_W, p.focal, p.cam, o, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = d[k];
  } else {
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = p.rays_o[3 * rl + k]; d[k] = p.rays_d[3 * rl + k]; v[k] = p.viewdirs[3 * rl + k]; }
  }
}

template <int KIND, bool X3, bool BF16, bool TRAIN>
__global__ void __launch_bounds__(TC_THREADS, 1) render_tc_kernel(const __grid_constant__ TcParams p) {
  using SP = SmemPlan<KIND, X3>;
  // no-swizzle operand layouts and bulk copies need 16-byte / 128-byte alignment only
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u);
  const uint32_t sm_u32 = ptx::smem_u32(sm);
  float* s_par = reinterpret_cast<float*>(sm + SP::PARAMS);
  float* s_xch = reinterpret_cast<float*>(sm + SP::XCH);
  const uint32_t bars = sm_u32 + SP::BARS;
  auto bar = [&](int slot) { return bars + 8u * slot; };
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(sm + SP::BARS + 8 * BAR_TMEM);
  volatile uint32_t* s_words = reinterpret_cast<volatile uint32_t*>(sm + SP::BARS);
  const uint32_t ticket = bars + 4u * WORD_TICKET;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (the control roles stay converged)
  const uint32_t rank = ptx::cluster_ctarank();   // 0 = leader (issues the pair's MMAs), 1 = peer
  const Program& P = p.prog;
  // Level 0: this CTA composites samples [s0, s0 + S_first) of its rays (a classic launch may cut the sample range into
  // segments, blockIdx.y); the fused kernel (n_levels == 2) walks all of level 0, samples the fine positions in-kernel
  // and walks all of level 1.
  const int n_levels = p.n_levels;
  const int s0 = (int)blockIdx.y * p.seg_len;
  const int S_first = min(p.seg_len, p.S_level[0] - s0);
  const int total_planes = S_first + (n_levels == 2 ? p.S_level[1] : 0);
  constexpr int NSTAGE = SP::NSTAGE;
  constexpr int E_OFF = off_E(X3), V_OFF = off_V(X3), P_OFF = off_P(X3), ONE_OFF = off_ONE(X3);

  // ---- one-time setup ---------------------------------------------------------------------------------
  if (tid == 0) {
    // leader's FULL: its own producer's arrive.expect_tx + the peer relay's "my half landed" arrive
    for (int i = 0; i < NSTAGE; ++i) { ptx::mbar_init(bar(BAR_FULL + i), rank == 0 ? 2 : 1); ptx::mbar_init(bar(BAR_EMPTY + i), 1); }
    // DEMPTY / CHUNK live in the leader and count the publishing warps of BOTH CTAs: 8 epilogue warps drain an
    // accumulator, one warp per lane quadrant (4) publishes a hidden chunk, the 4 encoder warps an aux chunk
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar(BAR_DFULL + i), 2); ptx::mbar_init(bar(BAR_DEMPTY + i), 16); }
    for (int i = 0; i < NUM_CHUNK_IDS; ++i) ptx::mbar_init(bar(BAR_CHUNK + i), 8);
    ptx::mbar_init(bar(BAR_EFREE), 2);   // both MMA issuers commit it
    ptx::mbar_init(bar(BAR_XW), 4);
    ptx::fence_mbar_init();
    s_words[WORD_TICKET] = 0u;
    uint32_t slot = 0;
    if (n_levels == 2) {
      // one scratch slot per resident CTA; n_slots >= number of SMs and one CTA fits per SM, so a free one exists
      for (uint32_t i = blockIdx.x % p.n_slots, tries = 0;; i = (i + 1 == (uint32_t)p.n_slots ? 0 : i + 1)) {
        const unsigned bit = 1u << (i & 31);
        if (!(atomicOr(p.slot_mask + (i >> 5), bit) & bit)) { slot = i; break; }
        if (++tries > (1u << 22)) {
          if (p.err_flag) atomicExch(p.err_flag, 9);
          __threadfence_system();
          __trap();
        }
      }
    }
    s_words[WORD_SLOT] = slot;
  }
  if (warp == 8) {
    ptx::tmem_alloc2(sm_u32 + SP::BARS + 8 * BAR_TMEM, 512);
    ptx::tmem_relinquish2();
  }
  {
    // head weights -> shared memory: [deform 3x128 + 4,] density 1x256 + 4, rgb 3x128 + 4 (PackedLayout::head_* order)
    // (both levels share the layout; the fused kernel reloads them between the levels)
    int off = 0;
    constexpr int NHEAD = KIND == AON_KIND_VANILLA ? 2 : 3;
    for (int h = 0; h < NHEAD; ++h) {
      const int n = (KIND == AON_KIND_AUTODECODER ? (h == 1 ? 256 : 384) : (h == 0 ? 256 : 384));
      const float* w = reinterpret_cast<const float*>(p.packed[0] + p.L.head_w[h]);
      const float* b = reinterpret_cast<const float*>(p.packed[0] + p.L.head_b[h]);
      for (int i = tid; i < n; i += TC_THREADS) s_par[off + i] = w[i] * (1.0f / SCALE_A);
      if (tid < 4) s_par[off + n + tid] = b[tid];
      off += n + 4;
    }
    // the constant "ones" operand chunk (K=16): columns 0..2 = 1.0 pick up the three bias pieces
    const uint16_t one = BF16 ? (uint16_t)0x3F80 : (uint16_t)0x3C00;
    for (int i = tid; i < 128 * 16; i += TC_THREADS) {
      const int row = i >> 4, k = i & 15;
      *reinterpret_cast<uint16_t*>(sm + ONE_OFF + (k >> 3) * 2048 + row * 16 + (k & 7) * 2) = k < 3 ? one : (uint16_t)0;
    }
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();   // barrier inits + TMEM allocation of both CTAs visible before any remote arrive / MMA
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);
  if (warp < 8) {   // both accumulators start zeroed (every MMA accumulates)
    const uint32_t zb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 256u;
#pragma unroll
    for (int i = 0; i < 8; ++i) ptx::tmem_st32_zero(zb + i * 32u);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  // the leader's CHUNK / DEMPTY / FULL barriers as shared::cluster addresses (valid from either CTA)
  const uint32_t lead_bars = ptx::mapa(bars, 0);
  auto lbar = [&](int slot) { return lead_bars + 8u * slot; };
  const long long t_start = clock64();
  auto mark = [&](int role, int s, int ui, int ev) {
    if (p.tl != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && s < 4) p.tl[((role * 4 + s) * MAX_UNITS + ui) * 4 + ev] = clock64() - t_start;
  };
  // optional wait-time accounting of the control roles (CTA 0 only, when a timeline buffer is attached): cycles spent in each
  // kind of wait, written behind the [3][4][MAX_UNITS][4] event stamps
  const bool prof = p.tl != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  constexpr int TL_STATS = 3 * 4 * MAX_UNITS * 4;
#define AON_TIMED(acc, stmt)                  \
  do {                                        \
    if (prof) {                               \
      const long long t__ = clock64();        \
      stmt;                                   \
      acc += clock64() - t__;                 \
    } else {                                  \
      stmt;                                   \
    }                                         \
  } while (0)
  // fused kernel: this CTA's scratch slot = coarse weights [128][S0] then fine t [128][S1]  (L2-resident: __ldcg / __stcg)
  const int S0L = p.S_level[0], S1L = p.S_level[1];
  float* slot_w = p.slots + (size_t)s_words[WORD_SLOT] * (size_t)(128 * (S0L + S1L));
  float* slot_t1 = slot_w + 128 * S0L;
  // In-kernel hierarchical sampling between the levels (A7; sampling.cuh), by the 12 epilogue + encoder warps: every
  // compositing weight of level 0 is in the slot, all MMAs of level 0 have completed (the epilogue warps waited for the
  // last accumulator), so the hidden-activation region of shared memory is free and serves as the per-warp scratch.
  auto sample_fine_positions = [&]() {
    level_bar_sync();
    float* scr = reinterpret_cast<float*>(sm + OFF_A) + warp * 384;
    static_assert(PDF_SCRATCH_FLOATS <= 384, "per-warp sampling scratch");
    for (int r = warp; r < 128; r += 12) {
      const long ray_r = (long)blockIdx.x * 128 + r;
      const long rl_r = ray_r < p.R ? ray_r : (long)p.R - 1;
      sample_pdf_ray<true>(p.t_vals + (p.t_stride ? rl_r * p.t_stride : 0), slot_w + r * S0L,
                           p.u ? p.u + rl_r * p.u_stride : nullptr, S0L, S1L - S0L, scr, slot_t1 + r * S1L, lane);
    }
    level_bar_sync();
  };

  if (warp >= 12) {
    ptx::setmaxnreg_dec<REGS_CONTROL>();
  if (warp == 12) {
    // ================================ weight producer (whole warp converged, one elected lane issues) ================================
    {
      uint32_t slot = 0, phase = 0;
      long long w_empty = 0;
      const long long t_role = clock64();
      for (int level = 0; level < n_levels; ++level) {
        const int S = level == 0 ? S_first : S1L;
        const char* fold_src = reinterpret_cast<const char*>(p.folded[level] + FOLD_STAGE_FLOAT0) + (size_t)rank * P.fold_half_bytes;
        for (int s = 0; s < S; ++s) {
          const char* src = p.packed[level] + (size_t)rank * P.stream_bytes;
          for (int ui = 0; ui < P.n_units; ++ui) {
            const Unit& u = P.u[ui];
            for (int c = 0; c < u.n_chunks; ++c) {
              const uint32_t w = u.ch[c];
              const uint32_t bytes = (uint32_t)stage_bytes(u, w, X3);
              const int nst = chunk_stages(w, X3);
              for (int i = 0; i < nst; ++i) {
                const char* from = src;
                if (KIND == AON_KIND_AUTODECODER && c == 0 && u.fold >= 0) from = fold_src + (size_t)u.fold * 16;
                AON_TIMED(w_empty, spin_bar(bar(BAR_EMPTY + slot), phase ^ 1));
                if (ptx::elect_one()) {
                  ptx::mbar_arrive_expect_tx(bar(BAR_FULL + slot), bytes);
                  ptx::bulk_g2s(sm_u32 + SP::RING + slot * SP::STAGE, from, bytes, bar(BAR_FULL + slot));
                }
                src += bytes;
                if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
              }
            }
          }
        }
      }
      if (prof && lane == 0) { p.tl[TL_STATS + 16] = w_empty; p.tl[TL_STATS + 17] = clock64() - t_role; }
    }
  } else if (warp == 14) {
    if (rank == 1) {
      // ================================ peer relay ================================
      // mbarrier.try_wait is CTA-local, so the leader cannot watch this CTA's FULL barriers: forward each
      // "my half of the stage landed" to the leader's FULL barrier (count 2) in stage order.
      uint32_t slot = 0, phase = 0;
      for (int s = 0; s < total_planes; ++s) {
        for (int ui = 0; ui < P.n_units; ++ui) {
          const Unit& u = P.u[ui];
          int nst = 0;
          for (int c = 0; c < u.n_chunks; ++c) nst += chunk_stages(u.ch[c], X3);
          for (int i = 0; i < nst; ++i) {
            spin_bar(bar(BAR_FULL + slot), phase);
            if (ptx::elect_one()) ptx::mbar_arrive_cluster(lbar(BAR_FULL + slot));
            if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  }
  if ((warp == 13 || warp == 14) && rank == 0) {
      // ================================ MMA issuers A (warp 14) and B (warp 13), leader CTA ================================
      // Issuing one weight stage (barrier test + fence + 2-3 tcgen05.mma + commit) costs one thread ~400 cycles of
      // serial latency (measured) -- as long as the MMAs of the stage take to execute -- so a single issuer cannot
      // keep the tensor pipe fed.  Two threads in different warps take alternate stages (global stage number
      // parity) and hand an issue ticket back and forth (ticket_wait above), so the MMAs enter the tensor core in stage
      // order: all waiting and the commits of the two threads overlap, only the issue itself is serialised.
      // tcgen05.commit tracks the committing thread's own MMAs, so both commit the accumulator-full (and
      // encoding-free) barriers, which count 2.
      const uint32_t me = warp == 14 ? 0u : 1u;
      constexpr uint32_t IDESC128 = ptx::idesc_f16(256, 128, BF16 ? 1 : 0);   // M = 256 over the pair
      constexpr uint32_t IDESC256 = ptx::idesc_f16(256, 256, BF16 ? 1 : 0);
      constexpr uint32_t A_LBO = (2048u >> 4) << 16;  // A operand: 128 rows x 16 B per k-group
      const uint32_t base16 = sm_u32 >> 4;
      const uint32_t ring16 = (sm_u32 + SP::RING) >> 4;
      uint32_t slot = 0, phase = 0, g = 0, n = 0;   // slot / phase / number of the NEXT stage of the common sequence
      long long w_full = 0, w_chunk = 0, w_dempty = 0, w_ticket = 0, n_own = 0;
      const long long t_role = clock64();
      auto advance = [&](uint32_t k) {
        n += k; slot += k;
        if (slot >= (uint32_t)NSTAGE) { slot -= NSTAGE; phase ^= 1; }
      };
      auto chunk_parity = [&](uint32_t w, int smp) { return (((uint32_t)smp & (w >> 21)) ^ (w >> 22)) & 1u; };
      for (int s = 0; s < total_planes; ++s) {
        for (int ui = 0; ui < P.n_units; ++ui, ++g) {
          const Unit& u = P.u[ui];
          const uint32_t b = g & 1;
          const uint32_t n128 = u.n128;
          const uint32_t idesc = n128 == 2 ? IDESC256 : IDESC128;
          const uint32_t b_lbo = (n128 * 64u) << 16;            // (N/2 rows * 16 B) >> 4 in the LBO field
          const uint32_t b_kstep16 = n128 * 128u;               // one K=16 step of B: N/2 * 32 B, >> 4
          const int n_chunks = u.n_chunks;
          const uint32_t d_tmem = tmem_base + b * 256u;
          // Every MMA of a unit accumulates (the epilogue leaves the accumulator zeroed); both issuers wait for the
          // accumulator to be drained before their first one.
          bool unit_wait = true;
          if ((n & 1u) == me) {
            // chunk 0 = bias rows x ones columns (one MMA in every mode)
            AON_TIMED(w_dempty, spin_bar(bar(BAR_DEMPTY + b), ((g >> 1) & 1) ^ 1));
            unit_wait = false;
            AON_TIMED(w_full, spin_bar(bar(BAR_FULL + slot), phase));
            ++n_own;
            ptx::tc_fence_after();
            if (me == 0 && lane == 0) mark(0, s, ui, 0);
            const uint32_t bd = (ring16 + slot * (SP::STAGE >> 4)) | b_lbo;
            AON_TIMED(w_ticket, ticket_wait(ticket, n));
            if (ptx::elect_one()) {
              ptx::mma2_f16_ss(d_tmem, mk_desc((base16 + (u.ch[0] & 0xFFFFu)) | A_LBO), mk_desc(bd), idesc, 1);
              ticket_set(ticket, n + 1u);
              ptx::mma_commit2(bar(BAR_EMPTY + slot), 3);
            }
          }
          advance(1);
          uint32_t w = u.ch[1];
          for (int c = 1; c < n_chunks; ++c) {
            const uint32_t wn = u.ch[c + 1 < n_chunks ? c + 1 : c];
            const uint32_t nst = X3 ? ((w >> 23) & 7u) : 1u;
            const uint32_t ks = (me ^ n) & 1u;          // which of the chunk's stages is mine (if it has two)
            if (nst == 2u || ks == 0u) {
              uint32_t my_slot = slot + ks, my_phase = phase;
              if (my_slot >= (uint32_t)NSTAGE) { my_slot -= NSTAGE; my_phase ^= 1; }
              const uint32_t id = (w >> 16) & 15u;
              if (unit_wait) { AON_TIMED(w_dempty, spin_bar(bar(BAR_DEMPTY + b), ((g >> 1) & 1) ^ 1)); unit_wait = false; }
              ++n_own;
              if (w & (1u << 20)) {   // first use of this generation of the chunk: wait until both CTAs have published it
                AON_TIMED(w_chunk, spin_bar(bar(BAR_CHUNK + id), chunk_parity(w, s)));
                if (w >> 31) AON_TIMED(w_chunk, spin_bar(bar(BAR_CHUNK + id + 1), chunk_parity(w, s)));   // merged pair
              }
              AON_TIMED(w_full, spin_bar(bar(BAR_FULL + my_slot), my_phase));
              ptx::tc_fence_after();
              const uint32_t bd = (ring16 + my_slot * (SP::STAGE >> 4)) | b_lbo;
              const uint32_t a_hi = ((base16 + (w & 0xFFFFu)) | A_LBO) + (X3 ? ks * 256u : 0u);
              AON_TIMED(w_ticket, ticket_wait(ticket, n + ks));
              if (ptx::elect_one()) {
                if (X3) {   // hi rows; lo rows follow at +N/2*32 B
                  const uint32_t a_lo = a_hi + (((w >> 26) & 31u) << 8);  // (lo offset >> 12) << 12 >> 4
                  ptx::mma2_f16_ss(d_tmem, mk_desc(a_hi), mk_desc(bd), idesc, 1);
                  ptx::mma2_f16_ss(d_tmem, mk_desc(a_lo), mk_desc(bd), idesc, 1);
                  ptx::mma2_f16_ss(d_tmem, mk_desc(a_hi), mk_desc(bd + b_kstep16), idesc, 1);
                } else {
                  const uint32_t nk = (w >> 23) & 7u;   // 2 (K = 32) or 4 (merged pair, K = 64)
                  ptx::mma2_f16_ss(d_tmem, mk_desc(a_hi), mk_desc(bd), idesc, 1);
                  ptx::mma2_f16_ss(d_tmem, mk_desc(a_hi + 256u), mk_desc(bd + b_kstep16), idesc, 1);
                  if (nk == 4u) {
                    ptx::mma2_f16_ss(d_tmem, mk_desc(a_hi + 512u), mk_desc(bd + 2u * b_kstep16), idesc, 1);
                    ptx::mma2_f16_ss(d_tmem, mk_desc(a_hi + 768u), mk_desc(bd + 3u * b_kstep16), idesc, 1);
                  }
                }
                ticket_set(ticket, n + ks + 1u);
                ptx::mma_commit2(bar(BAR_EMPTY + my_slot), 3);
              }
            }
            advance(nst);
            w = wn;
          }
          if (ptx::elect_one()) {
            ptx::mma_commit2(bar(BAR_DFULL + b), 3);
            if (u.last_e_use) ptx::mma_commit2(bar(BAR_EFREE), 3);
          }
          if (me == 0 && lane == 0) mark(0, s, ui, 3);
          if (!X3 && u.n_hidden) {
            // Parity waits are only sound if this warp never falls a whole generation behind a barrier it will test
            // later: in the one-pass modes a chunk entry is ONE stage, so a warp only waits for the chunks whose stage it
            // owns; it observes the generation of the unit's last hidden chunks here (each epilogue warp publishes in order:
            // last even + last odd chunk imply the rest).  In f16x3 every chunk has one stage per warp: both wait for all.
            const uint32_t nh = u.n_hidden, par = chunk_parity(u.ch[u.last_hidden], s);
            spin_bar(bar(BAR_CHUNK + nh - 1), par);
            spin_bar(bar(BAR_CHUNK + nh - 2), par);
          }
        }
      }
      if (prof && lane == 0) {
        long long* st = p.tl + TL_STATS + me * 8;
        st[0] = w_full; st[1] = w_chunk; st[2] = w_dempty; st[3] = w_ticket; st[4] = clock64() - t_role; st[5] = n_own;
      }
  }
  } else if (warp >= 8) {
    ptx::setmaxnreg_dec<REGS_ENCODER>();
    // ================================ encoder: sample positions -> operand chunks ================================
    const int row = tid - 256;   // one row per encoder thread
    float o[3], d[3], vd[3];
    load_ray(p, row, o, d, vd);
    const long ray = (long)blockIdx.x * 128 + row;
    const long rl = ray < p.R ? ray : (long)p.R - 1;
    auto publish = [&](int id) {
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(lbar(BAR_CHUNK + id));
    };
    // view-direction encoding, once per tile (model.py:174: pos_enc(viewdirs, 0, 4))
    encode_store<4, 32, X3, BF16>(sm + V_OFF, LO_V, row, vd);
    publish(CH_V);
    int pl = 0;   // plane number over both levels (barrier parities continue across the level boundary)
    for (int level = 0; level < n_levels; ++level) {
      const int S = level == 0 ? S_first : S1L;
      const int SG = level == 0 ? S0L : S1L;
      const int sb = level == 0 ? s0 : 0;
      if (level == 1) sample_fine_positions();
      const float* tv = level == 0 ? p.t_vals + (p.t_stride ? rl * p.t_stride : 0) : slot_t1 + row * S1L;
      float t = __ldcg(tv + sb);
      for (int s = 0; s < S; ++s, ++pl) {
        const float t_next = (sb + s + 1 < SG) ? __ldcg(tv + sb + s + 1) : 0.f;
        // cast_rays (helper.py:25-26)
        float x[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) x[k] = __fadd_rn(o[k], __fmul_rn(t, d[k]));
        if (pl > 0) wait_bar(bar(BAR_EFREE), (uint32_t)((pl - 1) & 1), p.err_flag, 7);
        if (tid == 256) mark(2, pl, 0, 0);
        if (KIND == AON_KIND_AUTODECODER) {
          // raw position -> operand chunk P (deformation MLP input; model_autodecoder.py:196-198)
          constexpr int PK = X3 ? 16 : 32;
#pragma unroll
          for (int k = 0; k < 3; ++k) put16<X3, BF16>(sm + P_OFF, LO_P, row, k, x[k]);
#pragma unroll
          for (int k = 3; k < PK; ++k) put16<X3, BF16>(sm + P_OFF, LO_P, row, k, 0.f);
          publish(CH_P);
          // warped position x' = x + deformation(x), handed over by the epilogue warps as fp32 in the
          // (by then consumed) P region
          wait_bar(bar(BAR_XW), (uint32_t)(pl & 1), p.err_flag, 8);
          const float* xw = reinterpret_cast<const float*>(sm + P_OFF);
#pragma unroll
          for (int k = 0; k < 3; ++k) x[k] = xw[128 * k + row];
        }
        encode_store<10, 64, X3, BF16>(sm + E_OFF, LO_E, row, x);
        if (TRAIN) {   // this thread's row of the encoding operand (its own shared-memory writes) -> packed plane in HBM
          const size_t tile = (size_t)blockIdx.x * SG + (sb + s);
#pragma unroll
          for (int kg = 0; kg < 8; ++kg) {
            p.dump.e_hi[(tile * 8 + kg) * 128 + row] = *reinterpret_cast<const uint4*>(sm + E_OFF + kg * 2048 + row * 16);
            if (X3) p.dump.e_lo[(tile * 8 + kg) * 128 + row] = *reinterpret_cast<const uint4*>(sm + E_OFF + LO_E + kg * 2048 + row * 16);
          }
        }
        publish(CH_E0);
        publish(CH_E0 + 1);
        if (tid == 256) mark(2, pl, 0, 1);
        t = t_next;
      }
    }
  } else {
    ptx::setmaxnreg_inc<REGS_EPILOGUE>();
    // ================================ epilogue / per-ray state ================================
    const int quad = warp & 3;              // TMEM lane quadrant of this warp
    const int half = warp >> 2;             // 0: even chunks + owner of the per-ray state; 1: odd chunks
    const int row = quad * 32 + lane;       // ray within the tile == TMEM lane
    const bool owner = half == 0;
    const long ray = (long)blockIdx.x * 128 + row;
    const bool valid = ray < p.R;
    const long rl = valid ? ray : (long)p.R - 1;
    float ro[3], rd[3], rv[3];
    load_ray(p, row, ro, rd, rv);
    const float ox = ro[0], oy = ro[1], oz = ro[2], dx = rd[0], dy = rd[1], dz = rd[2];
    const float dnorm = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);

    // head weights / biases in shared memory
    const float* hw_def = s_par;                                             // auto-decoder only
    const float* hw_sig = s_par + (KIND == AON_KIND_AUTODECODER ? 388 : 0);
    const float* hw_rgb = hw_sig + 260;

    uint32_t g = 0;
    int pl = 0;
    for (int level = 0; level < n_levels; ++level) {
      const int S = level == 0 ? S_first : S1L;
      const int SG = level == 0 ? S0L : S1L;
      const int sb = level == 0 ? s0 : 0;
      if (level == 1) {
        sample_fine_positions();
        // the fine MLP's head weights replace the coarse ones (all 12 warps are past the level barrier: nobody reads them)
        int off = 0;
        constexpr int NHEAD = KIND == AON_KIND_VANILLA ? 2 : 3;
        for (int h = 0; h < NHEAD; ++h) {
          const int n = (KIND == AON_KIND_AUTODECODER ? (h == 1 ? 256 : 384) : (h == 0 ? 256 : 384));
          const float* w = reinterpret_cast<const float*>(p.packed[1] + p.L.head_w[h]);
          const float* b = reinterpret_cast<const float*>(p.packed[1] + p.L.head_b[h]);
          for (int i = tid; i < n; i += 256) s_par[off + i] = w[i] * (1.0f / SCALE_A);
          if (tid < 4) s_par[off + n + tid] = b[tid];
          off += n + 4;
        }
        __syncwarp();
        asm volatile("bar.sync 6, 256;" ::: "memory");
      }
      const float* tv = level == 0 ? p.t_vals + (p.t_stride ? rl * p.t_stride : 0) : slot_t1 + row * S1L;
      const bool last_level = level == n_levels - 1;
      float trans = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, cdepth = 0.f, cacc = 0.f;
      float t_cur = __ldcg(tv + sb);
    for (int s = 0; s < S; ++s, ++pl) {
      const float t_next = (sb + s + 1 < SG) ? __ldcg(tv + sb + s + 1) : 0.f;
      float sig = 0.f, h0 = 0.f, h1 = 0.f, h2 = 0.f;  // head accumulators (density; rgb / deformation)

      for (int ui = 0; ui < P.n_units; ++ui, ++g) {
        const Unit& u = P.u[ui];
        const uint32_t b = g & 1;
        const int epi = u.epi;
        const bool relu = u.relu != 0;
        const int n_mine = u.n128 * 2;  // my 32-column chunks of this unit's output: cc = 2 j + half
        if (epi == EPI_RGB || epi == EPI_DEFORM) { h0 = h1 = h2 = 0.f; }
        const float* hw = epi == EPI_STORE_SIGMA ? hw_sig : (epi == EPI_RGB ? hw_rgb : hw_def);
        unsigned char* out_base = sm + OFF_A + row * 16;
        wait_bar(bar(BAR_DFULL + b), (g >> 1) & 1, p.err_flag, 5);
        ptx::tc_fence_after();
        if (tid == 0) mark(1, pl, ui, 0);
        const uint32_t d_addr = lane_base + b * 256u + (uint32_t)half * 32u;

        uint32_t r[2][32];
        ptx::tmem_ld32(d_addr, r[0]);
#pragma unroll 2
        for (int j = 0; j < n_mine; ++j) {
          const int cc = 2 * j + half;
          float v[32];
          // the two register buffers alternate; written so that all indexing stays static
          if ((j & 1) == 0) {
            tmem_ld_wait_dep(r[0]);
            if (j + 1 < n_mine) ptx::tmem_ld32(d_addr + (uint32_t)(j + 1) * 64u, r[1]);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[0][i]) * (1.0f / SCALE_W);
          } else {
            tmem_ld_wait_dep(r[1]);
            if (j + 1 < n_mine) ptx::tmem_ld32(d_addr + (uint32_t)(j + 1) * 64u, r[0]);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[1][i]) * (1.0f / SCALE_W);
          }
          // leave the columns just read zeroed: the next unit that uses this accumulator only accumulates
          ptx::tmem_st32_zero(d_addr + (uint32_t)j * 64u);
          if (j == n_mine - 1) {
            // accumulator drained and cleared (my part): hand the TMEM buffer back to the MMA issuers
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lbar(BAR_DEMPTY + b));
          }
          if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && pl == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) p.dbg[((size_t)ui * 128 + row) * 256 + cc * 32 + i] = v[i] * (1.0f / SCALE_A);
          }
          // training forward: this thread's 32 outputs go to HBM once, as 4 x 16-byte rows of the packed plane (+ a ReLU mask
          // word: bit i = pre-activation i is not negative, collected from the sign bits with one funnel shift per value)
          uint4* gh = nullptr;
          uint4* gl = nullptr;
          if (TRAIN) {
            const size_t tile = (size_t)blockIdx.x * SG + (sb + s);
            const size_t g0 = (tile * (size_t)(u.n128 * 16) + cc * 4) * 128 + row;
            gh = p.dump.hi[ui] + g0;
            if (X3) gl = p.dump.lo[ui] + g0;
            if (relu && p.dump.bits[ui] != nullptr && !(p.dump.dbg_flags & 2)) {
              uint32_t neg = 0;
#pragma unroll
              for (int i = 31; i >= 0; --i) neg = __funnelshift_l(__float_as_uint(v[i]), neg, 1);
              p.dump.bits[ui][(tile * 128 + row) * (size_t)(u.n128 * 4) + cc] = ~neg;
            }
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
            if (TRAIN) {   // the head's input layer is needed by the backward (wgrad operand of the head, ReLU mask)
#pragma unroll
              for (int kg = 0; kg < 4; ++kg) {
                uint4 hi, lo;
                pack8<X3, BF16>(v + kg * 8, hi, lo);
                gh[kg * 128] = hi;
                if (X3) gl[kg * 128] = lo;
              }
            }
          } else {
            // next layer's A operand: 32 hidden features of this row -> chunk cc
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
              uint4 hi, lo;
              pack8<X3, BF16>(v + kg * 8, hi, lo);
              *reinterpret_cast<uint4*>(out_base + cc * 8192 + kg * 2048) = hi;
              if (X3) *reinterpret_cast<uint4*>(out_base + LO_A + cc * 8192 + kg * 2048) = lo;
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lbar(BAR_CHUNK + cc));
            if (tid == 0 && j == 0) mark(1, pl, ui, 1);
            if (TRAIN) {
              // The finished chunk (32 features x 128 rows: 8 KB hi + 8 KB lo) sits in shared memory in exactly the layout of
              // its slice of the packed plane in HBM: once the four warps that wrote it have met (named barrier 7 / 8 of this
              // half), ONE thread hands it to the TMA engine as two bulk copies -- no per-thread global stores in the
              // epilogue.  The region is overwritten by the next unit's epilogue, >= 1 such barrier later: the issuing thread
              // waits for the previous chunk's shared-memory reads before it arrives at the barrier.
              const bool issuer = quad == 0 && lane == 0;
              if (issuer) ptx::bulk_wait_read0();
              __syncwarp();
              if (!(p.dump.dbg_flags & 4)) asm volatile("bar.sync %0, 128;" ::"r"(7 + half) : "memory");
              if (issuer && !(p.dump.dbg_flags & 1)) {
                const uint32_t src = sm_u32 + OFF_A + cc * 8192;
                // evict_first: 4 GB of planes stream out per level through the L2 that serves the weight stream of every sample
                // plane; without the hint the level takes 1.53 ms instead of 1.27 ms (measured; evict_last on the weights: no gain)
                // (the one-pass mode writes half the bytes and has a deeper weight ring: the hint does not pay there -- 0.71 vs 0.74 ms)
                if (X3) {
                  const uint64_t pol = ptx::l2_policy_evict_first();
                  ptx::bulk_s2g_hint(gh - row, src, 8192, pol);
                  ptx::bulk_s2g_hint(gl - row, src + LO_A, 8192, pol);
                } else {
                  ptx::bulk_s2g(gh - row, src, 8192);
                }
                ptx::bulk_commit();
              }
            }
          }
        }
        if (tid == 0) mark(1, pl, ui, 2);

        // head partial sums of the odd warp -> even warp (owner of the per-ray state)
        if (epi != EPI_STORE) {
          if (!owner) {
            if (epi == EPI_STORE_SIGMA) s_xch[row] = sig;
            else { s_xch[128 + row] = h0; s_xch[256 + row] = h1; s_xch[384 + row] = h2; }
            __threadfence_block();
            pair_bar_arrive(1 + quad);
          } else {
            pair_bar_sync(1 + quad);
            if (epi == EPI_STORE_SIGMA) sig += s_xch[row];
            else { h0 += s_xch[128 + row]; h1 += s_xch[256 + row]; h2 += s_xch[384 + row]; }
          }
        }

        if (KIND == AON_KIND_AUTODECODER && epi == EPI_DEFORM && owner) {
          // model_autodecoder.py:203: x' = deformation_layer(h) + pos  (pos via cast_rays, helper.py:25-26)
          const float* hb = hw_def + 384;
          float* xw = reinterpret_cast<float*>(sm + P_OFF);
          xw[row] = __fadd_rn(h0 + hb[0], __fadd_rn(ox, __fmul_rn(t_cur, dx)));
          xw[128 + row] = __fadd_rn(h1 + hb[1], __fadd_rn(oy, __fmul_rn(t_cur, dy)));
          xw[256 + row] = __fadd_rn(h2 + hb[2], __fadd_rn(oz, __fmul_rn(t_cur, dz)));
          if (TRAIN) {
            const size_t m = ((size_t)blockIdx.x * SG + (sb + s)) * 128 + row;
            p.dump.warped[3 * m + 0] = xw[row]; p.dump.warped[3 * m + 1] = xw[128 + row]; p.dump.warped[3 * m + 2] = xw[256 + row];
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar(BAR_XW));
        }
      }

      // ---- activations + alpha compositing of sample s (helper.py:157-195) ----
      if (TRAIN) {
        // training forward: the raw network outputs leave the kernel; activations + compositing (and their adjoints) are
        // aon_composite / aon_composite_backward on [R,S,4]
        if (owner && valid) {
          const float* hb = hw_rgb + 384;
          p.dump.raw[ray * SG + sb + s] = make_float4(h0 + hb[0], h1 + hb[1], h2 + hb[2], sig + hw_sig[256]);
        }
      } else if (owner) {
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
        const float delta = (sb + s + 1 < SG) ? __fsub_rn(t_next, t_cur) : 1e10f;
        const float dist = __fmul_rn(delta, dnorm);
        const float alpha = __fsub_rn(1.0f, expf(__fmul_rn(-sigma, dist)));
        if (p.n_seg > 1) {
          // sample-segmented launch: the ray's samples are composited in order by composite_samples_kernel
          if (valid) p.seg_samples[ray * SG + sb + s] = make_float4(alpha, rr, gg, bb);
        } else {
          const float w = __fmul_rn(alpha, trans);
          cr = fmaf(w, rr, cr); cg = fmaf(w, gg, cg); cb = fmaf(w, bb, cb);
          cdepth = fmaf(w, t_cur, cdepth);
          cacc += w;
          trans = __fmul_rn(trans, __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f));
          if (n_levels == 2) {
            if (level == 0) __stcg(slot_w + row * S0L + s, w);
          } else if (p.weights && valid) {
            p.weights[ray * SG + s] = w;
          }
        }
      }
      t_cur = t_next;
    }

      if (!TRAIN && valid && owner && p.n_seg == 1) {
        if (isnan(cdepth)) cdepth = INFINITY;  // helper.py:179 nan_to_num(depth, nan=inf)
        else if (isinf(cdepth)) cdepth = cdepth > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
        if (p.white_bkgd) {
          const float bg = __fsub_rn(1.0f, cacc);
          cr += bg; cg += bg; cb += bg;
        }
        float* o5 = last_level ? p.out5 : p.coarse_out5;
        if (o5 != nullptr) {
          o5[5 * ray + 0] = cr; o5[5 * ray + 1] = cg; o5[5 * ray + 2] = cb; o5[5 * ray + 3] = cacc; o5[5 * ray + 4] = cdepth;
        } else if (last_level) {
          p.comp_rgb[3 * ray + 0] = cr; p.comp_rgb[3 * ray + 1] = cg; p.comp_rgb[3 * ray + 2] = cb;
          p.acc[ray] = cacc;
          p.depth[ray] = cdepth;
        }
      }
    }
  }

  if (TRAIN && (tid == 0 || tid == 128)) ptx::bulk_wait0();   // the two bulk-store issuers: all activation planes have landed
  // ---- teardown ----
  ptx::tc_fence_before();
  ptx::cluster_sync_all();   // no CTA of the pair may exit (or free TMEM) while its partner can still touch it
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc2(tmem_base, 512);
  }
  if (tid == 0 && n_levels == 2) {
    const uint32_t slot = s_words[WORD_SLOT];
    atomicAnd(p.slot_mask + (slot >> 5), ~(1u << (slot & 31)));
  }
}

// Sample-segmented launches (small ray batches, the tail wave of a large one) leave (alpha, r, g, b) of every sample in a
// scratch tensor; this kernel composites them per ray in sample order with exactly the operation sequence of the fused
// kernel's epilogue (helper.py:157-195), so a segmented render is bit-identical to an unsegmented one.
__global__ void composite_samples_kernel(const float4* __restrict__ smp, const float* __restrict__ t_vals, long t_stride, int R, int S,
                                         int white_bkgd, float* __restrict__ comp_rgb, float* __restrict__ acc,
                                         float* __restrict__ depth, float* __restrict__ out5, float* __restrict__ weights) {
  const long ray = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= R) return;
  const float* tv = t_vals + (t_stride ? ray * t_stride : 0);
  float trans = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, cdepth = 0.f, cacc = 0.f;
  for (int s = 0; s < S; ++s) {
    const float4 a = smp[ray * S + s];
    const float t_cur = tv[s];
    const float w = __fmul_rn(a.x, trans);
    cr = fmaf(w, a.y, cr); cg = fmaf(w, a.z, cg); cb = fmaf(w, a.w, cb);
    cdepth = fmaf(w, t_cur, cdepth);
    cacc += w;
    trans = __fmul_rn(trans, __fadd_rn(__fsub_rn(1.0f, a.x), 1e-10f));
    if (weights != nullptr) weights[ray * S + s] = w;
  }
  if (isnan(cdepth)) cdepth = INFINITY;
  else if (isinf(cdepth)) cdepth = cdepth > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  if (white_bkgd) {
    const float bg = __fsub_rn(1.0f, cacc);
    cr += bg; cg += bg; cb += bg;
  }
  if (out5 != nullptr) {
    out5[5 * ray + 0] = cr; out5[5 * ray + 1] = cg; out5[5 * ray + 2] = cb; out5[5 * ray + 3] = cacc; out5[5 * ray + 4] = cdepth;
  } else {
    comp_rgb[3 * ray + 0] = cr; comp_rgb[3 * ray + 1] = cg; comp_rgb[3 * ray + 2] = cb;
    acc[ray] = cacc;
    depth[ray] = cdepth;
  }
}

// Samples-per-CTA split: small ray batches (the reference's 3840-ray chunks = 15 CTA pairs on 74 pair slots) leave most
// SMs idle; cutting each ray's sample range into n segments gives n times as many CTAs.  Cost model per wave:
// ceil(S / n) sample planes + ~3 planes of start-up (view encoding, pipeline fill, combine).
static int choose_segments(int pairs, int S, int pair_slots) {
  int best = 1;
  long best_cost = -1;
  for (int n = 1; n <= 16 && n <= S / 4; ++n) {
    const long waves = ((long)pairs * n + pair_slots - 1) / pair_slots;
    const long cost = waves * ((S + n - 1) / n + 3);
    if (best_cost < 0 || cost * 100 < best_cost * 90) { best = n; best_cost = cost; }   // needs a > 10 % win over fewer segments
  }
  return best;
}

// ---- host side -------------------------------------------------------------------------------------------
template <int KIND, bool X3, bool BF16, bool TRAIN = false>
static int launch(const TcParams& p, int grid, int grid_y, cudaStream_t st) {
  using SP = SmemPlan<KIND, X3>;
  auto kern = render_tc_kernel<KIND, X3, BF16, TRAIN>;
  AON_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SP::TOTAL));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, (unsigned)grid_y);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SP::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AON_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  AON_LAUNCH_CHECK();
  return AON_OK;
}

static int launch_any(const TcParams& p, int kind, int precision, int grid, int grid_y, cudaStream_t st) {
  const bool van = kind == AON_KIND_VANILLA;
  switch (precision) {
    case AON_PREC_TC_F16X3:
      return van ? launch<AON_KIND_VANILLA, true, false>(p, grid, grid_y, st) : launch<AON_KIND_AUTODECODER, true, false>(p, grid, grid_y, st);
    case AON_PREC_TC_F16:
      return van ? launch<AON_KIND_VANILLA, false, false>(p, grid, grid_y, st) : launch<AON_KIND_AUTODECODER, false, false>(p, grid, grid_y, st);
    case AON_PREC_TC_BF16:
      return van ? launch<AON_KIND_VANILLA, false, true>(p, grid, grid_y, st) : launch<AON_KIND_AUTODECODER, false, true>(p, grid, grid_y, st);
    default:
      set_error("bad precision %d", precision);
      return AON_E_ARG;
  }
}

static int launch_train(const TcParams& p, int kind, int precision, int grid, int grid_y, cudaStream_t st) {
  const bool van = kind == AON_KIND_VANILLA;
  switch (precision) {
    case AON_PREC_TC_F16X3:
      return van ? launch<AON_KIND_VANILLA, true, false, true>(p, grid, grid_y, st) : launch<AON_KIND_AUTODECODER, true, false, true>(p, grid, grid_y, st);
    case AON_PREC_TC_F16:
      return van ? launch<AON_KIND_VANILLA, false, false, true>(p, grid, grid_y, st) : launch<AON_KIND_AUTODECODER, false, false, true>(p, grid, grid_y, st);
    default:
      set_error("training forward: precision must be f16x3 or f16 (got %d)", precision);
      return AON_E_ARG;
  }
}

static int device_info(int* dev, int* sms) {
  int major = 0;
  AON_CUDA_CHECK(cudaGetDevice(dev));
  AON_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, *dev));
  if (major != 10) {
    set_error("tensor-core precision modes need an sm_100 device (found compute capability %d.x)", major);
    return AON_E_UNSUPPORTED;
  }
  *sms = 148;
  cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, *dev);
  return AON_OK;
}

// Classic single-level launch over rays [0, R) of the arrays in `p` (already offset by the caller) with n_seg sample
// segments per ray; seg_buf (>= R * S float4) is only touched when n_seg > 1.
static int render_range(TcParams p, int kind, int precision, int R, int n_seg, float4* seg_buf, cudaStream_t st) {
  const int S = p.S_level[0];
  p.R = R;
  const int grid = ((R + 255) / 256) * 2;   // CTA pairs; an odd last tile leaves the peer with no valid rays
  p.n_seg = n_seg;
  p.seg_len = (S + n_seg - 1) / n_seg;
  p.seg_samples = n_seg > 1 ? seg_buf : nullptr;
  int rc = launch_any(p, kind, precision, grid, n_seg, st);
  if (n_seg > 1 && rc == AON_OK) {
    composite_samples_kernel<<<(R + 127) / 128, 128, 0, st>>>(seg_buf, p.t_vals, p.t_stride, R, S, p.white_bkgd, p.comp_rgb, p.acc,
                                                              p.depth, p.out5, p.weights);
    g_launches++;
    if (cudaGetLastError() != cudaSuccess) { set_error("composite_samples_kernel launch failed"); rc = AON_E_CUDA; }
  }
  return rc;
}

static void fill_common(TcParams& p, int kind, int precision, const AonRenderOpts* opts) {
  memset(&p, 0, sizeof(p));
  p.prog = build_program(kind, precision);
  p.L = layout_tc(kind, precision);
  p.n_seg = 1;
  if (opts) { p.dbg = opts->dbg; p.err_flag = opts->err_flag; p.tl = opts->timeline; }
}

size_t seg_bytes_tc(int R, int S) {
  // worst case of the sample-segmented part of a launch: every ray of a small batch, or the last partial wave (< 80 CTA pairs)
  const long rays = R < MAX_SLOTS / 2 * 256 ? R : MAX_SLOTS / 2 * 256;
  return (size_t)rays * S * sizeof(float4);
}

// Split of a fused render of R rays: returns the number of rays (a multiple of 256, or R) the fused kernel takes as full
// waves of CTA pairs; the remainder goes through the sample-segmented three-launch path.  A remainder that would not be
// segmented anyway (it nearly fills a wave) stays in the fused launch.
int tail_split_tc(int R, int sms, int* n_seg_coarse, int* n_seg_fine) {
  const int pairs = (R + 255) / 256, slots = sms / 2;
  const int rem = pairs % slots;
  *n_seg_coarse = *n_seg_fine = 1;
  if (rem == 0) return R;
  const int nc = choose_segments(rem, N_COARSE, slots), nf = choose_segments(rem, N_TOTAL, slots);
  if (nc == 1 && nf == 1) return R;
  *n_seg_coarse = nc;
  *n_seg_fine = nf;
  return (pairs - rem) * 256;
}

int render_level_tc(int kind, int precision, const void* packed, const float* folded, const float* rays_o,
                    const float* rays_d, const float* viewdirs, const float* t_vals, long t_stride, int R, int S,
                    int white_bkgd, float* comp_rgb, float* acc, float* depth, float* out5, float* weights, void* workspace,
                    size_t workspace_bytes, const AonRenderOpts* opts, cudaStream_t st) {
  int dev = 0, sms = 148;
  int rc = device_info(&dev, &sms);
  if (rc != AON_OK) return rc;
  TcParams p;
  fill_common(p, kind, precision, opts);
  p.packed[0] = (const char*)packed;
  p.folded[0] = folded;
  p.S_level[0] = S;
  p.n_levels = 1;
  p.rays_o = rays_o; p.rays_d = rays_d; p.viewdirs = viewdirs;
  p.t_vals = t_vals; p.t_stride = t_stride;
  p.R = R; p.white_bkgd = white_bkgd;
  p.comp_rgb = comp_rgb; p.acc = acc; p.depth = depth; p.out5 = out5; p.weights = weights;
  float4* seg_buf = (float4*)workspace;
  const size_t seg_cap = workspace_bytes / sizeof(float4);
  const int pairs = (R + 255) / 256, slots = sms / 2;
  const int force = opts ? opts->force_segments : 0;
  auto need = [&](int rays, int n_seg) -> bool { return n_seg <= 1 || (size_t)rays * S <= seg_cap; };
  if (force > 0) {
    const int n = force <= S ? force : 1;
    if (!need(R, n)) { set_error("aon_render_level: workspace too small for %d sample segments (see aon_workspace_bytes)", n); return AON_E_SIZE; }
    return render_range(p, kind, precision, R, n, seg_buf, st);
  }
  if (pairs <= slots || (opts && opts->no_tail_split)) {
    int n = choose_segments(pairs, S, slots);
    if (!need(R, n)) n = 1;   // no workspace: unsegmented (slower on small batches, same values)
    return render_range(p, kind, precision, R, n, seg_buf, st);
  }
  // Large batch = several waves of ray tiles over the pair slots.  The last, partly filled wave (640x480: 1200 tiles on 74
  // slots = 16.2 waves) would hold all but a few SMs idle for a whole tile time: render the full waves unsplit and the
  // remainder tiles in a second launch with sample segments, so the remainder spreads over every SM.
  const int rem = pairs % slots;
  int n_tail = rem ? choose_segments(rem, S, slots) : 1;
  const int R_main = (pairs - rem) * 256;
  if (n_tail > 1 && !need(R - R_main, n_tail)) n_tail = 1;
  if (n_tail == 1) return render_range(p, kind, precision, R, 1, seg_buf, st);
  rc = render_range(p, kind, precision, R_main, 1, seg_buf, st);
  if (rc != AON_OK) return rc;
  p.rays_o += 3 * (size_t)R_main; p.rays_d += 3 * (size_t)R_main; p.viewdirs += 3 * (size_t)R_main;
  p.t_vals += (size_t)R_main * t_stride;
  if (p.out5) p.out5 += 5 * (size_t)R_main;
  else { p.comp_rgb += 3 * (size_t)R_main; p.acc += R_main; p.depth += R_main; }
  if (p.weights) p.weights += (size_t)R_main * S;
  return render_range(p, kind, precision, R - R_main, n_tail, seg_buf, st);
}

// Training forward of ONE level over rays [0, R): the MLP chain of all R * S samples in one launch, every layer output written
// once (TrainDump), raw outputs in ray-major order.  Sample segments spread a small ray batch over all SMs (no compositing
// here, so segments need no scratch).  Row tiles: ((R + 255) / 256 * 2) * S.
int forward_train_tc(int kind, int precision, const void* packed, const float* folded, const float* rays_o, const float* rays_d,
                     const float* viewdirs, const float* t_vals, long t_stride, int R, int S, const TrainDump& dump, cudaStream_t st) {
  int dev = 0, sms = 148;
  int rc = device_info(&dev, &sms);
  if (rc != AON_OK) return rc;
  TcParams p;
  fill_common(p, kind, precision, nullptr);
  p.dump = dump;
  p.packed[0] = (const char*)packed;
  p.folded[0] = folded;
  p.S_level[0] = S;
  p.n_levels = 1;
  p.rays_o = rays_o; p.rays_d = rays_d; p.viewdirs = viewdirs;
  p.t_vals = t_vals; p.t_stride = t_stride;
  p.R = R;
  const int pairs = (R + 255) / 256, slots = sms / 2;
  // samples per CTA: as many segments as it takes to fill the pair slots (cost model of choose_segments)
  int n_seg = 1;
  {
    long best = -1;
    for (int n = 1; n <= S; ++n) {
      const long waves = ((long)pairs * n + slots - 1) / slots;
      const long cost = waves * ((S + n - 1) / n + 2);
      if (best < 0 || cost < best) { best = cost; n_seg = n; }
    }
  }
  p.n_seg = n_seg;
  p.seg_len = (S + n_seg - 1) / n_seg;
  n_seg = (S + p.seg_len - 1) / p.seg_len;     // no empty trailing segment
  p.n_seg = n_seg;
  return launch_train(p, kind, precision, pairs * 2, n_seg, st);
}

// The fused image kernel over rays [0, R): coarse level, in-kernel hierarchical sampling, fine level in ONE launch; no
// per-sample tensor reaches HBM.  slots / slot_mask live in the caller's workspace (mask zeroed by the caller).
int render_fused_tc(int kind, int precision, const void* packed_c, const void* packed_f, const float* folded_c,
                    const float* folded_f, const float* rays_o, const float* rays_d, const float* viewdirs, const Cam* cam,
                    int H, int W, float focal, long ray0, const float* t0, long t0_stride, const float* u, long u_stride,
                    int R, int S0, int S1, int white_bkgd, float* out5, float* coarse_out5, float* slots, unsigned* slot_mask,
                    int n_slots, const AonRenderOpts* opts, cudaStream_t st) {
  int dev = 0, sms = 148;
  int rc = device_info(&dev, &sms);
  if (rc != AON_OK) return rc;
  if (sms > n_slots) { set_error("fused render: device has %d SMs, workspace has %d scratch slots", sms, n_slots); return AON_E_UNSUPPORTED; }
  TcParams p;
  fill_common(p, kind, precision, opts);
  p.packed[0] = (const char*)packed_c; p.packed[1] = (const char*)packed_f;
  p.folded[0] = folded_c; p.folded[1] = folded_f;
  p.S_level[0] = S0; p.S_level[1] = S1;
  p.n_levels = 2;
  p.rays_o = rays_o; p.rays_d = rays_d; p.viewdirs = viewdirs;
  if (cam) { p.cam_on = 1; p.cam = *cam; p.img_H = H; p.img_W = W; p.focal = focal; p.ray0 = ray0; }
  p.t_vals = t0; p.t_stride = t0_stride;
  p.u = u; p.u_stride = u_stride;
  p.slots = slots; p.slot_mask = slot_mask; p.n_slots = n_slots;
  p.R = R; p.white_bkgd = white_bkgd;
  p.seg_len = S0;
  p.out5 = out5; p.coarse_out5 = coarse_out5;
  return launch_any(p, kind, precision, ((R + 255) / 256) * 2, 1, st);
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
  const int x3 = precision == AON_PREC_TC_F16X3;
  PackSrc src;
  memset(&src, 0, sizeof(src));
  for (int i = 0; i < num_layers(kind); ++i) { src.w[i] = w[i]; src.b[i] = b[i]; }
  long byte0 = 0;
  for (int i = 0; i < num_gemm(kind); ++i) {
    src.g[i] = gemm_layers(kind)[i];
    src.in_features[i] = layer_shapes(kind)[src.g[i].src][1];
    src.unit_byte0[i] = byte0;
    byte0 += unit_bytes(P.u[i], x3);
  }
  src.unit_byte0[num_gemm(kind)] = byte0;
  uint16_t* out = (uint16_t*)packed;
  const int blocks = 592;
  if (precision == AON_PREC_TC_F16X3) pack_stream_kernel<true, false><<<blocks, 256, 0, st>>>(P, src, out);
  else if (precision == AON_PREC_TC_F16) pack_stream_kernel<false, false><<<blocks, 256, 0, st>>>(P, src, out);
  else pack_stream_kernel<false, true><<<blocks, 256, 0, st>>>(P, src, out);
  AON_LAUNCH_CHECK();
  return pack_tail(kind, L, w, b, (char*)packed, st);
}

extern "C" int aon_train_tiles(int R, int S) { return (R <= 0 || S <= 0) ? 0 : ((R + 255) / 256) * 2 * S; }

extern "C" int aon_forward_train(int kind, int precision, const void* packed, const float* folded, const float* rays_o,
                                 const float* rays_d, const float* viewdirs, const float* t_vals, long t_stride, int R, int S,
                                 const AonTrainDump* d, aon_stream_t stream) {
  AON_REQUIRE(kind == AON_KIND_VANILLA || kind == AON_KIND_AUTODECODER, "bad kind %d", kind);
  AON_REQUIRE(packed && rays_o && rays_d && viewdirs && t_vals && d, "aon_forward_train: null pointer");
  AON_REQUIRE(kind == AON_KIND_VANILLA || (folded != nullptr && d->warped != nullptr),
              "aon_forward_train: auto-decoder needs the folded biases of aon_fold_latents() and a warped-position buffer");
  AON_REQUIRE(R >= 1 && S >= 1 && (t_stride == 0 || t_stride >= S), "aon_forward_train: bad sizes R=%d S=%d t_stride=%ld", R, S, t_stride);
  AON_REQUIRE(((uintptr_t)packed & 255) == 0, "packed buffer must be 256-byte aligned");
  const bool x3 = precision == AON_PREC_TC_F16X3;
  AON_REQUIRE(d->raw && d->enc_hi && (!x3 || d->enc_lo), "aon_forward_train: raw / encoding buffers missing");
  TrainDump td;
  memset(&td, 0, sizeof(td));
  const int ng = num_gemm(kind);
  for (int i = 0; i < ng; ++i) {
    AON_REQUIRE(d->act_hi[i] && (!x3 || d->act_lo[i]), "aon_forward_train: activation plane of unit %d missing", i);
    td.hi[i] = (uint4*)d->act_hi[i]; td.lo[i] = (uint4*)d->act_lo[i]; td.bits[i] = (uint32_t*)d->relu_bits[i];
  }
  td.e_hi = (uint4*)d->enc_hi; td.e_lo = (uint4*)d->enc_lo; td.raw = (float4*)d->raw; td.warped = d->warped;
  if (const char* e = getenv("AON_TRAIN_DEBUG")) td.dbg_flags = atoi(e);   // profiling experiments (tools/prof_fwd_train.py); results are wrong
  return forward_train_tc(kind, precision, packed, folded, rays_o, rays_d, viewdirs, t_vals, t_stride, R, S, td, (cudaStream_t)stream);
}

// Introspection for tools / tests (not part of the public ABI in include/aon.h): program size and shared-memory plan.
extern "C" int aon_debug_program_info(int kind, int precision, int* n_units, int* n_stages, int* smem_bytes) {
  const Program P = build_program(kind, precision);
  if (n_units) *n_units = P.n_units;
  if (n_stages) *n_stages = (int)(P.stream_bytes / 1024);  // KB of weights streamed per sample per CTA
  if (smem_bytes) {
    const bool x3 = precision == AON_PREC_TC_F16X3;
    *smem_bytes = kind == AON_KIND_VANILLA ? (x3 ? SmemPlan<0, true>::TOTAL : SmemPlan<0, false>::TOTAL)
                                           : (x3 ? SmemPlan<1, true>::TOTAL : SmemPlan<1, false>::TOTAL);
  }
  return 0;
}
