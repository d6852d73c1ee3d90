// render_simt.cu -- AON_PREC_FP32: one fused kernel per level (encode -> MLP -> activations ->
// alpha compositing) with the MLP contraction in fp32 FFMA on the CUDA cores.  This is the parity
// anchor (no operand rounding at all); the tcgen05 kernel in render_tc.cu is the throughput path.
//
// Work item = 128 rays; the CTA walks the S samples of those rays one "sample plane" (128 rows) at a
// time, so each ray's transmittance / colour / depth accumulators are running scalars in the
// registers of one thread and no [rays x samples x features] tensor ever reaches HBM.
//
// Reference functions replaced: cast_rays + pos_enc (helper.py:25-26,136-140), NeRFMLP.forward
// (model.py:95-120 / model_autodecoder.py:171-239), activations (model.py:186-187 /
// model_autodecoder.py:321-323), volumetric_rendering (helper.py:157-195).
#include "aon_common.cuh"

namespace aon {

constexpr int RT = 128;        // rows (rays) per CTA
constexpr int NT = 256;        // threads per CTA
constexpr int KC = 16;         // K rows of weights per pipeline stage
constexpr int XLD = RT;        // leading dimension of the [k][row] activation buffers

struct SimtParams {
  const char* packed;
  PackedLayout L;
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
  float* out5;     // interleaved [R,5] = (r, g, b, acc, depth); when set, the three planes above are not written
  float* weights;
};

struct Smem {
  float X[256][XLD];      // hidden activations, [feature][row]
  float E[KE][XLD];       // positional encoding of the (warped) sample position
  float V[KV][XLD];       // view-direction encoding (per ray, constant over samples)
  float P[KP][XLD];       // raw sample position (auto-decoder deformation input), rows 3.. are zero
  float W[2][KC][256];    // weight pipeline stages
  float red[2][4][RT];    // head partial sums: [half][channel][row]
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// One GEMM layer: X[0:N] <- act(Wt^T [X[0:K1]; aux] + bias), in place.
// Thread (tr, tc): rows {tr*4+i, 64+tr*4+i}, cols {tc*4 + 64*q + j}.
template <int N, int K1, int KAUX, bool RELU>
__device__ __forceinline__ void gemm_layer(Smem& sm, const float (*aux)[XLD], const float* __restrict__ Wt,
                                           const float* __restrict__ bias) {
  constexpr int NQ = N / 64;            // column groups of 4 per thread: 4 (N=256) or 2 (N=128)
  constexpr int K = K1 + KAUX;
  static_assert(K % KC == 0, "K must be a multiple of KC");
  constexpr int NCHUNK = K / KC;
  const int tid = threadIdx.x;
  const int tr = tid & 15, tc = tid >> 4;
  float acc[8][NQ * 4];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + tc * 4 + 64 * q));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i][q * 4 + 0] = b4.x; acc[i][q * 4 + 1] = b4.y; acc[i][q * 4 + 2] = b4.z; acc[i][q * 4 + 3] = b4.w;
    }
  }
  auto prefetch = [&](int c, int stage) {
    // KC rows x N floats = KC*N/4 16-byte packets
    constexpr int PKTS = KC * N / 4;
    const float* src = Wt + (size_t)c * KC * N;
#pragma unroll
    for (int p = tid; p < PKTS; p += NT) {
      const int k = p / (N / 4), n4 = p % (N / 4);
      cp_async16(&sm.W[stage][k][n4 * 4], src + (size_t)k * N + n4 * 4);
    }
    cp_async_commit();
  };
  prefetch(0, 0);
  for (int c = 0; c < NCHUNK; ++c) {
    if (c + 1 < NCHUNK) {
      prefetch(c + 1, (c + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int st = c & 1;
    const int k0 = c * KC;
    const float(*src)[XLD] = (k0 < K1) ? (const float(*)[XLD])(&sm.X[k0]) : (aux + (k0 - K1));
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&src[kk][tr * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&src[kk][64 + tr * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(&sm.W[st][kk][tc * 4 + 64 * q]);
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][q * 4 + j] = fmaf(a[i], b[j], acc[i][q * 4 + j]);
      }
    }
    __syncthreads();  // all reads of stage st (and, on the last chunk, of X) are done
  }
  // epilogue: write act(acc) over X[0:N]
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = tc * 4 + 64 * q + j;
      float4 o0, o1;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = RELU ? fmaxf(acc[i][q * 4 + j], 0.f) : acc[i][q * 4 + j];
      o0 = make_float4(v[0], v[1], v[2], v[3]);
      o1 = make_float4(v[4], v[5], v[6], v[7]);
      *reinterpret_cast<float4*>(&sm.X[col][tr * 4]) = o0;
      *reinterpret_cast<float4*>(&sm.X[col][64 + tr * 4]) = o1;
    }
  __syncthreads();
}

// Small head: out[c][row] = b[c] + sum_k w[c][k] * X[k][row], NOUT <= 3; two threads per row split K.
template <int NOUT, int K>
__device__ __forceinline__ void head_partial(Smem& sm, const float* __restrict__ w) {
  const int row = threadIdx.x & (RT - 1), half = threadIdx.x >> 7;
  float s[NOUT];
#pragma unroll
  for (int c = 0; c < NOUT; ++c) s[c] = 0.f;
  const int k0 = half * (K / 2);
#pragma unroll 8
  for (int k = k0; k < k0 + K / 2; ++k) {
    const float x = sm.X[k][row];
#pragma unroll
    for (int c = 0; c < NOUT; ++c) s[c] = fmaf(__ldg(w + c * K + k), x, s[c]);
  }
#pragma unroll
  for (int c = 0; c < NOUT; ++c) sm.red[half][c][row] = s[c];
}

// pos_enc (helper.py:136-140) of 3 values into rows of a [k][row] buffer; this thread does
// frequencies [f0, f1).
template <int L>
__device__ __forceinline__ void encode_rows(float (*dst)[XLD], int row, float x, float y, float z, int f0,
                                            int f1, bool write_identity) {
  if (write_identity) {
    dst[0][row] = x; dst[1][row] = y; dst[2][row] = z;
  }
  const float v[3] = {x, y, z};
  for (int f = f0; f < f1; ++f) {
    const float sc = (float)(1 << f);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float xb = v[d] * sc;  // exact (power of two)
      dst[3 + f * 3 + d][row] = sinf(xb);
      dst[3 + 3 * L + f * 3 + d][row] = sinf(__fadd_rn(xb, AON_HALF_PI_F));
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(NT, 1) render_simt_kernel(const SimtParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int row = tid & (RT - 1), half = tid >> 7;
  const long ray0 = (long)blockIdx.x * RT;
  const long ray = ray0 + row;
  const bool valid = ray < p.R;
  const long rl = valid ? ray : (long)p.R - 1;  // clamp loads of the ragged tail
  const int S = p.S;
  const PackedLayout& L = p.L;
  auto Wp = [&](int i) { return reinterpret_cast<const float*>(p.packed + L.w[i]); };
  auto Bp = [&](int i) {
    return (KIND == AON_KIND_AUTODECODER && L.fold[i] >= 0) ? p.folded + L.fold[i]
                                                             : reinterpret_cast<const float*>(p.packed + L.bias[i]);
  };
  auto HW = [&](int i) { return reinterpret_cast<const float*>(p.packed + L.head_w[i]); };
  auto HB = [&](int i) { return reinterpret_cast<const float*>(p.packed + L.head_b[i]); };

  const float ox = p.rays_o[3 * rl + 0], oy = p.rays_o[3 * rl + 1], oz = p.rays_o[3 * rl + 2];
  const float dx = p.rays_d[3 * rl + 0], dy = p.rays_d[3 * rl + 1], dz = p.rays_d[3 * rl + 2];

  // zero the padding rows once; view encoding once per work item
  for (int i = tid; i < RT; i += NT) sm.E[KE - 1][i] = 0.f;
  for (int i = tid; i < (KV - 27) * RT; i += NT) sm.V[27 + i / RT][i % RT] = 0.f;
  for (int i = tid; i < (KP - 3) * RT; i += NT) sm.P[3 + i / RT][i % RT] = 0.f;
  {
    const float vx = p.viewdirs[3 * rl + 0], vy = p.viewdirs[3 * rl + 1], vz = p.viewdirs[3 * rl + 2];
    encode_rows<4>(sm.V, row, vx, vy, vz, half * 2, half * 2 + 2, half == 0);
  }
  // |d| (helper.py:168)
  const float dnorm = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
  const float* tv = p.t_vals + (p.t_stride ? rl * p.t_stride : 0);

  // per-ray running state (threads 0..127)
  float trans = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, cdepth = 0.f, cacc = 0.f;
  float t_cur = tv[0];
  __syncthreads();

  for (int s = 0; s < S; ++s) {
    const float t_next = (s + 1 < S) ? tv[s + 1] : 0.f;
    // cast_rays (helper.py:25-26): o + t * d with separately rounded mul and add
    const float px = __fadd_rn(ox, __fmul_rn(t_cur, dx));
    const float py = __fadd_rn(oy, __fmul_rn(t_cur, dy));
    const float pz = __fadd_rn(oz, __fmul_rn(t_cur, dz));
    float ex = px, ey = py, ez = pz;

    if (KIND == AON_KIND_AUTODECODER) {
      if (half == 0) { sm.P[0][row] = px; sm.P[1][row] = py; sm.P[2][row] = pz; }
      __syncthreads();
      gemm_layer<128, 0, KP, true>(sm, sm.P, Wp(0), Bp(0));
      gemm_layer<128, 128, 0, true>(sm, nullptr, Wp(1), Bp(1));
      gemm_layer<128, 128, 0, true>(sm, nullptr, Wp(2), Bp(2));
      gemm_layer<128, 128, 0, true>(sm, nullptr, Wp(3), Bp(3));
      head_partial<3, 128>(sm, HW(0));
      __syncthreads();
      // model_autodecoder.py:203: x' = deformation_layer(h) + pos
      ex = __fadd_rn((sm.red[0][0][row] + sm.red[1][0][row]) + __ldg(HB(0) + 0), px);
      ey = __fadd_rn((sm.red[0][1][row] + sm.red[1][1][row]) + __ldg(HB(0) + 1), py);
      ez = __fadd_rn((sm.red[0][2][row] + sm.red[1][2][row]) + __ldg(HB(0) + 2), pz);
    }
    encode_rows<10>(sm.E, row, ex, ey, ez, half * 5, half * 5 + 5, half == 0);
    __syncthreads();

    constexpr int G0 = KIND == AON_KIND_AUTODECODER ? 4 : 0;  // index of pts_linears.0
    gemm_layer<256, 0, KE, true>(sm, sm.E, Wp(G0 + 0), Bp(G0 + 0));
    gemm_layer<256, 256, 0, true>(sm, nullptr, Wp(G0 + 1), Bp(G0 + 1));
    gemm_layer<256, 256, 0, true>(sm, nullptr, Wp(G0 + 2), Bp(G0 + 2));
    gemm_layer<256, 256, 0, true>(sm, nullptr, Wp(G0 + 3), Bp(G0 + 3));
    gemm_layer<256, 256, 0, true>(sm, nullptr, Wp(G0 + 4), Bp(G0 + 4));
    gemm_layer<256, 256, KE, true>(sm, sm.E, Wp(G0 + 5), Bp(G0 + 5));
    gemm_layer<256, 256, 0, true>(sm, nullptr, Wp(G0 + 6), Bp(G0 + 6));
    gemm_layer<256, 256, 0, true>(sm, nullptr, Wp(G0 + 7), Bp(G0 + 7));
    constexpr int HD = KIND == AON_KIND_AUTODECODER ? 1 : 0;  // density head index
    head_partial<1, 256>(sm, HW(HD));
    __syncthreads();
    const float raw_sigma = (sm.red[0][0][row] + sm.red[1][0][row]) + __ldg(HB(HD));
    gemm_layer<256, 256, 0, false>(sm, nullptr, Wp(G0 + 8), Bp(G0 + 8));    // bottleneck
    gemm_layer<128, 256, KV, true>(sm, sm.V, Wp(G0 + 9), Bp(G0 + 9));       // views_linear.0
    if (KIND == AON_KIND_AUTODECODER) {
      gemm_layer<128, 128, 0, true>(sm, nullptr, Wp(G0 + 10), Bp(G0 + 10));
      gemm_layer<128, 128, 0, true>(sm, nullptr, Wp(G0 + 11), Bp(G0 + 11));
      gemm_layer<128, 128, 0, true>(sm, nullptr, Wp(G0 + 12), Bp(G0 + 12));
    }
    head_partial<3, 128>(sm, HW(HD + 1));
    __syncthreads();

    if (half == 0) {
      const float* hb = HB(HD + 1);
      float r = (sm.red[0][0][row] + sm.red[1][0][row]) + __ldg(hb + 0);
      float g = (sm.red[0][1][row] + sm.red[1][1][row]) + __ldg(hb + 1);
      float b = (sm.red[0][2][row] + sm.red[1][2][row]) + __ldg(hb + 2);
      float sigma;
      if (KIND == AON_KIND_VANILLA) {
        r = sigmoidf_ref(r); g = sigmoidf_ref(g); b = sigmoidf_ref(b);   // model.py:186
        sigma = fmaxf(raw_sigma, 0.f);                                     // model.py:187
      } else {
        // model_autodecoder.py:321-323
        r = __fsub_rn(__fmul_rn(sigmoidf_ref(r), 1.002f), 0.001f);
        g = __fsub_rn(__fmul_rn(sigmoidf_ref(g), 1.002f), 0.001f);
        b = __fsub_rn(__fmul_rn(sigmoidf_ref(b), 1.002f), 0.001f);
        sigma = softplusf_ref(__fadd_rn(raw_sigma, -1.0f));
      }
      // helper.py:160-176
      const float delta = (s + 1 < S) ? __fsub_rn(t_next, t_cur) : 1e10f;
      const float dist = __fmul_rn(delta, dnorm);
      const float alpha = __fsub_rn(1.0f, expf(__fmul_rn(-sigma, dist)));
      const float w = __fmul_rn(alpha, trans);
      cr = fmaf(w, r, cr); cg = fmaf(w, g, cg); cb = fmaf(w, b, cb);
      cdepth = fmaf(w, t_cur, cdepth);
      cacc += w;
      trans = __fmul_rn(trans, __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f));
      if (p.weights && valid) p.weights[ray * S + s] = w;
    }
    t_cur = t_next;
    __syncthreads();  // red[] reuse in the next sample
  }

  if (half == 0 && valid) {
    // helper.py:179-180: nan_to_num(depth, nan=inf) (+-inf -> +-FLT_MAX); the chunk-global clamp
    // that follows is the identity and is not reproduced (it would couple rays across a chunk).
    if (isnan(cdepth)) cdepth = INFINITY;
    else if (isinf(cdepth)) cdepth = cdepth > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    if (p.white_bkgd) {  // helper.py:185-186
      const float bg = __fsub_rn(1.0f, cacc);
      cr += bg; cg += bg; cb += bg;
    }
    if (p.out5 != nullptr) {
      p.out5[5 * ray + 0] = cr; p.out5[5 * ray + 1] = cg; p.out5[5 * ray + 2] = cb; p.out5[5 * ray + 3] = cacc; p.out5[5 * ray + 4] = cdepth;
    } else {
      p.comp_rgb[3 * ray + 0] = cr; p.comp_rgb[3 * ray + 1] = cg; p.comp_rgb[3 * ray + 2] = cb;
      p.acc[ray] = cacc;
      p.depth[ray] = cdepth;
    }
  }
}

int render_level_simt(int kind, const void* packed, const float* folded, const float* rays_o,
                      const float* rays_d, const float* viewdirs, const float* t_vals, long t_stride,
                      int R, int S, int white_bkgd, float* comp_rgb, float* acc, float* depth, float* out5,
                      float* weights, cudaStream_t st) {
  SimtParams p;
  p.packed = (const char*)packed;
  p.L = layout_fp32(kind);
  p.folded = folded;
  p.rays_o = rays_o; p.rays_d = rays_d; p.viewdirs = viewdirs;
  p.t_vals = t_vals; p.t_stride = t_stride;
  p.R = R; p.S = S; p.white_bkgd = white_bkgd;
  p.comp_rgb = comp_rgb; p.acc = acc; p.depth = depth; p.out5 = out5; p.weights = weights;
  const int grid = (R + RT - 1) / RT;
  const size_t smem = sizeof(Smem);
  if (kind == AON_KIND_VANILLA) {
    AON_CUDA_CHECK(cudaFuncSetAttribute(render_simt_kernel<AON_KIND_VANILLA>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    render_simt_kernel<AON_KIND_VANILLA><<<grid, NT, smem, st>>>(p);
  } else {
    AON_CUDA_CHECK(cudaFuncSetAttribute(render_simt_kernel<AON_KIND_AUTODECODER>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    render_simt_kernel<AON_KIND_AUTODECODER><<<grid, NT, smem, st>>>(p);
  }
  AON_LAUNCH_CHECK();
  return AON_OK;
}

// One level in the precision's kernel; out5 (interleaved) or the three planes receive the composite.
int render_level_any(int kind, int precision, const void* packed, const float* folded, const float* rays_o,
                     const float* rays_d, const float* viewdirs, const float* t_vals, long t_stride, int R, int S,
                     int white_bkgd, float* comp_rgb, float* acc, float* depth, float* out5, float* weights,
                     void* workspace, size_t workspace_bytes, const AonRenderOpts* opts, cudaStream_t st) {
  if (precision == AON_PREC_FP32)
    return render_level_simt(kind, packed, folded, rays_o, rays_d, viewdirs, t_vals, t_stride, R, S, white_bkgd, comp_rgb,
                             acc, depth, out5, weights, st);
  if (precision >= AON_PREC_TC_F16X3 && precision <= AON_PREC_TC_BF16)
    return render_level_tc(kind, precision, packed, folded, rays_o, rays_d, viewdirs, t_vals, t_stride, R, S, white_bkgd,
                           comp_rgb, acc, depth, out5, weights, workspace, workspace_bytes, opts, st);
  set_error("bad precision %d", precision);
  return AON_E_ARG;
}

}  // namespace aon

using namespace aon;

extern "C" int aon_render_level(int kind, int precision, const void* packed, const float* folded,
                                const float* rays_o, const float* rays_d, const float* viewdirs,
                                const float* t_vals, long t_stride, int R, int S, int white_bkgd,
                                float* comp_rgb, float* acc, float* depth, float* weights, void* workspace,
                                size_t workspace_bytes, const AonRenderOpts* opts, aon_stream_t stream) {
  AON_REQUIRE(kind == AON_KIND_VANILLA || kind == AON_KIND_AUTODECODER, "bad kind %d", kind);
  AON_REQUIRE(packed && rays_o && rays_d && viewdirs && t_vals && comp_rgb && acc && depth,
              "aon_render_level: null pointer");
  AON_REQUIRE(kind == AON_KIND_VANILLA || folded != nullptr,
              "aon_render_level: auto-decoder needs the folded biases of aon_fold_latents()");
  AON_REQUIRE(R >= 0 && S >= 1, "aon_render_level: bad sizes R=%d S=%d", R, S);
  AON_REQUIRE(t_stride == 0 || t_stride >= S, "aon_render_level: bad t_stride %ld", t_stride);
  AON_REQUIRE(((uintptr_t)packed & 255) == 0, "packed buffer must be 256-byte aligned");
  AON_REQUIRE(workspace == nullptr || ((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  if (R == 0) return AON_OK;
  return render_level_any(kind, precision, packed, folded, rays_o, rays_d, viewdirs, t_vals, t_stride, R, S, white_bkgd,
                          comp_rgb, acc, depth, nullptr, weights, workspace, workspace ? workspace_bytes : 0, opts,
                          (cudaStream_t)stream);
}
