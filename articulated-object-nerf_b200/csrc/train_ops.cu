// train_ops.cu -- per-ray / per-element stages of the TRAINING path (stage 1) with hand-written
// adjoints: positional encoding, activations + alpha compositing, and the Adam update.  The MLP contractions of the
// training step still run as library GEMMs (torch.nn.functional.linear); everything around them is here so that the
// training step launches no elementwise torch kernels on [rays x samples x features] tensors.
//
// Reference functions replaced (forward) and differentiated (backward):
//   pos_enc                      helper.py:136-140
//   activations                  model.py:186-187 / model_autodecoder.py:321-323
//   volumetric_rendering         helper.py:157-195
//   Adam step + LR               model.py:386-419 (torch.optim.Adam, betas (0.9, 0.999), eps 1e-8)
#include "aon_common.cuh"

namespace aon {

// ---- pos_enc --------------------------------------------------------------------------------------------------------
// out[n, 3 + 6L] = [x, sin(2^k x) (k-major, xyz inner), sin(2^k x + pi/2)]; one thread per (point, frequency) pair.
__global__ void pos_enc_fwd_kernel(const float* __restrict__ x, long n, int L, float* __restrict__ out) {
  const int C = 3 + 6 * L;
  const long total = n * (L + 1);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long i = idx / (L + 1);
    const int k = (int)(idx % (L + 1)) - 1;
    const float v[3] = {x[3 * i + 0], x[3 * i + 1], x[3 * i + 2]};
    float* o = out + i * C;
    if (k < 0) {
      o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    } else {
      const float sc = (float)(1 << k);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float xb = v[d] * sc;
        o[3 + 3 * k + d] = sinf(xb);
        o[3 + 3 * L + 3 * k + d] = sinf(__fadd_rn(xb, AON_HALF_PI_F));
      }
    }
  }
}

// gx[n,3] = g_id + sum_k 2^k (g_sin * cos(2^k x) + g_cos * cos(2^k x + pi/2));  one thread per point.
// Both cosines are evaluated at the forward's own (rounded) arguments, which is the derivative autograd takes of
// helper.py:139.
__global__ void pos_enc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, long n, int L, float* __restrict__ gx) {
  const int C = 3 + 6 * L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float* gi = g + i * C;
    float acc[3] = {gi[0], gi[1], gi[2]};
    for (int k = 0; k < L; ++k) {
      const float sc = (float)(1 << k);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float xb = x[3 * i + d] * sc;
        const float c0 = cosf(xb), c1 = cosf(__fadd_rn(xb, AON_HALF_PI_F));
        acc[d] = fmaf(sc, fmaf(gi[3 + 3 * k + d], c0, gi[3 + 3 * L + 3 * k + d] * c1), acc[d]);
      }
    }
    gx[3 * i + 0] = acc[0]; gx[3 * i + 1] = acc[1]; gx[3 * i + 2] = acc[2];
  }
}

// ---- activations + alpha compositing ------------------------------------------------------------------------------------
// act_mode 0: rgb = sigmoid(raw), sigma = relu(raw)             (model.py:186-187)
// act_mode 1: rgb = sigmoid(raw) * 1.002 - 0.001, sigma = softplus(raw - 1)   (model_autodecoder.py:321-323)
__device__ __forceinline__ float act_rgb(float raw, int mode) {
  const float s = sigmoidf_ref(raw);
  return mode ? __fsub_rn(__fmul_rn(s, 1.002f), 0.001f) : s;
}
__device__ __forceinline__ float act_sigma(float raw, int mode) {
  return mode ? softplusf_ref(__fadd_rn(raw, -1.0f)) : fmaxf(raw, 0.f);
}

// One WARP per ray: lane l handles samples l, l + 32, ...; the exclusive transmittance product (helper.py:171-176) is a
// multiplicative warp scan per 32-sample chunk with a running carry (the reference's sequential cumprod re-associated: equal
// to fp32 rounding), the ray sums are warp reductions -- the "warp shuffles in registers" form of the compositing.  All global
// accesses are coalesced along the sample axis.
constexpr int COMP_WARPS = 4;

__global__ void __launch_bounds__(COMP_WARPS * 32)
composite_fwd_kernel(const float* __restrict__ raw_rgb, const float* __restrict__ raw_sigma, const float* __restrict__ t_vals,
                     long t_stride, const float* __restrict__ dirs, int R, int S, int white_bkgd, int act_mode,
                     float* __restrict__ comp_rgb, float* __restrict__ acc, float* __restrict__ depth,
                     float* __restrict__ weights, float* __restrict__ trans_out) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
  if (ray >= R) return;                                   // whole warps leave together
  const float dx = dirs[3 * (size_t)ray], dy = dirs[3 * (size_t)ray + 1], dz = dirs[3 * (size_t)ray + 2];
  const float dnorm = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));   // helper.py:168
  const float* tv = t_vals + (size_t)ray * t_stride;
  const float* rr = raw_rgb + (size_t)ray * S * 3;
  const float* rs = raw_sigma + (size_t)ray * S;
  float carry = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, cdepth = 0.f, cacc = 0.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool valid = s < S;
    float r = 0.f, g = 0.f, b = 0.f, alpha = 0.f, t_cur = 0.f;
    if (valid) {
      t_cur = tv[s];
      r = act_rgb(rr[3 * s + 0], act_mode); g = act_rgb(rr[3 * s + 1], act_mode); b = act_rgb(rr[3 * s + 2], act_mode);
      const float sigma = act_sigma(rs[s], act_mode);
      const float delta = (s + 1 < S) ? __fsub_rn(tv[s + 1], t_cur) : 1e10f;      // helper.py:160-166
      alpha = __fsub_rn(1.0f, expf(__fmul_rn(-sigma, __fmul_rn(delta, dnorm))));   // helper.py:170
    }
    const float f = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;      // helper.py:171-175
    float p = f;                                            // inclusive product scan over the chunk
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float q = __shfl_up_sync(0xffffffffu, p, o);
      if (lane >= o) p *= q;
    }
    float excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.0f;
    const float trans = carry * excl;
    carry *= __shfl_sync(0xffffffffu, p, 31);
    const float w = __fmul_rn(alpha, trans);                // helper.py:176
    cr = fmaf(w, r, cr); cg = fmaf(w, g, cg); cb = fmaf(w, b, cb);
    cdepth = fmaf(w, t_cur, cdepth);
    cacc += w;
    if (valid) {
      if (weights) weights[(size_t)ray * S + s] = w;
      if (trans_out) trans_out[(size_t)ray * S + s] = trans;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cr += __shfl_xor_sync(0xffffffffu, cr, o); cg += __shfl_xor_sync(0xffffffffu, cg, o); cb += __shfl_xor_sync(0xffffffffu, cb, o);
    cdepth += __shfl_xor_sync(0xffffffffu, cdepth, o); cacc += __shfl_xor_sync(0xffffffffu, cacc, o);
  }
  if (lane == 0) {
    if (isnan(cdepth)) cdepth = INFINITY;                                            // helper.py:179
    else if (isinf(cdepth)) cdepth = cdepth > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    if (white_bkgd) {                                                                // helper.py:185-186
      const float bg = __fsub_rn(1.0f, cacc);
      cr += bg; cg += bg; cb += bg;
    }
    comp_rgb[3 * (size_t)ray + 0] = cr; comp_rgb[3 * (size_t)ray + 1] = cg; comp_rgb[3 * (size_t)ray + 2] = cb;
    acc[ray] = cacc;
    depth[ray] = cdepth;
  }
}

// Adjoint of composite_fwd_kernel.  With w_s = alpha_s T_s, T_s = prod_{j<s} (1 - alpha_j + 1e-10):
//   dL/dw_s     = gC . c_s + gD t_s + gA - [white] (gC_r + gC_g + gC_b)
//   dL/dc_s     = w_s gC
//   dL/dalpha_s = dL/dw_s T_s - B_s / (1 - alpha_s + 1e-10),   B_s = sum_{j>s} dL/dw_j w_j
//   dL/dsigma_s = dL/dalpha_s dist_s exp(-sigma_s dist_s)
// One warp per ray, chunks of 32 samples from the last to the first; B is an exclusive SUFFIX sum (shuffle-down scan per chunk
// + carry); T_s and w_s come from the forward.
__global__ void __launch_bounds__(COMP_WARPS * 32)
composite_bwd_kernel(const float* __restrict__ raw_rgb, const float* __restrict__ raw_sigma, const float* __restrict__ t_vals,
                     long t_stride, const float* __restrict__ dirs, const float* __restrict__ weights,
                     const float* __restrict__ trans_in, const float* __restrict__ g_rgb, const float* __restrict__ g_acc,
                     const float* __restrict__ g_depth, int R, int S, int white_bkgd, int act_mode,
                     float* __restrict__ g_raw_rgb, float* __restrict__ g_raw_sigma) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float dx = dirs[3 * (size_t)ray], dy = dirs[3 * (size_t)ray + 1], dz = dirs[3 * (size_t)ray + 2];
  const float dnorm = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
  const float* tv = t_vals + (size_t)ray * t_stride;
  const float* rr = raw_rgb + (size_t)ray * S * 3;
  const float* rs = raw_sigma + (size_t)ray * S;
  const float* ww = weights + (size_t)ray * S;
  const float* tt = trans_in + (size_t)ray * S;
  const float gr = g_rgb ? g_rgb[3 * (size_t)ray + 0] : 0.f, gg = g_rgb ? g_rgb[3 * (size_t)ray + 1] : 0.f,
              gb = g_rgb ? g_rgb[3 * (size_t)ray + 2] : 0.f;
  const float ga = g_acc ? g_acc[ray] : 0.f, gd = g_depth ? g_depth[ray] : 0.f;
  const float gbg = white_bkgd ? (gr + gg + gb) : 0.f;
  const float k = act_mode ? 1.002f : 1.0f;
  float carry = 0.f;                                        // sum of dL/dw_j w_j over the chunks behind this one
  for (int s0 = ((S - 1) / 32) * 32; s0 >= 0; s0 -= 32) {
    const int s = s0 + lane;
    const bool valid = s < S;
    float x = 0.f, gw = 0.f, w = 0.f, T = 0.f, alpha = 0.f, e = 0.f, dist = 0.f, raw_s = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
    if (valid) {
      const float t_cur = tv[s];
      sr = sigmoidf_ref(rr[3 * s + 0]); sg = sigmoidf_ref(rr[3 * s + 1]); sb = sigmoidf_ref(rr[3 * s + 2]);
      const float r = act_mode ? __fsub_rn(__fmul_rn(sr, 1.002f), 0.001f) : sr;
      const float g = act_mode ? __fsub_rn(__fmul_rn(sg, 1.002f), 0.001f) : sg;
      const float b = act_mode ? __fsub_rn(__fmul_rn(sb, 1.002f), 0.001f) : sb;
      raw_s = rs[s];
      const float sigma = act_sigma(raw_s, act_mode);
      const float delta = (s + 1 < S) ? __fsub_rn(tv[s + 1], t_cur) : 1e10f;
      dist = __fmul_rn(delta, dnorm);
      e = expf(__fmul_rn(-sigma, dist));
      alpha = __fsub_rn(1.0f, e);
      w = ww[s]; T = tt[s];
      gw = fmaf(gr, r, fmaf(gg, g, fmaf(gb, b, fmaf(gd, t_cur, ga - gbg))));
      x = gw * w;
    }
    float p = x;                                            // inclusive suffix sum over the chunk
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float q = __shfl_down_sync(0xffffffffu, p, o);
      if (lane + o < 32) p += q;
    }
    float B = __shfl_down_sync(0xffffffffu, p, 1);
    if (lane == 31) B = 0.f;
    B += carry;
    carry += __shfl_sync(0xffffffffu, p, 0);
    if (valid) {
      // colour: dL/draw = w gC act'(raw)
      g_raw_rgb[((size_t)ray * S + s) * 3 + 0] = w * gr * k * sr * (1.0f - sr);
      g_raw_rgb[((size_t)ray * S + s) * 3 + 1] = w * gg * k * sg * (1.0f - sg);
      g_raw_rgb[((size_t)ray * S + s) * 3 + 2] = w * gb * k * sb * (1.0f - sb);
      // density
      const float galpha = gw * T - B / __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      float gsigma = galpha * dist * e;
      if (e == 0.f) gsigma = 0.f;                                // dist = 1e10 (last sample): 0 * huge stays 0, never nan
      const float dact = act_mode ? sigmoidf_ref(__fadd_rn(raw_s, -1.0f)) : (raw_s > 0.f ? 1.0f : 0.f);
      g_raw_sigma[(size_t)ray * S + s] = dact == 0.f ? 0.f : gsigma * dact;
    }
  }
}

// ---- Adam (torch.optim.Adam single-tensor formulas, flat buffer) --------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long n, float omb1, float beta2, float omb2, float eps, float step_size, float sqrt_bc2, float grad_scale) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = fmaf(gi - m[i], omb1, m[i]);                         // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(gi * gi, omb2, v[i] * beta2);                   // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrt_bc2 + eps;                       // (sqrt(v) / sqrt(bc2)).add_(eps)
    p[i] = p[i] - step_size * (mi / denom);                                // addcdiv_(exp_avg, denom, value=-lr/bc1)
  }
}

// same update with the seven step-dependent scalars read from DEVICE memory: a captured CUDA graph of the training step bakes
// kernel arguments in, so the learning-rate schedule and the bias corrections have to arrive through a buffer that the host
// refreshes (one 28-byte copy) before every replay
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long n, const float* __restrict__ sc) {
  const float omb1 = sc[0], beta2 = sc[1], omb2 = sc[2], eps = sc[3], step_size = sc[4], sqrt_bc2 = sc[5], grad_scale = sc[6];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = fmaf(gi - m[i], omb1, m[i]);
    const float vi = fmaf(gi * gi, omb2, v[i] * beta2);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrt_bc2 + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

static inline int grid_for(long n, int block) {
  long g = (n + block - 1) / block;
  const long cap = 148L * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace aon

using namespace aon;

extern "C" int aon_pos_enc(const float* x, long n, int max_deg, float* out, aon_stream_t stream) {
  AON_REQUIRE(x && out, "aon_pos_enc: null pointer");
  AON_REQUIRE(n >= 0 && max_deg >= 1 && max_deg <= 16, "aon_pos_enc: bad sizes n=%ld max_deg=%d", n, max_deg);
  if (n == 0) return AON_OK;
  pos_enc_fwd_kernel<<<grid_for(n * (max_deg + 1), 256), 256, 0, (cudaStream_t)stream>>>(x, n, max_deg, out);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_pos_enc_backward(const float* x, const float* g_out, long n, int max_deg, float* g_x, aon_stream_t stream) {
  AON_REQUIRE(x && g_out && g_x, "aon_pos_enc_backward: null pointer");
  AON_REQUIRE(n >= 0 && max_deg >= 1 && max_deg <= 16, "aon_pos_enc_backward: bad sizes n=%ld max_deg=%d", n, max_deg);
  if (n == 0) return AON_OK;
  pos_enc_bwd_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(x, g_out, n, max_deg, g_x);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_composite(const float* raw_rgb, const float* raw_sigma, const float* t_vals, long t_stride,
                             const float* dirs, int R, int S, int white_bkgd, int act_mode, float* comp_rgb, float* acc,
                             float* depth, float* weights, float* trans, aon_stream_t stream) {
  AON_REQUIRE(raw_rgb && raw_sigma && t_vals && dirs && comp_rgb && acc && depth, "aon_composite: null pointer");
  AON_REQUIRE(R >= 0 && S >= 1, "aon_composite: bad sizes R=%d S=%d", R, S);
  AON_REQUIRE(t_stride == 0 || t_stride >= S, "aon_composite: bad t_stride %ld", t_stride);
  AON_REQUIRE(act_mode == 0 || act_mode == 1, "aon_composite: bad act_mode %d", act_mode);
  if (R == 0) return AON_OK;
  composite_fwd_kernel<<<(R + COMP_WARPS - 1) / COMP_WARPS, COMP_WARPS * 32, 0, (cudaStream_t)stream>>>(raw_rgb, raw_sigma, t_vals, t_stride, dirs, R, S, white_bkgd,
                                                                       act_mode, comp_rgb, acc, depth, weights, trans);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_composite_backward(const float* raw_rgb, const float* raw_sigma, const float* t_vals, long t_stride,
                                      const float* dirs, const float* weights, const float* trans, const float* g_comp_rgb,
                                      const float* g_acc, const float* g_depth, int R, int S, int white_bkgd, int act_mode,
                                      float* g_raw_rgb, float* g_raw_sigma, aon_stream_t stream) {
  AON_REQUIRE(raw_rgb && raw_sigma && t_vals && dirs && weights && trans && g_raw_rgb && g_raw_sigma,
              "aon_composite_backward: null pointer");
  AON_REQUIRE(R >= 0 && S >= 1, "aon_composite_backward: bad sizes R=%d S=%d", R, S);
  AON_REQUIRE(t_stride == 0 || t_stride >= S, "aon_composite_backward: bad t_stride %ld", t_stride);
  AON_REQUIRE(act_mode == 0 || act_mode == 1, "aon_composite_backward: bad act_mode %d", act_mode);
  if (R == 0) return AON_OK;
  composite_bwd_kernel<<<(R + COMP_WARPS - 1) / COMP_WARPS, COMP_WARPS * 32, 0, (cudaStream_t)stream>>>(raw_rgb, raw_sigma, t_vals, t_stride, dirs, weights, trans,
                                                                       g_comp_rgb, g_acc, g_depth, R, S, white_bkgd, act_mode,
                                                                       g_raw_rgb, g_raw_sigma);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, double lr, double beta1,
                             double beta2, double eps, long step, double grad_scale, aon_stream_t stream) {
  AON_REQUIRE(params && grads && exp_avg && exp_avg_sq, "aon_adam_step: null pointer");
  AON_REQUIRE(n >= 0 && step >= 1, "aon_adam_step: bad n=%ld / step=%ld (step counts from 1)", n, step);
  if (n == 0) return AON_OK;
  // scalars are doubles like torch's Python-side hyper-parameters: 1 - beta2 = 0.001 must not inherit the rounding of 0.999f
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2,
                                                                  (float)(1.0 - beta2), (float)eps, (float)(lr / bc1),
                                                                  (float)sqrt(bc2), (float)grad_scale);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

extern "C" int aon_adam_scalars(double lr, double beta1, double beta2, double eps, long step, double grad_scale, float* out7_host) {
  AON_REQUIRE(out7_host && step >= 1, "aon_adam_scalars: bad argument");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  out7_host[0] = (float)(1.0 - beta1); out7_host[1] = (float)beta2; out7_host[2] = (float)(1.0 - beta2); out7_host[3] = (float)eps;
  out7_host[4] = (float)(lr / bc1); out7_host[5] = (float)sqrt(bc2); out7_host[6] = (float)grad_scale;
  return AON_OK;
}

extern "C" int aon_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, const float* scalars7_dev,
                                 aon_stream_t stream) {
  AON_REQUIRE(params && grads && exp_avg && exp_avg_sq && scalars7_dev && n >= 0, "aon_adam_step_dev: bad argument");
  if (n == 0) return AON_OK;
  adam_dev_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, scalars7_dev);
  AON_LAUNCH_CHECK();
  return AON_OK;
}
