// aon_api.cu -- C-ABI entry points that are not the render kernel itself: introspection, weight
// packing, latent folding, ray generation (A1+A2), coarse sampling (A3), hierarchical sampling (A7)
// and the host-buffer whole-image call (A8/A11).  See include/aon.h for the contract.
#include <stdarg.h>
#include <string.h>

#include "aon_common.cuh"
#include "sampling.cuh"

namespace aon {

thread_local char g_err[512] = "";
thread_local long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_layers(int kind) { return kind == AON_KIND_VANILLA ? V_NUM_LAYERS : A_NUM_LAYERS; }
const int (*layer_shapes(int kind))[2] { return kind == AON_KIND_VANILLA ? V_SHAPES : A_SHAPES; }
int num_gemm(int kind) { return kind == AON_KIND_VANILLA ? V_NUM_GEMM : A_NUM_GEMM; }
const GemmLayer* gemm_layers(int kind) { return kind == AON_KIND_VANILLA ? V_GEMM : A_GEMM; }
int num_heads(int kind) { return kind == AON_KIND_VANILLA ? 2 : 3; }
static const Head V_HEADS[2] = {V_HEAD_DENSITY, V_HEAD_RGB};
static const Head A_HEADS[3] = {A_HEAD_DEFORM, A_HEAD_DENSITY, A_HEAD_RGB};
const Head* heads(int kind) { return kind == AON_KIND_VANILLA ? V_HEADS : A_HEADS; }

static int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Everything after the weight blocks (biases, latent blocks, heads) is fp32 in every precision.
void layout_tail(int kind, PackedLayout& L, int64_t off) {
  const GemmLayer* g = gemm_layers(kind);
  int fold = 0;
  for (int i = 0; i < num_gemm(kind); ++i) {
    L.bias[i] = off;
    off += g[i].N * 4;
    if (g[i].lat_col0 >= 0) {
      L.wlat[i] = off;
      off += (int64_t)g[i].lat_cnt * g[i].N * 4;
      L.fold[i] = fold;
      fold += g[i].N;
    } else {
      L.wlat[i] = -1;
      L.fold[i] = -1;
    }
  }
  const Head* h = heads(kind);
  for (int i = 0; i < num_heads(kind); ++i) {
    L.head_w[i] = off;
    off += (int64_t)h[i].N * h[i].K * 4;
    L.head_b[i] = off;
    off += 16;
  }
  L.total_bytes = align_up(off, 256);
}

PackedLayout layout_fp32(int kind) {
  PackedLayout L;
  memset(&L, 0, sizeof(L));
  const GemmLayer* g = gemm_layers(kind);
  int64_t off = 0;
  for (int i = 0; i < num_gemm(kind); ++i) {
    L.w[i] = off;
    off += (int64_t)(g[i].K1 + kauxOf(g[i].aux)) * g[i].N * 4;
  }
  layout_tail(kind, L, off);
  return L;
}

// ---- pack kernels -------------------------------------------------------------------------------
// Wt[k][n] = W[n][col(k)] with k over [X rows | aux rows (zero padded)].
__global__ void pack_gemm_fp32_kernel(const float* __restrict__ W, int in_features, GemmLayer g,
                                      float* __restrict__ Wt) {
  const int K = g.K1 + kauxOf(g.aux);
  const int total = K * g.N;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx / g.N, n = idx % g.N;
    float v = 0.f;
    if (k < g.K1) {
      v = W[(size_t)n * in_features + k];
    } else if (k - g.K1 < g.aux_cnt) {
      v = W[(size_t)n * in_features + g.aux_col0 + (k - g.K1)];
    }
    Wt[idx] = v;
  }
}

// bias copy + latent block Wlat[j][n] = W[n][lat_col0 + j]
__global__ void pack_tail_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                 int in_features, GemmLayer g, float* __restrict__ bias,
                                 float* __restrict__ wlat) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int n = tid; n < g.N; n += nth) bias[n] = b[n];
  if (wlat != nullptr) {
    for (int idx = tid; idx < g.lat_cnt * g.N; idx += nth) {
      const int j = idx / g.N, n = idx % g.N;
      wlat[idx] = W[(size_t)n * in_features + g.lat_col0 + j];
    }
  }
}

__global__ void pack_head_kernel(const float* __restrict__ W, const float* __restrict__ b, int N,
                                 int K, float* __restrict__ hw, float* __restrict__ hb) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < N * K; i += nth) hw[i] = W[i];
  if (tid < 4) hb[tid] = tid < N ? b[tid] : 0.f;
}

int pack_tail(int kind, const PackedLayout& L, const float* const* w, const float* const* b,
              char* packed, cudaStream_t st) {
  const GemmLayer* g = gemm_layers(kind);
  const int(*shp)[2] = layer_shapes(kind);
  for (int i = 0; i < num_gemm(kind); ++i) {
    float* wl = L.wlat[i] >= 0 ? (float*)(packed + L.wlat[i]) : nullptr;
    pack_tail_kernel<<<32, 256, 0, st>>>(w[g[i].src], b[g[i].src], shp[g[i].src][1], g[i],
                                         (float*)(packed + L.bias[i]), wl);
    AON_LAUNCH_CHECK();
  }
  const Head* h = heads(kind);
  for (int i = 0; i < num_heads(kind); ++i) {
    pack_head_kernel<<<4, 256, 0, st>>>(w[h[i].src], b[h[i].src], h[i].N, h[i].K,
                                        (float*)(packed + L.head_w[i]),
                                        (float*)(packed + L.head_b[i]));
    AON_LAUNCH_CHECK();
  }
  return AON_OK;
}

// folded[f + n] = bias[n] + sum_j Wlat[j][n] * lat[lat_off + j]
__global__ void fold_kernel(const float* __restrict__ bias, const float* __restrict__ wlat, int N,
                            int cnt, const float* __restrict__ lat, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int j = 0; j < cnt; ++j) acc = fmaf(wlat[(size_t)j * N + n], lat[j], acc);
  out[n] = bias[n] + acc;
}

__global__ void gather_latents_kernel(const float* __restrict__ shape,
                                      const float* __restrict__ app,
                                      const float* __restrict__ art, float* __restrict__ lat) {
  const int i = threadIdx.x;
  if (i < 128) lat[i] = shape[i];
  else if (i < 160) lat[i] = art[i - 128];
  else if (i < 288) lat[i] = app[i - 160];
}

// ---- A1+A2 ray generation -------------------------------------------------------------------------
__global__ void raygen_kernel(int H, int W, float focal, Cam c, float* __restrict__ rays_o,
                              float* __restrict__ rays_d) {
  const int n = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    float o[3], d[3];
    ray_from_camera(p, H, W, focal, c, o, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) { rays_d[3 * p + k] = d[k]; rays_o[3 * p + k] = o[k]; }
  }
}

// ---- A3 coarse sampling -------------------------------------------------------------------------------
__global__ void sample_along_rays_kernel(float near, float far, int n, const float* __restrict__ t_rand,
                                         int R, float* __restrict__ t_vals, RngDev rng) {
  const long total = (t_rand || rng.on) ? (long)R * n : n;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
       idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n);
    const float t = coarse_t(i, n, near, far);
    if (!t_rand && !rng.on) {
      t_vals[idx] = t;
      continue;
    }
    const float draw = t_rand ? t_rand[idx] : rng_uniform(rng, RNG_STREAM_STRATIFIED, (unsigned)(idx / n), (unsigned)i);
    // helper.py:122-127: stratified jitter between midpoints
    const float tp = i > 0 ? coarse_t(i - 1, n, near, far) : t;
    const float tn = i < n - 1 ? coarse_t(i + 1, n, near, far) : t;
    const float lower = i > 0 ? __fmul_rn(0.5f, __fadd_rn(t, tp)) : t;
    const float upper = i < n - 1 ? __fmul_rn(0.5f, __fadd_rn(tn, t)) : t;
    t_vals[idx] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), draw));
  }
}

// ---- A7 hierarchical sampling ----------------------------------------------------------------------------
// One warp per ray (sampling.cuh: sample_pdf_ray; the fused image kernel runs the same function in-kernel).
constexpr int PDF_WARPS = 8;

__global__ void __launch_bounds__(PDF_WARPS * 32)
sample_pdf_kernel(const float* __restrict__ t_coarse, long t_stride, const float* __restrict__ weights,
                  const float* __restrict__ u_in, long u_stride, int R, int nc, int nf,
                  float* __restrict__ t_fine, RngDev rng) {
  __shared__ float s_scr[PDF_WARPS][PDF_SCRATCH_FLOATS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long ray = (long)blockIdx.x * PDF_WARPS + warp; ray < R; ray += (long)gridDim.x * PDF_WARPS)
    sample_pdf_ray<false>(t_coarse + ray * t_stride, weights + ray * (long)nc, u_in ? u_in + ray * u_stride : nullptr, nc, nf,
                          s_scr[warp], t_fine + ray * (long)(nc + nf), lane, &rng, (unsigned)ray);
}

// the draws themselves, [rows, cols] of one stream (tests, debugging: the sampling kernels never materialise them)
__global__ void rng_uniform_kernel(RngDev rng, unsigned stream, int rows, int cols, float* __restrict__ out) {
  const long total = (long)rows * cols;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x)
    out[idx] = rng_uniform(rng, stream, (unsigned)(idx / cols), (unsigned)(idx % cols));
}
__global__ void rng_advance_kernel(unsigned long long* offset_dev, unsigned long long by) { *offset_dev += by; }

// rays of pixels [p0, p0 + n) of an H x W view (tail of a fused image render that takes the three-launch path)
__global__ void raygen_range_kernel(int H, int W, float focal, Cam c, long p0, int n, float* __restrict__ rays_o,
                                    float* __restrict__ rays_d) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float o[3], d[3];
    ray_from_camera(p0 + i, H, W, focal, c, o, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) { rays_d[3 * (size_t)i + k] = d[k]; rays_o[3 * (size_t)i + k] = o[k]; }
  }
}

// ---- workspace layout (aon_workspace_bytes) --------------------------------------------------------------------
// [slot mask 256 B | coarse t table 512 B | fused scratch slots | tail: rays 6 | w0 65 | t1 193 | coarse out 5 | seg 193 x 4
//  (floats per tail ray) | host staging: rays 9 + out 5 + coarse out 5 floats per ray]
struct WsLayout {
  size_t mask, t0, slots, tail_rays, tail_w0, tail_t1, tail_c5, tail_seg, seg_bytes, stage, total;
  long tail_cap;   // rays the three-launch path can take
};
static size_t up256(size_t x) { return (x + 255) / 256 * 256; }
static bool ws_layout(int precision, long R, WsLayout* L) {
  const bool tc = precision >= AON_PREC_TC_F16X3 && precision <= AON_PREC_TC_BF16;
  if (!tc && precision != AON_PREC_FP32) return false;
  if (R < 0) return false;
  size_t off = 0;
  L->mask = off; off += 256;
  L->t0 = off; off += 512;
  L->slots = off;
  if (tc) off += up256((size_t)MAX_SLOTS * 128 * (N_COARSE + N_TOTAL) * 4);
  // tensor-core modes: only the last partial wave (< MAX_SLOTS / 2 CTA pairs) or a batch smaller than one wave takes the
  // three-launch path; fp32: every ray does
  const long cap = tc ? (R < MAX_SLOTS / 2 * 256 ? R : MAX_SLOTS / 2 * 256) : R;
  L->tail_cap = cap;
  L->tail_rays = off; off += up256((size_t)cap * 6 * 4);
  L->tail_w0 = off; off += up256((size_t)cap * N_COARSE * 4);
  L->tail_t1 = off; off += up256((size_t)cap * N_TOTAL * 4);
  L->tail_c5 = off; off += up256((size_t)cap * 5 * 4);
  L->tail_seg = off;
  L->seg_bytes = tc ? up256(seg_bytes_tc((int)R, N_TOTAL)) : 0;
  off += L->seg_bytes;
  L->stage = off; off += up256((size_t)R * 19 * 4);
  L->total = off;
  return true;
}

size_t folded_floats_tc(int kind);                                                            // render_tc.cu
int fold_stages_tc(int kind, int precision, const PackedLayout& L, float* folded, cudaStream_t st);  // render_tc.cu

}  // namespace aon

using namespace aon;

extern "C" {

int aon_version(void) { return AON_ABI_VERSION; }
const char* aon_last_error(void) { return g_err; }
long aon_launch_count(int reset) {
  long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int aon_num_layers(int kind) {
  if (kind != AON_KIND_VANILLA && kind != AON_KIND_AUTODECODER) return AON_E_ARG;
  return num_layers(kind);
}

int aon_layer_shape(int kind, int i, int* out_features, int* in_features) {
  AON_REQUIRE(kind == AON_KIND_VANILLA || kind == AON_KIND_AUTODECODER, "bad kind %d", kind);
  AON_REQUIRE(i >= 0 && i < num_layers(kind) && out_features && in_features, "bad layer index %d", i);
  *out_features = layer_shapes(kind)[i][0];
  *in_features = layer_shapes(kind)[i][1];
  return AON_OK;
}

size_t aon_packed_bytes(int kind, int precision) {
  if (kind != AON_KIND_VANILLA && kind != AON_KIND_AUTODECODER) return 0;
  if (precision == AON_PREC_FP32) return (size_t)layout_fp32(kind).total_bytes;
  if (precision >= AON_PREC_TC_F16X3 && precision <= AON_PREC_TC_BF16)
    return (size_t)layout_tc(kind, precision).total_bytes;
  return 0;
}

int aon_pack_weights_tc(int kind, int precision, const float* const* w, const float* const* b,
                        void* packed, size_t packed_bytes, cudaStream_t st);  // render_tc.cu

int aon_pack_weights(int kind, int precision, const float* const* w, const float* const* b,
                     void* packed, size_t packed_bytes, aon_stream_t stream) {
  AON_REQUIRE(kind == AON_KIND_VANILLA || kind == AON_KIND_AUTODECODER, "bad kind %d", kind);
  AON_REQUIRE(w && b && packed, "null pointer");
  const size_t need = aon_packed_bytes(kind, precision);
  AON_REQUIRE(need > 0, "bad precision %d", precision);
  if (packed_bytes < need) {
    set_error("packed buffer too small: %zu < %zu", packed_bytes, need);
    return AON_E_SIZE;
  }
  AON_REQUIRE(((uintptr_t)packed & 255) == 0, "packed buffer must be 256-byte aligned");
  for (int i = 0; i < num_layers(kind); ++i) AON_REQUIRE(w[i] && b[i], "null layer pointer %d", i);
  cudaStream_t st = (cudaStream_t)stream;
  if (precision != AON_PREC_FP32)
    return aon_pack_weights_tc(kind, precision, w, b, packed, packed_bytes, st);
  const PackedLayout L = layout_fp32(kind);
  const GemmLayer* g = gemm_layers(kind);
  const int(*shp)[2] = layer_shapes(kind);
  for (int i = 0; i < num_gemm(kind); ++i) {
    pack_gemm_fp32_kernel<<<64, 256, 0, st>>>(w[g[i].src], shp[g[i].src][1], g[i],
                                             (float*)((char*)packed + L.w[i]));
    AON_LAUNCH_CHECK();
  }
  return pack_tail(kind, L, w, b, (char*)packed, st);
}

size_t aon_folded_floats(int kind) {
  // fp32 folded biases | latent staging | per-call bias stages of the tensor-core kernel (both CTA ranks)
  return kind == AON_KIND_AUTODECODER ? folded_floats_tc(kind) : 0;
}

int aon_fold_latents(int kind, int precision, const void* packed, const float* shape,
                     const float* appearance, const float* articulation, float* folded,
                     aon_stream_t stream) {
  AON_REQUIRE(kind == AON_KIND_AUTODECODER, "aon_fold_latents: kind must be AON_KIND_AUTODECODER");
  AON_REQUIRE(packed && shape && appearance && articulation && folded, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const PackedLayout L = precision == AON_PREC_FP32 ? layout_fp32(kind) : layout_tc(kind, precision);
  AON_REQUIRE(L.total_bytes > 0, "bad precision %d", precision);
  // the latent vector is staged at the tail of `folded` (caller provides A_FOLDED + A_LATENT floats)
  float* lat = folded + A_FOLDED_FLOATS;
  gather_latents_kernel<<<1, 288, 0, st>>>(shape, appearance, articulation, lat);
  AON_LAUNCH_CHECK();
  const GemmLayer* g = gemm_layers(kind);
  for (int i = 0; i < num_gemm(kind); ++i) {
    if (g[i].lat_col0 < 0) continue;
    fold_kernel<<<(g[i].N + 127) / 128, 128, 0, st>>>(
        (const float*)((const char*)packed + L.bias[i]), (const float*)((const char*)packed + L.wlat[i]),
        g[i].N, g[i].lat_cnt, lat + g[i].lat_off, folded + L.fold[i]);
    AON_LAUNCH_CHECK();
  }
  if (precision != AON_PREC_FP32) return fold_stages_tc(kind, precision, L, folded, st);
  return AON_OK;
}

int aon_raygen(int H, int W, float focal, const float* c2w_host, float* rays_o, float* rays_d,
               aon_stream_t stream) {
  AON_REQUIRE(H > 0 && W > 0 && focal > 0.f && c2w_host && rays_o && rays_d, "aon_raygen: bad argument");
  Cam c;
  memcpy(c.m, c2w_host, sizeof(c.m));
  const int n = H * W;
  raygen_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(H, W, focal, c, rays_o, rays_d);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

static RngDev rng_dev(const AonRng* r) {
  RngDev g;
  g.seed = r ? r->seed : 0; g.offset = r ? r->offset : 0; g.offset_dev = r ? r->offset_dev : nullptr; g.on = r != nullptr;
  return g;
}

static int sample_along_rays_impl(float near, float far, int n_points, const float* t_rand, const AonRng* rng, int R,
                                  float* t_vals, aon_stream_t stream) {
  const bool per_ray = t_rand != nullptr || rng != nullptr;
  AON_REQUIRE(n_points >= 2 && t_vals && (!per_ray || R > 0), "aon_sample_along_rays: bad argument");
  const long total = per_ray ? (long)R * n_points : n_points;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  sample_along_rays_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(near, far, n_points, t_rand, R, t_vals, rng_dev(rng));
  AON_LAUNCH_CHECK();
  return AON_OK;
}

int aon_sample_along_rays(float near, float far, int n_points, const float* t_rand, int R,
                          float* t_vals, aon_stream_t stream) {
  return sample_along_rays_impl(near, far, n_points, t_rand, nullptr, R, t_vals, stream);
}

int aon_sample_along_rays_rng(float near, float far, int n_points, const AonRng* rng, int R, float* t_vals, aon_stream_t stream) {
  AON_REQUIRE(rng != nullptr, "aon_sample_along_rays_rng: null rng");
  return sample_along_rays_impl(near, far, n_points, nullptr, rng, R, t_vals, stream);
}

int aon_rng_uniform(const AonRng* rng, int stream_id, int rows, int cols, float* out, aon_stream_t stream) {
  AON_REQUIRE(rng && out && rows >= 0 && cols >= 1 && (stream_id == 0 || stream_id == 1), "aon_rng_uniform: bad argument");
  if (rows == 0) return AON_OK;
  const long total = (long)rows * cols;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  rng_uniform_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rng_dev(rng), (unsigned)stream_id, rows, cols, out);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

int aon_rng_advance(unsigned long long* offset_dev, unsigned long long by, aon_stream_t stream) {
  AON_REQUIRE(offset_dev != nullptr, "aon_rng_advance: null pointer");
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(offset_dev, by);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

static int sample_pdf_impl(const float* t_coarse, long t_stride, const float* weights, const float* u, long u_stride,
                           const AonRng* rng, int R, int n_coarse, int n_fine, float* t_fine, aon_stream_t stream);

int aon_sample_pdf(const float* t_coarse, long t_stride, const float* weights, const float* u,
                   long u_stride, int R, int n_coarse, int n_fine, float* t_fine, aon_stream_t stream) {
  return sample_pdf_impl(t_coarse, t_stride, weights, u, u_stride, nullptr, R, n_coarse, n_fine, t_fine, stream);
}

int aon_sample_pdf_rng(const float* t_coarse, long t_stride, const float* weights, const AonRng* rng, int R, int n_coarse,
                       int n_fine, float* t_fine, aon_stream_t stream) {
  AON_REQUIRE(rng != nullptr, "aon_sample_pdf_rng: null rng");
  return sample_pdf_impl(t_coarse, t_stride, weights, nullptr, 0, rng, R, n_coarse, n_fine, t_fine, stream);
}

static int sample_pdf_impl(const float* t_coarse, long t_stride, const float* weights, const float* u, long u_stride,
                           const AonRng* rng, int R, int n_coarse, int n_fine, float* t_fine, aon_stream_t stream) {
  AON_REQUIRE(t_coarse && weights && t_fine, "aon_sample_pdf: null pointer");
  AON_REQUIRE(R >= 0 && n_coarse >= 4 && n_coarse <= PDF_MAX_COARSE && n_fine >= 1 && n_fine <= PDF_MAX_FINE,
              "aon_sample_pdf: unsupported sizes R=%d n_coarse=%d n_fine=%d", R, n_coarse, n_fine);
  AON_REQUIRE(t_stride == 0 || t_stride >= n_coarse, "aon_sample_pdf: bad t_stride");
  if (R == 0) return AON_OK;
  const int blocks = (R + PDF_WARPS - 1) / PDF_WARPS;
  sample_pdf_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, PDF_WARPS * 32, 0, (cudaStream_t)stream>>>(
      t_coarse, t_stride, weights, u, u_stride, R, n_coarse, n_fine, t_fine, rng_dev(rng));
  AON_LAUNCH_CHECK();
  return AON_OK;
}

size_t aon_workspace_bytes(int precision, int R) {
  WsLayout L;
  return ws_layout(precision, R, &L) ? L.total : 0;
}

// Three-launch path (coarse level, sample_pdf, fine level) over n rays given as device arrays; t0 as in aon_render_rays.
static int render_unfused(int kind, int precision, const void* pc, const void* pf, const float* fc, const float* ff,
                          const float* o, const float* d, const float* v, const float* t0, long t0_stride, const float* u,
                          long u_stride, int n, int white_bkgd, float* out5, float* coarse_out5, char* ws, const WsLayout& L,
                          const AonRenderOpts* opts, cudaStream_t st) {
  float* w0 = (float*)(ws + L.tail_w0);
  float* t1 = (float*)(ws + L.tail_t1);
  float* c5 = coarse_out5 ? coarse_out5 : (float*)(ws + L.tail_c5);
  void* seg = ws + L.tail_seg;
  int rc = render_level_any(kind, precision, pc, fc, o, d, v, t0, t0_stride, n, N_COARSE, white_bkgd, nullptr, nullptr,
                            nullptr, c5, w0, seg, L.seg_bytes, opts, st);
  if (rc != AON_OK) return rc;
  if ((rc = aon_sample_pdf(t0, t0_stride, w0, u, u_stride, n, N_COARSE, N_FINE, t1, (aon_stream_t)st)) != AON_OK) return rc;
  return render_level_any(kind, precision, pf, ff, o, d, v, t1, N_TOTAL, n, N_TOTAL, white_bkgd, nullptr, nullptr, nullptr,
                          out5, nullptr, seg, L.seg_bytes, opts, st);
}

// common body of aon_render_rays (cam == nullptr) and aon_render_image (rays_* == nullptr)
static int render_all(int kind, int precision, const void* pc, const void* pf, const float* fc, const float* ff,
                      const float* rays_o, const float* rays_d, const float* viewdirs, const Cam* cam, int H, int W,
                      float focal, long ray0, const float* t_coarse, const float* u, int R, float near, float far,
                      int white_bkgd, float* out, float* coarse_out, void* workspace, size_t workspace_bytes,
                      const AonRenderOpts* opts, cudaStream_t st, const char* who) {
  AON_REQUIRE(kind == AON_KIND_VANILLA || kind == AON_KIND_AUTODECODER, "%s: bad kind %d", who, kind);
  AON_REQUIRE(pc && pf && out, "%s: null pointer", who);
  AON_REQUIRE(kind == AON_KIND_VANILLA || (fc && ff), "%s: auto-decoder needs the folded biases of aon_fold_latents()", who);
  AON_REQUIRE(R >= 0, "%s: R must be non-negative", who);
  AON_REQUIRE((((uintptr_t)pc | (uintptr_t)pf) & 255) == 0, "%s: packed buffers must be 256-byte aligned", who);
  if (R == 0) return AON_OK;
  WsLayout L;
  if (!ws_layout(precision, R, &L)) { set_error("%s: bad precision %d", who, precision); return AON_E_ARG; }
  AON_REQUIRE(workspace && ((uintptr_t)workspace & 255) == 0, "%s: workspace must be a 256-byte aligned device buffer", who);
  if (workspace_bytes < L.total) {
    set_error("%s: workspace too small: %zu < aon_workspace_bytes() = %zu", who, workspace_bytes, L.total);
    return AON_E_SIZE;
  }
  char* ws = (char*)workspace;
  int rc;
  const float* t0 = t_coarse;
  long t0_stride = N_COARSE;
  if (t0 == nullptr) {
    float* tab = (float*)(ws + L.t0);
    if ((rc = aon_sample_along_rays(near, far, N_COARSE, nullptr, R, tab, (aon_stream_t)st)) != AON_OK) return rc;
    t0 = tab;
    t0_stride = 0;
  }
  const long u_stride = u ? N_FINE : 0;
  const bool tc = precision != AON_PREC_FP32;
  int R_main = 0, nsc = 1, nsf = 1;
  if (tc && !(opts && opts->no_fuse)) {
    int dev = 0, sms = 148;
    AON_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    R_main = (opts && opts->no_tail_split) ? R : tail_split_tc(R, sms, &nsc, &nsf);
    if (R - R_main > L.tail_cap) R_main = R;
  }
  if (R_main > 0) {
    AON_CUDA_CHECK(cudaMemsetAsync(ws + L.mask, 0, 256, st));
    rc = render_fused_tc(kind, precision, pc, pf, fc, ff, rays_o, rays_d, viewdirs, cam, H, W, focal, ray0, t0, t0_stride, u,
                         u_stride, R_main, N_COARSE, N_TOTAL, white_bkgd, out, coarse_out, (float*)(ws + L.slots),
                         (unsigned*)(ws + L.mask), MAX_SLOTS, opts, st);
    if (rc != AON_OK) return rc;
  }
  const int n = R - R_main;
  if (n == 0) return AON_OK;
  const float *o, *d, *v;
  if (cam) {
    float* ro = (float*)(ws + L.tail_rays);
    float* rd = ro + 3 * (size_t)n;
    raygen_range_kernel<<<(n + 255) / 256, 256, 0, st>>>(H, W, focal, *cam, ray0 + R_main, n, ro, rd);
    AON_LAUNCH_CHECK();
    o = ro; d = rd; v = rd;
  } else {
    o = rays_o + 3 * (size_t)R_main; d = rays_d + 3 * (size_t)R_main; v = viewdirs + 3 * (size_t)R_main;
  }
  return render_unfused(kind, precision, pc, pf, fc, ff, o, d, v, t0 + (size_t)R_main * t0_stride, t0_stride,
                        u ? u + (size_t)R_main * u_stride : nullptr, u_stride, n, white_bkgd, out + 5 * (size_t)R_main,
                        coarse_out ? coarse_out + 5 * (size_t)R_main : nullptr, ws, L, opts, st);
}

int aon_render_rays(int kind, int precision, const void* packed_coarse, const void* packed_fine,
                    const float* folded_coarse, const float* folded_fine, const float* rays_o,
                    const float* rays_d, const float* viewdirs, const float* t_coarse, const float* u,
                    int R, float near, float far, int white_bkgd, float* out, float* coarse_out,
                    void* workspace, size_t workspace_bytes, const AonRenderOpts* opts,
                    aon_stream_t stream) {
  AON_REQUIRE(R == 0 || (rays_o && rays_d && viewdirs), "aon_render_rays: null ray pointer");
  return render_all(kind, precision, packed_coarse, packed_fine, folded_coarse, folded_fine, rays_o, rays_d, viewdirs,
                    nullptr, 0, 0, 0.f, 0, t_coarse, u, R, near, far, white_bkgd, out, coarse_out, workspace,
                    workspace_bytes, opts, (cudaStream_t)stream, "aon_render_rays");
}

int aon_render_image(int kind, int precision, const void* packed_coarse, const void* packed_fine,
                     const float* folded_coarse, const float* folded_fine, const float* c2w_host,
                     float focal, int H, int W, long ray0, int R, float near, float far,
                     int white_bkgd, float* out, float* coarse_out, void* workspace,
                     size_t workspace_bytes, const AonRenderOpts* opts, aon_stream_t stream) {
  AON_REQUIRE(c2w_host && H > 0 && W > 0 && focal > 0.f, "aon_render_image: bad camera");
  AON_REQUIRE(ray0 >= 0 && R >= 0 && ray0 + R <= (long)H * W, "aon_render_image: pixel range [%ld, %ld) outside the %d x %d image",
              ray0, ray0 + R, H, W);
  Cam c;
  memcpy(c.m, c2w_host, sizeof(c.m));
  return render_all(kind, precision, packed_coarse, packed_fine, folded_coarse, folded_fine, nullptr, nullptr, nullptr, &c,
                    H, W, focal, ray0, nullptr, nullptr, R, near, far, white_bkgd, out, coarse_out, workspace,
                    workspace_bytes, opts, (cudaStream_t)stream, "aon_render_image");
}

int aon_render_image_host(int kind, int precision, const void* packed_coarse, const void* packed_fine,
                          const float* folded_coarse, const float* folded_fine,
                          const float* rays_o_host, const float* rays_d_host,
                          const float* viewdirs_host, int R, float near, float far, int white_bkgd,
                          float* out_host, float* coarse_out_host, void* workspace, size_t workspace_bytes,
                          const AonRenderOpts* opts, aon_stream_t stream) {
  AON_REQUIRE(packed_coarse && packed_fine && rays_o_host && rays_d_host && viewdirs_host && out_host,
              "aon_render_image_host: null pointer");
  AON_REQUIRE(R > 0, "aon_render_image_host: R must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  WsLayout L;
  if (!ws_layout(precision, R, &L)) { set_error("aon_render_image_host: bad precision %d", precision); return AON_E_ARG; }
  AON_REQUIRE(workspace && workspace_bytes >= L.total, "aon_render_image_host: workspace missing or smaller than aon_workspace_bytes()");
  float* f = (float*)((char*)workspace + L.stage);
  float* d_o = f;               f += (size_t)3 * R;
  float* d_d = f;               f += (size_t)3 * R;
  float* d_v = f;               f += (size_t)3 * R;
  float* d_out = f;             f += (size_t)5 * R;
  float* d_out0 = f;
  const size_t rb = (size_t)3 * R * 4;
  AON_CUDA_CHECK(cudaMemcpyAsync(d_o, rays_o_host, rb, cudaMemcpyHostToDevice, st));
  AON_CUDA_CHECK(cudaMemcpyAsync(d_d, rays_d_host, rb, cudaMemcpyHostToDevice, st));
  AON_CUDA_CHECK(cudaMemcpyAsync(d_v, viewdirs_host, rb, cudaMemcpyHostToDevice, st));
  int rc = aon_render_rays(kind, precision, packed_coarse, packed_fine, folded_coarse, folded_fine, d_o, d_d, d_v, nullptr,
                           nullptr, R, near, far, white_bkgd, d_out, coarse_out_host ? d_out0 : nullptr, workspace,
                           workspace_bytes, opts, stream);
  if (rc != AON_OK) return rc;
  AON_CUDA_CHECK(cudaMemcpyAsync(out_host, d_out, (size_t)R * 20, cudaMemcpyDeviceToHost, st));
  if (coarse_out_host)
    AON_CUDA_CHECK(cudaMemcpyAsync(coarse_out_host, d_out0, (size_t)R * 20, cudaMemcpyDeviceToHost, st));
  AON_CUDA_CHECK(cudaStreamSynchronize(st));
  return AON_OK;
}

}  // extern "C"
