// aon_api.cu -- C-ABI entry points that are not the render kernel itself: introspection, weight
// packing, latent folding, ray generation (A1+A2), coarse sampling (A3), hierarchical sampling (A7)
// and the host-buffer whole-image call (A8/A11).  See include/aon.h for the contract.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "aon_common.cuh"

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
struct Cam {
  float m[12];
};
__global__ void raygen_kernel(int H, int W, float focal, Cam c, float* __restrict__ rays_o,
                              float* __restrict__ rays_d) {
  const int n = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int row = p / W, col = p % W;
    // datasets/ray_utils.py:86-88: [(x - W/2)/f, -(y - H/2)/f, -1]
    const float dx = __fdiv_rn((float)col - 0.5f * (float)W, focal);
    const float dy = -__fdiv_rn((float)row - 0.5f * (float)H, focal);
    const float dz = -1.0f;
    // ray_utils.py:133: d_world = dirs @ c2w[:, :3].T
    float wx = fmaf(dz, c.m[2], fmaf(dy, c.m[1], dx * c.m[0]));
    float wy = fmaf(dz, c.m[6], fmaf(dy, c.m[5], dx * c.m[4]));
    float wz = fmaf(dz, c.m[10], fmaf(dy, c.m[9], dx * c.m[8]));
    // ray_utils.py:146-147: in-place normalisation (rays_d aliases viewdirs)
    const float nrm = sqrtf(fmaf(wz, wz, fmaf(wy, wy, wx * wx)));
    rays_d[3 * p + 0] = __fdiv_rn(wx, nrm);
    rays_d[3 * p + 1] = __fdiv_rn(wy, nrm);
    rays_d[3 * p + 2] = __fdiv_rn(wz, nrm);
    rays_o[3 * p + 0] = c.m[3];
    rays_o[3 * p + 1] = c.m[7];
    rays_o[3 * p + 2] = c.m[11];
  }
}

// ---- A3 coarse sampling -------------------------------------------------------------------------------
// torch.linspace(0,1,n) (fp32, CPU and CUDA): i < n/2 ? step*i : fma(-step, n-1-i, 1), step = 1/(n-1).
__device__ __forceinline__ float linspace01(int i, int n, float end) {
  const float step = __fdiv_rn(end, (float)(n - 1));
  return i < n / 2 ? __fmul_rn(step, (float)i) : fmaf(-step, (float)(n - 1 - i), end);
}
__device__ __forceinline__ float coarse_t(int i, int n, float near, float far) {
  const float s = linspace01(i, n, 1.0f);
  // helper.py:120: near * (1 - s) + far * s, three separately rounded fp32 ops
  return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, s)), __fmul_rn(far, s));
}
__global__ void sample_along_rays_kernel(float near, float far, int n, const float* __restrict__ t_rand,
                                         int R, float* __restrict__ t_vals) {
  const long total = t_rand ? (long)R * n : n;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
       idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n);
    const float t = coarse_t(i, n, near, far);
    if (!t_rand) {
      t_vals[idx] = t;
      continue;
    }
    // helper.py:122-127: stratified jitter between midpoints
    const float tp = i > 0 ? coarse_t(i - 1, n, near, far) : t;
    const float tn = i < n - 1 ? coarse_t(i + 1, n, near, far) : t;
    const float lower = i > 0 ? __fmul_rn(0.5f, __fadd_rn(t, tp)) : t;
    const float upper = i < n - 1 ? __fmul_rn(0.5f, __fadd_rn(tn, t)) : t;
    t_vals[idx] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[idx]));
  }
}

// ---- A7 hierarchical sampling ----------------------------------------------------------------------------
// One warp per ray.  helper.py:203-252 + model.py:162-166.  The bracket search
// (idx = #{cdf <= u}) is bit-equivalent to the reference's mask-max/min formulation
// (oracle: sorted_piecewise_constant_pdf_bracket); the reference's full sort of the concatenated
// [t_coarse | samples] is done as a rank sort (stable on ties), which is also correct for the
// unsorted u of randomized training.
constexpr int PDF_MAX_COARSE = 65;
constexpr int PDF_MAX_FINE = 128;
constexpr int PDF_WARPS = 8;

__global__ void __launch_bounds__(PDF_WARPS * 32)
sample_pdf_kernel(const float* __restrict__ t_coarse, long t_stride, const float* __restrict__ weights,
                  const float* __restrict__ u_in, long u_stride, int R, int nc, int nf,
                  float* __restrict__ t_fine) {
  __shared__ float s_t[PDF_WARPS][PDF_MAX_COARSE + PDF_MAX_FINE + 3];
  __shared__ float s_bins[PDF_WARPS][PDF_MAX_COARSE];
  __shared__ float s_cdf[PDF_WARPS][PDF_MAX_COARSE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = nc - 1;  // bins = midpoints (64); interior weights = nb - 1 (63)
  const int nw = nb - 1;
  const int ntot = nc + nf;
  for (long ray = (long)blockIdx.x * PDF_WARPS + warp; ray < R; ray += (long)gridDim.x * PDF_WARPS) {
    float* st = s_t[warp];
    float* sb = s_bins[warp];
    float* sc = s_cdf[warp];
    const float* tc = t_coarse + ray * t_stride;
    const float* w = weights + ray * (long)nc;
    for (int i = lane; i < nc; i += 32) st[i] = tc[i];
    __syncwarp();
    for (int i = lane; i < nb; i += 32) sb[i] = __fmul_rn(0.5f, __fadd_rn(st[i + 1], st[i]));
    // weight sum over the interior weights w[1 .. nc-2]
    float part = 0.f;
    for (int i = lane; i < nw; i += 32) part += w[1 + i];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const float wsum0 = part;
    const float padding = fmaxf(0.f, __fsub_rn(1e-5f, wsum0));
    const float padw = __fdiv_rn(padding, (float)nw);
    const float wsum = __fadd_rn(wsum0, padding);
    // cdf[0]=0, cdf[k]=min(1, cumsum(pdf)[k-1]) for k=1..nw-1, cdf[nw]=1   (nw+1 = nb entries)
    // pdf in parallel, then the sequential fp32 cumsum of torch.cumsum (CPU) by lane 0 (62 adds).
    for (int k = lane; k < nw - 1; k += 32) sc[k + 1] = __fdiv_rn(__fadd_rn(w[1 + k], padw), wsum);
    __syncwarp();
    if (lane == 0) {
      float c = 0.f;
      sc[0] = 0.f;
      for (int k = 0; k < nw - 1; ++k) {
        c = __fadd_rn(c, sc[k + 1]);
        sc[k + 1] = fminf(1.0f, c);
      }
      sc[nw] = 1.0f;
    }
    __syncwarp();
    for (int j = lane; j < nf; j += 32) {
      float u;
      if (u_in) {
        u = u_in[ray * u_stride + j];
      } else {
        // helper.py:229: linspace(0, 1 - 2^-32, nf); the end point rounds to 1.0f
        u = linspace01(j, nf, 1.0f);
      }
      // idx = #{k : cdf[k] <= u}; cdf is non-decreasing -> binary search
      int lo = 0, hi = nb;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sc[mid] <= u) lo = mid + 1; else hi = mid;
      }
      const int i0 = max(lo - 1, 0), i1 = min(lo, nb - 1);
      const float c0 = sc[i0], c1 = sc[i1], b0 = sb[i0], b1 = sb[i1];
      float t = __fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0));
      if (isnan(t)) t = 0.f;               // nan_to_num(., 0): 0/0 -> 0
      else if (isinf(t)) t = t > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
      t = fminf(fmaxf(t, 0.f), 1.f);
      st[nc + j] = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
    }
    __syncwarp();
    // The reference sorts cat[t_coarse, samples] (helper.py:250).  t_coarse is sorted; if the samples came out
    // non-decreasing too (always, up to rounding, for the deterministic u table) the stable sort is a merge:
    //   rank(coarse i) = i + #{samples < t_i},  rank(sample j) = #{coarse <= s_j} + j      (two binary searches)
    // otherwise (unsorted random u, or a 1-ulp inversion at a bracket boundary) fall back to a rank sort.
    float* out = t_fine + ray * (long)ntot;
    bool sorted = true;
    for (int j = lane; j < nf - 1; j += 32) sorted = sorted && (st[nc + j] <= st[nc + j + 1]);
    for (int i = lane; i < nc - 1; i += 32) sorted = sorted && (st[i] <= st[i + 1]);
    sorted = __all_sync(0xffffffffu, sorted);
    if (sorted) {
      for (int i = lane; i < ntot; i += 32) {
        const float v = st[i];
        int lo = 0, hi, rank;
        if (i < nc) {          // first sample index with s >= v
          hi = nf;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (st[nc + mid] < v) lo = mid + 1; else hi = mid; }
          rank = i + lo;
        } else {               // first coarse index with t > v
          hi = nc;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (st[mid] <= v) lo = mid + 1; else hi = mid; }
          rank = (i - nc) + lo;
        }
        out[rank] = v;
      }
    } else {
      for (int i = lane; i < ntot; i += 32) {
        const float v = st[i];
        int rank = 0;
        for (int j = 0; j < ntot; ++j) {
          const float x = st[j];
          rank += (x < v) || (x == v && j < i);
        }
        out[rank] = v;
      }
    }
    __syncwarp();
  }
}

__global__ void interleave5_kernel(const float* __restrict__ planes, int R, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  out[5 * (size_t)r + 0] = planes[3 * (size_t)r + 0];
  out[5 * (size_t)r + 1] = planes[3 * (size_t)r + 1];
  out[5 * (size_t)r + 2] = planes[3 * (size_t)r + 2];
  out[5 * (size_t)r + 3] = planes[(size_t)3 * R + r];
  out[5 * (size_t)r + 4] = planes[(size_t)4 * R + r];
}

// scratch arena for aon_render_image_host (per process, grown on demand, guarded by a mutex)
struct Arena {
  std::mutex mu;
  char* base = nullptr;
  size_t cap = 0;
  int device = -1;
};
static Arena g_arena;

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

int aon_sample_along_rays(float near, float far, int n_points, const float* t_rand, int R,
                          float* t_vals, aon_stream_t stream) {
  AON_REQUIRE(n_points >= 2 && t_vals && (t_rand == nullptr || R > 0), "aon_sample_along_rays: bad argument");
  const long total = t_rand ? (long)R * n_points : n_points;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  sample_along_rays_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(near, far, n_points, t_rand, R, t_vals);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

int aon_sample_pdf(const float* t_coarse, long t_stride, const float* weights, const float* u,
                   long u_stride, int R, int n_coarse, int n_fine, float* t_fine, aon_stream_t stream) {
  AON_REQUIRE(t_coarse && weights && t_fine, "aon_sample_pdf: null pointer");
  AON_REQUIRE(R >= 0 && n_coarse >= 4 && n_coarse <= PDF_MAX_COARSE && n_fine >= 1 && n_fine <= PDF_MAX_FINE,
              "aon_sample_pdf: unsupported sizes R=%d n_coarse=%d n_fine=%d", R, n_coarse, n_fine);
  AON_REQUIRE(t_stride == 0 || t_stride >= n_coarse, "aon_sample_pdf: bad t_stride");
  if (R == 0) return AON_OK;
  const int blocks = (R + PDF_WARPS - 1) / PDF_WARPS;
  sample_pdf_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, PDF_WARPS * 32, 0, (cudaStream_t)stream>>>(
      t_coarse, t_stride, weights, u, u_stride, R, n_coarse, n_fine, t_fine);
  AON_LAUNCH_CHECK();
  return AON_OK;
}

int aon_render_image_host(int kind, int precision, const void* packed_coarse, const void* packed_fine,
                          const float* folded_coarse, const float* folded_fine,
                          const float* rays_o_host, const float* rays_d_host,
                          const float* viewdirs_host, int R, float near, float far, int white_bkgd,
                          float* out_host, float* coarse_out_host, aon_stream_t stream) {
  AON_REQUIRE(packed_coarse && packed_fine && rays_o_host && rays_d_host && viewdirs_host && out_host,
              "aon_render_image_host: null pointer");
  AON_REQUIRE(R > 0, "aon_render_image_host: R must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const int S0 = 65, NF = 128, S1 = S0 + NF;
  // arena: rays 9R | t0 S0 | w0 R*S0 | t1 R*S1 | out0 5R | out1 5R   (floats)
  const size_t nfl = (size_t)R * (9 + S0 + S1 + 10) + 256;
  int dev = 0;
  AON_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_arena.mu);
  if (g_arena.device != dev || g_arena.cap < nfl * 4) {
    if (g_arena.base) cudaFree(g_arena.base);
    g_arena.base = nullptr;
    g_arena.cap = 0;
    AON_CUDA_CHECK(cudaMalloc(&g_arena.base, nfl * 4));
    g_arena.cap = nfl * 4;
    g_arena.device = dev;
  }
  float* f = (float*)g_arena.base;
  float* d_o = f;               f += (size_t)3 * R;
  float* d_d = f;               f += (size_t)3 * R;
  float* d_v = f;               f += (size_t)3 * R;
  float* d_t0 = f;              f += 256;
  float* d_w0 = f;              f += (size_t)R * S0;
  float* d_t1 = f;              f += (size_t)R * S1;
  float* d_out0 = f;            f += (size_t)5 * R;
  float* d_out1 = f;
  const size_t rb = (size_t)3 * R * 4;
  AON_CUDA_CHECK(cudaMemcpyAsync(d_o, rays_o_host, rb, cudaMemcpyHostToDevice, st));
  AON_CUDA_CHECK(cudaMemcpyAsync(d_d, rays_d_host, rb, cudaMemcpyHostToDevice, st));
  AON_CUDA_CHECK(cudaMemcpyAsync(d_v, viewdirs_host, rb, cudaMemcpyHostToDevice, st));
  int rc;
  if ((rc = aon_sample_along_rays(near, far, S0, nullptr, R, d_t0, stream)) != AON_OK) return rc;
  // outputs are packed [R,5] on the host side; on the device rgb [R,3] | acc [R] | depth [R]
  if ((rc = aon_render_level(kind, precision, packed_coarse, folded_coarse, d_o, d_d, d_v, d_t0, 0, R, S0,
                             white_bkgd, d_out0, d_out0 + (size_t)3 * R, d_out0 + (size_t)4 * R, d_w0,
                             stream)) != AON_OK)
    return rc;
  if ((rc = aon_sample_pdf(d_t0, 0, d_w0, nullptr, 0, R, S0, NF, d_t1, stream)) != AON_OK) return rc;
  if ((rc = aon_render_level(kind, precision, packed_fine, folded_fine, d_o, d_d, d_v, d_t1, S1, R, S1,
                             white_bkgd, d_out1, d_out1 + (size_t)3 * R, d_out1 + (size_t)4 * R, nullptr,
                             stream)) != AON_OK)
    return rc;
  // interleave (rgb | acc | depth) planes into the [R,5] host layout, then one contiguous D2H each
  float* d_pack = d_t1;  // t1 is dead after the fine level
  interleave5_kernel<<<(R + 255) / 256, 256, 0, st>>>(d_out1, R, d_pack);
  AON_LAUNCH_CHECK();
  AON_CUDA_CHECK(cudaMemcpyAsync(out_host, d_pack, (size_t)R * 20, cudaMemcpyDeviceToHost, st));
  if (coarse_out_host) {
    interleave5_kernel<<<(R + 255) / 256, 256, 0, st>>>(d_out0, R, d_pack + (size_t)5 * R);
    AON_LAUNCH_CHECK();
    AON_CUDA_CHECK(cudaMemcpyAsync(coarse_out_host, d_pack + (size_t)5 * R, (size_t)R * 20,
                                   cudaMemcpyDeviceToHost, st));
  }
  AON_CUDA_CHECK(cudaStreamSynchronize(st));
  return AON_OK;
}

}  // extern "C"
