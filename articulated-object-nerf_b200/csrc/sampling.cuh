// sampling.cuh -- device functions shared by the stand-alone kernels of aon_api.cu and the fused image
// kernel of render_tc.cu, so that both produce bit-identical values: ray generation (A1+A2), the coarse
// sample table (A3) and hierarchical sampling of one ray by one warp (A7).
#pragma once
#include "aon_common.cuh"

namespace aon {

struct Cam {
  float m[12];   // c2w, row-major [3,4]
};

// datasets/ray_utils.py:71-90 (get_ray_directions) + :118-159 (get_rays): pixel p = row * W + col ->
// origin o and unit direction d (the reference normalises in place, so rays_d == viewdirs).
__device__ __forceinline__ void ray_from_camera(long p, int H, int W, float focal, const Cam& c, float (&o)[3], float (&d)[3]) {
  const int row = (int)(p / W), col = (int)(p % W);
  // ray_utils.py:86-88: [(x - W/2)/f, -(y - H/2)/f, -1]
  const float dx = __fdiv_rn((float)col - 0.5f * (float)W, focal);
  const float dy = -__fdiv_rn((float)row - 0.5f * (float)H, focal);
  const float dz = -1.0f;
  // ray_utils.py:133: d_world = dirs @ c2w[:, :3].T
  const float wx = fmaf(dz, c.m[2], fmaf(dy, c.m[1], dx * c.m[0]));
  const float wy = fmaf(dz, c.m[6], fmaf(dy, c.m[5], dx * c.m[4]));
  const float wz = fmaf(dz, c.m[10], fmaf(dy, c.m[9], dx * c.m[8]));
  // ray_utils.py:146-147: in-place normalisation (rays_d aliases viewdirs)
  const float nrm = sqrtf(fmaf(wz, wz, fmaf(wy, wy, wx * wx)));
  d[0] = __fdiv_rn(wx, nrm);
  d[1] = __fdiv_rn(wy, nrm);
  d[2] = __fdiv_rn(wz, nrm);
  o[0] = c.m[3];
  o[1] = c.m[7];
  o[2] = c.m[11];
}

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

// ---- in-kernel random draws (helper.py:126 stratified jitter, helper.py:227 inverse-cdf draws) -----------------------------
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the generator behind torch.rand on CUDA),
// used as a pure FUNCTION of (seed, step offset, stream, row, column): a draw does not depend on the launch shape, so no
// [R,65] / [R,128] tensor of uniforms is ever written to HBM and a CUDA-graph replay only needs the step offset to advance
// in device memory.  key = seed (lo, hi); counter = (column / 4, row, offset lo, offset hi[0:30) | stream << 30); the draw of
// column c is word c % 4 of the block, mapped to [0, 1) as (x >> 8) * 2^-24.  oracle/philox.py restates it in numpy.
struct RngDev {
  unsigned long long seed, offset;
  const unsigned long long* offset_dev;   // optional: added to `offset` (the step counter of a captured training step)
  int on;
};
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float rng_uniform(const RngDev& g, unsigned stream, unsigned row, unsigned col) {
  const unsigned long long off = g.offset + (g.offset_dev ? *g.offset_dev : 0ull);
  const uint4 c = make_uint4(col >> 2, row, (unsigned)off, ((unsigned)(off >> 32) & 0x3FFFFFFFu) | (stream << 30));
  const uint4 x = philox4x32_10(c, make_uint2((unsigned)g.seed, (unsigned)(g.seed >> 32)));
  const unsigned w = (col & 3u) == 0 ? x.x : (col & 3u) == 1 ? x.y : (col & 3u) == 2 ? x.z : x.w;
  return (float)(w >> 8) * 5.9604644775390625e-08f;   // 2^-24: [0, 1)
}
constexpr unsigned RNG_STREAM_STRATIFIED = 0, RNG_STREAM_INVERSE_CDF = 1;

constexpr int PDF_MAX_COARSE = 65;
constexpr int PDF_MAX_FINE = 128;
constexpr int PDF_ST_FLOATS = PDF_MAX_COARSE + PDF_MAX_FINE + 3;   // per-warp scratch: values to sort
constexpr int PDF_SCRATCH_FLOATS = PDF_ST_FLOATS + 2 * PDF_MAX_COARSE;   // + bins + cdf

// sum of x[0 .. n) in the order of ATen's CPU sum kernel for a contiguous inner reduction
// (aten/src/ATen/native/cpu/SumKernel.cpp: vectorized_inner_sum -> row_sum -> multi_row_sum; the sum stub is built
// for AVX2 even on AVX-512 hosts, so the vector width is 8 floats): per vector lane l, four interleaved partial sums
// over the 8-wide vectors, the left-over vectors into partial 0, partials folded 0 <- 1, 2, 3; then, sequentially, the
// scalar tail x[8*(n/8) ..] and the eight lane partials.  weights[..., 1:-1].sum(-1) of helper.py:213 with n = 63.
// Every lane of the warp returns the sum.  (n < 8 takes ATen's scalar path, which is a different order; the reference
// only ever uses n = 63.)
__device__ __forceinline__ float aten_row_sum(const float* x, int n, int lane) {
  const int nvec = n >> 3, ilp = nvec >> 2;
  float ps0 = 0.f;
  if (lane < 8) {
    float ps[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < ilp; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) ps[k] = __fadd_rn(ps[k], x[(4 * i + k) * 8 + lane]);
    for (int i = 4 * ilp; i < nvec; ++i) ps[0] = __fadd_rn(ps[0], x[i * 8 + lane]);
    ps0 = __fadd_rn(__fadd_rn(__fadd_rn(ps[0], ps[1]), ps[2]), ps[3]);
  }
  float acc = 0.f;
  for (int k = nvec * 8; k < n; ++k) acc = __fadd_rn(acc, x[k]);
#pragma unroll
  for (int l = 0; l < 8; ++l) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, ps0, l));
  return acc;
}

// Hierarchical sampling of ONE ray by ONE warp: helper.py:203-252 + model.py:162-166.
//   tc [nc] coarse t of the ray; w [nc] compositing weights (the first and last are dropped here, model.py:165);
//   u [nf] inverse-cdf draws or nullptr = the deterministic table of helper.py:229 (or, with rng->on, draws generated here);
//   out [nc + nf] sorted.
//   scr: PDF_SCRATCH_FLOATS floats of shared memory owned by this warp.
// The bracket search (idx = #{cdf <= u}) is bit-equivalent to the reference's mask-max/min formulation
// (oracle: sorted_piecewise_constant_pdf_bracket); the reference's sort of the concatenated [t_coarse | samples] is a
// stable two-list merge when both lists are non-decreasing, with a rank-sort fallback for the unsorted u of
// randomized training or a 1-ulp inversion at a bracket boundary.
template <bool STREAM_LD>
__device__ __forceinline__ void sample_pdf_ray(const float* __restrict__ tc, const float* __restrict__ w,
                                               const float* __restrict__ u_in, int nc, int nf, float* scr,
                                               float* __restrict__ out, int lane, const RngDev* rng = nullptr,
                                               unsigned rng_row = 0) {
  float* st = scr;
  float* sb = scr + PDF_ST_FLOATS;
  float* sc = sb + PDF_MAX_COARSE;
  const int nb = nc - 1;  // bins = midpoints (64); interior weights = nb - 1 (63)
  const int nw = nb - 1;
  const int ntot = nc + nf;
  auto ldw = [&](int i) { return STREAM_LD ? __ldcg(w + i) : w[i]; };
  for (int i = lane; i < nc; i += 32) st[i] = tc[i];
  // interior weights w[1 .. nc-2] staged behind the samples' slots (free until the bracket search writes them)
  float* sw = st + nc;
  for (int i = lane; i < nw; i += 32) sw[i] = ldw(1 + i);
  __syncwarp();
  for (int i = lane; i < nb; i += 32) sb[i] = __fmul_rn(0.5f, __fadd_rn(st[i + 1], st[i]));
  const float wsum0 = aten_row_sum(sw, nw, lane);
  const float padding = fmaxf(0.f, __fsub_rn(1e-5f, wsum0));
  const float padw = __fdiv_rn(padding, (float)nw);
  const float wsum = __fadd_rn(wsum0, padding);
  // cdf[0]=0, cdf[k]=min(1, cumsum(pdf)[k-1]) for k=1..nw-1, cdf[nw]=1   (nw+1 = nb entries)
  // pdf in parallel, then torch.cumsum's CPU algorithm by lane 0: a sequential scan whose running sum is a DOUBLE
  // (ATen ReduceOpsKernel.cpp cumsum_cpu_kernel: at::acc_type<float, false> = double) rounded to fp32 per element.
  for (int k = lane; k < nw - 1; k += 32) sc[k + 1] = __fdiv_rn(__fadd_rn(sw[k], padw), wsum);
  __syncwarp();
  if (lane == 0) {
    double c = 0.0;
    sc[0] = 0.f;
    for (int k = 0; k < nw - 1; ++k) {
      c += (double)sc[k + 1];
      sc[k + 1] = fminf(1.0f, (float)c);
    }
    sc[nw] = 1.0f;
  }
  __syncwarp();
  for (int j = lane; j < nf; j += 32) {
    float u;
    if (u_in) {
      u = u_in[j];
    } else if (rng != nullptr && rng->on) {
      u = rng_uniform(*rng, RNG_STREAM_INVERSE_CDF, rng_row, (unsigned)j);   // helper.py:227: torch.rand(..., num_samples)
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
      if (STREAM_LD) __stcg(out + rank, v); else out[rank] = v;
    }
  } else {
    for (int i = lane; i < ntot; i += 32) {
      const float v = st[i];
      int rank = 0;
      for (int j = 0; j < ntot; ++j) {
        const float x = st[j];
        rank += (x < v) || (x == v && j < i);
      }
      if (STREAM_LD) __stcg(out + rank, v); else out[rank] = v;
    }
  }
  __syncwarp();
}

}  // namespace aon
