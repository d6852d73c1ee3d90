// aon_common.cuh -- shared host/device helpers for libaon_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/aon.h"
#include "aon_spec.h"

namespace aon {

// ---- error plumbing (thread-local message, never throws) ---------------------------------------
void set_error(const char* fmt, ...);
extern thread_local long g_launches;

#define AON_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::aon::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                         \
      return AON_E_CUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define AON_LAUNCH_CHECK()            \
  do {                                \
    ::aon::g_launches++;              \
    AON_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

#define AON_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) {                   \
      ::aon::set_error(__VA_ARGS__); \
      return AON_E_ARG;              \
    }                                \
  } while (0)

// ---- layouts -----------------------------------------------------------------------------------
int num_layers(int kind);
const int (*layer_shapes(int kind))[2];
int num_gemm(int kind);
const GemmLayer* gemm_layers(int kind);
int num_heads(int kind);
const Head* heads(int kind);  // order: [deform,] density, rgb
// fp32 (AON_PREC_FP32) packing: Wt[(K1+Kaux)][N] per GEMM layer, then bias, latent block, heads.
PackedLayout layout_fp32(int kind);
// tensor-core packing (defined in render_tc.cu)
PackedLayout layout_tc(int kind, int precision);

// ---- render entry points shared between the translation units ----------------------------------------------------
struct Cam;   // sampling.cuh
constexpr int MAX_SLOTS = 160;   // per-CTA scratch slots of the fused image kernel (>= SM count of any sm_100 part)
constexpr int N_COARSE = 65, N_FINE = 128, N_TOTAL = N_COARSE + N_FINE;   // samples per ray (model.py:128-129: 64 + 1, 128)
int render_level_tc(int kind, int precision, const void* packed, const float* folded, const float* rays_o,
                    const float* rays_d, const float* viewdirs, const float* t_vals, long t_stride, int R, int S,
                    int white_bkgd, float* comp_rgb, float* acc, float* depth, float* out5, float* weights, void* workspace,
                    size_t workspace_bytes, const AonRenderOpts* opts, cudaStream_t st);
int render_level_any(int kind, int precision, const void* packed, const float* folded, const float* rays_o,
                     const float* rays_d, const float* viewdirs, const float* t_vals, long t_stride, int R, int S,
                     int white_bkgd, float* comp_rgb, float* acc, float* depth, float* out5, float* weights,
                     void* workspace, size_t workspace_bytes, const AonRenderOpts* opts, cudaStream_t st);
int render_fused_tc(int kind, int precision, const void* packed_c, const void* packed_f, const float* folded_c,
                    const float* folded_f, const float* rays_o, const float* rays_d, const float* viewdirs, const Cam* cam,
                    int H, int W, float focal, long ray0, const float* t0, long t0_stride, const float* u, long u_stride,
                    int R, int S0, int S1, int white_bkgd, float* out5, float* coarse_out5, float* slots, unsigned* slot_mask,
                    int n_slots, const AonRenderOpts* opts, cudaStream_t st);
size_t seg_bytes_tc(int R, int S);
int tail_split_tc(int R, int sms, int* n_seg_coarse, int* n_seg_fine);

// ---- device math -------------------------------------------------------------------------------
// The reference encodes the cosine half as sin(fl32(2^k x) + fl32(pi/2)) (helper.py:139); 2^k x is
// exact in fp32, the +pi/2 is a rounded fp32 add.  __fadd_rn keeps the compiler from contracting.
#define AON_HALF_PI_F 1.57079637050628662109375f

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplusf_ref(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

}  // namespace aon
