/* aon.h -- C ABI of libaon_b200.so: the B200-native (sm_100a) volume-rendering hot path of
 * zubair-irshad/articulated-object-nerf.
 *
 * The reference has NO native / FFI layer (it is 100 % Python on stock ATen ops);
 * its de-facto boundary is the Python call NeRF.forward(rays, randomized, white_bkgd, near, far)
 * (models/vanilla_nerf/model.py:147-199) and NeRF_AE_Art.forward(..., latents)
 * (models/vanilla_nerf/model_autodecoder.py:278-337).  Each entry point below replaces the group of
 * reference functions cited next to it; INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add to route those calls here.
 *
 * Conventions
 *   - every function returns 0 on success or a negative AON_E_* code; it never throws, exits or
 *     prints.  aon_last_error() returns a thread-local message for the last failure.
 *   - all tensor pointers are DEVICE pointers to contiguous row-major fp32 unless the name ends in
 *     _host; the caller allocates everything (no hidden allocation, no global mutable state).
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*; NULL = legacy
 *     default stream); no entry point synchronises unless it says so.
 *   - scratch memory is the CALLER's: entry points that need it take (workspace, workspace_bytes), sized by
 *     aon_workspace_bytes(); the library never allocates device memory and keeps no global switches
 *     (per-call knobs travel in AonRenderOpts).
 *   - the device is the calling thread's current CUDA device.
 */
#ifndef AON_H_
#define AON_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AON_ABI_VERSION 2

/* error codes */
#define AON_OK 0
#define AON_E_ARG (-1)      /* bad argument (null pointer, bad size, misaligned pointer)        */
#define AON_E_CUDA (-2)     /* a CUDA runtime call failed; see aon_last_error()                 */
#define AON_E_UNSUPPORTED (-3) /* valid request this build/device cannot serve (e.g. not sm_100) */
#define AON_E_SIZE (-4)     /* caller buffer too small                                          */

/* model kind: which of the reference's two hard-coded MLPs */
#define AON_KIND_VANILLA 0     /* models/vanilla_nerf/model.py:39-120            (12 linears)  */
#define AON_KIND_AUTODECODER 1 /* models/vanilla_nerf/model_autodecoder.py:60-239 (20 linears)  */

/* arithmetic mode of the MLP contraction */
#define AON_PREC_FP32 0    /* fp32 FFMA on CUDA cores (exact-order-free fp32; the parity anchor)   */
#define AON_PREC_TC_F16X3 1 /* tcgen05 kind::f16, operands split hi+lo fp16 (3 MMAs, ~22-bit)      */
#define AON_PREC_TC_F16 2  /* tcgen05 kind::f16, single fp16 operands (fast mode)                  */
#define AON_PREC_TC_BF16 3 /* tcgen05 kind::f16, single bf16 operands (fast mode, wide range)      */

typedef void* aon_stream_t; /* cudaStream_t */

int aon_version(void);
const char* aon_last_error(void);

/* Number of nn.Linear layers of a kind, in the reference's state_dict order
 * (vanilla: pts_linears.0-7, views_linear.0, bottleneck_layer, density_layer, rgb_layer;
 *  auto-decoder: deformations_linear.0-3, deformation_layer, pts_linears.0-7, views_linear.0-3,
 *  bottleneck_layer, density_layer, rgb_layer). */
int aon_num_layers(int kind);
/* out/in features of layer `i` of `kind` (the reference [out,in] weight shape). */
int aon_layer_shape(int kind, int i, int* out_features, int* in_features);

/* ---- weight packing -------------------------------------------------------------------------
 * Replaces nothing in the reference (it keeps nn.Linear [out,in] fp32); converts ONE MLP
 * (coarse_mlp or fine_mlp) from that layout into the kernel layout of `precision`.
 * w[i]/b[i]: device pointers to layer i's weight [out,in] and bias [out], i < aon_num_layers(kind).
 * `w`/`b` themselves are HOST arrays of device pointers. */
size_t aon_packed_bytes(int kind, int precision);
int aon_pack_weights(int kind, int precision, const float* const* w, const float* const* b,
                     void* packed, size_t packed_bytes, aon_stream_t stream);

/* ---- latent folding (auto-decoder only) --------------------------------------------------------
 * Replaces the einops.repeat broadcast + concat of the latent codes to every sample
 * (model_autodecoder.py:186-198, :226-228): the latent columns of deformations_linear.0,
 * pts_linears.0, pts_linears.5 and views_linear.0 are contracted with the codes ONCE per call into
 * per-call bias vectors.  shape/appearance [128], articulation [32] (device).  `folded` receives
 * aon_folded_floats(kind) floats and is passed to aon_render_level(). */
size_t aon_folded_floats(int kind);
int aon_fold_latents(int kind, int precision, const void* packed, const float* shape,
                     const float* appearance, const float* articulation, float* folded,
                     aon_stream_t stream);

/* ---- A1+A2  ray generation ---------------------------------------------------------------------
 * Replaces get_ray_directions + get_rays(output_view_dirs=True) (datasets/ray_utils.py:71-90,
 * :118-159).  c2w_host: 12 floats, row-major [3,4], read on the host at call time.
 * rays_o / rays_d: [H*W,3]; rays_d is unit-norm and doubles as viewdirs (the reference returns the
 * same values for both, ray_utils.py:146-147). */
int aon_raygen(int H, int W, float focal, const float* c2w_host, float* rays_o, float* rays_d,
               aon_stream_t stream);

/* ---- A3  coarse sampling -----------------------------------------------------------------------
 * Replaces sample_along_rays (helper.py:106-133), lindisp=False.  n_points = num_samples+1 (65).
 * t_rand == NULL: deterministic -> writes the shared table t_vals[n_points] (the reference
 * broadcasts it);  else t_rand [R,n_points] uniform -> stratified jitter, t_vals [R,n_points]. */
int aon_sample_along_rays(float near, float far, int n_points, const float* t_rand, int R,
                          float* t_vals, aon_stream_t stream);

/* In-kernel random draws for the two randomized sampling steps of training (helper.py:126 `torch.rand(batch, num_samples+1)`,
 * helper.py:227 `torch.rand(..., num_samples)`): Philox4x32-10 evaluated as a function of (seed, offset, stream, ray, column)
 * inside the sampling kernels -- no [R,65] / [R,128] tensor of uniforms exists.  The effective step offset is
 * offset + *offset_dev (offset_dev may be NULL); a training step captured in a CUDA graph keeps its step counter in device
 * memory and bumps it with aon_rng_advance inside the graph.  The bit stream is this library's (oracle/philox.py restates
 * it), not torch's: the reference draws from torch's global generator, whose stream depends on the launch shape.
 * stream ids: 0 = stratified jitter, 1 = inverse-cdf draws.  aon_rng_uniform writes the draws [rows, cols] themselves
 * (tests / debugging; the sampling kernels never materialise them). */
typedef struct AonRng {
  unsigned long long seed;
  unsigned long long offset;
  const unsigned long long* offset_dev;   /* device pointer or NULL */
} AonRng;
int aon_sample_along_rays_rng(float near, float far, int n_points, const AonRng* rng, int R,
                              float* t_vals, aon_stream_t stream);
int aon_rng_uniform(const AonRng* rng, int stream_id, int rows, int cols, float* out, aon_stream_t stream);
int aon_rng_advance(unsigned long long* offset_dev, unsigned long long by, aon_stream_t stream);

/* ---- workspace + per-call options ---------------------------------------------------------------------
 * aon_workspace_bytes: bytes of device scratch (256-byte aligned) that aon_render_level / aon_render_rays /
 * aon_render_image / aon_render_image_host need for up to R rays in `precision`; 0 for a bad precision.
 * It covers: the per-CTA scratch slots of the fused image kernel (coarse weights + fine sample positions of the
 * CTA's 128 rays, L2-resident, 160 slots x 132 KB), the per-sample (alpha, rgb) buffer of sample-segmented
 * launches (small ray batches, the last partial wave of a large one), the intermediates of the three-launch path
 * those rays take, and device staging for the _host call.
 * AonRenderOpts: optional per-call knobs for debugging and A/B timing; NULL = defaults.  Nothing is sticky. */
size_t aon_workspace_bytes(int precision, int R);
typedef struct AonRenderOpts {
  int force_segments;   /* > 0: aon_render_level cuts every ray's sample range into this many segments              */
  int no_tail_split;    /* 1: do not render the last partial wave of ray tiles in a separate, sample-segmented pass */
  int no_fuse;          /* 1: aon_render_rays / _image take the three-launch path (coarse, sample_pdf, fine)        */
  int reserved;
  float* dbg;           /* device [n_units][128][256]: pre-activation output of every GEMM unit, tile 0 / sample 0  */
  int* err_flag;        /* device int: receives a code if a pipeline barrier ever times out                          */
  long long* timeline;  /* device [3][4][18][4]: SM-clock stamps of the pipeline roles of CTA 0                      */
} AonRenderOpts;

/* ---- A4+A5/A9+A6  one level: encode, MLP, activations, alpha compositing ---------------------------
 * Replaces, per level, cast_rays + pos_enc + NeRFMLP.forward + activations + volumetric_rendering
 * (helper.py:25-26,136-140,157-195; model.py:95-120,174-195; model_autodecoder.py:171-239,306-331).
 * t_vals: [R,S] (t_stride = S) or a shared table [S] (t_stride = 0).
 * folded: from aon_fold_latents() (auto-decoder) or NULL (vanilla).
 * Outputs: comp_rgb [R,3], acc [R], depth [R]; weights [R,S] may be NULL.
 * workspace may be NULL (small ray batches then run unsegmented: same values, lower occupancy). */
int aon_render_level(int kind, int precision, const void* packed, const float* folded,
                     const float* rays_o, const float* rays_d, const float* viewdirs,
                     const float* t_vals, long t_stride, int R, int S, int white_bkgd,
                     float* comp_rgb, float* acc, float* depth, float* weights, void* workspace,
                     size_t workspace_bytes, const AonRenderOpts* opts, aon_stream_t stream);

/* ---- A7  hierarchical sampling ------------------------------------------------------------------
 * Replaces the t_mids / weights[...,1:-1] slicing (model.py:162-166) + sample_pdf
 * (helper.py:203-252): t_coarse [R,n_coarse] (stride 0 = shared table), weights [R,n_coarse]
 * (the full compositing weights; the kernel drops the first and last itself),
 * u [R,n_fine] or [n_fine] (u_stride 0) or NULL = the reference's deterministic linspace;
 * t_fine [R, n_coarse+n_fine] sorted.  n_coarse <= 65, n_fine <= 128.
 * The 63-term weight sum follows ATen's CPU reduction order (see csrc/sampling.cuh), so t_fine is bit-equal
 * to the reference's on the same weights. */
int aon_sample_pdf(const float* t_coarse, long t_stride, const float* weights, const float* u,
                   long u_stride, int R, int n_coarse, int n_fine, float* t_fine,
                   aon_stream_t stream);
/* the same with the inverse-cdf draws u [R,n_fine] generated in the kernel (AonRng above, stream 1) */
int aon_sample_pdf_rng(const float* t_coarse, long t_stride, const float* weights, const AonRng* rng,
                       int R, int n_coarse, int n_fine, float* t_fine, aon_stream_t stream);

/* ---- A8/A11  the whole level loop of NeRF.forward in ONE kernel ---------------------------------------
 * Replaces NeRF.forward / NeRF_AE_Art.forward (model.py:147-199; model_autodecoder.py:278-337) and the
 * chunk loop of render_rays / render_rays_test around it (model.py:295-348; model_autodecoder.py:479-541):
 * coarse level (65 samples), hierarchical sampling, fine level (193 samples) of every ray inside one fused
 * kernel launch -- each CTA keeps the coarse weights and fine sample positions of its 128 rays in an
 * L2-resident scratch slot, so no [rays x samples] tensor is written to HBM.
 *   t_coarse  NULL = deterministic coarse table (eval); else [R,65] per-ray positions (randomized training draws)
 *   u         NULL = deterministic inverse-cdf table; else [R,128] uniform draws
 *   out       [R,5] = (r, g, b, acc, depth) of the fine level;  coarse_out [R,5] of the coarse level or NULL
 * Rays beyond the last full wave of CTA pairs (and whole batches smaller than one wave) take a
 * sample-segmented three-launch path through the workspace; results do not depend on the split. */
int aon_render_rays(int kind, int precision, const void* packed_coarse, const void* packed_fine,
                    const float* folded_coarse, const float* folded_fine, const float* rays_o,
                    const float* rays_d, const float* viewdirs, const float* t_coarse, const float* u,
                    int R, float near, float far, int white_bkgd, float* out, float* coarse_out,
                    void* workspace, size_t workspace_bytes, const AonRenderOpts* opts,
                    aon_stream_t stream);

/* Same, with ray generation (A1+A2; datasets/ray_utils.py:71-90,118-159) fused in: renders pixels
 * [ray0, ray0 + R) (row-major) of the H x W view of camera c2w_host (12 floats, [3,4] row-major, read on
 * the host at call time).  ray0 / R let N ranks render contiguous pixel blocks of one image. */
int aon_render_image(int kind, int precision, const void* packed_coarse, const void* packed_fine,
                     const float* folded_coarse, const float* folded_fine, const float* c2w_host,
                     float focal, int H, int W, long ray0, int R, float near, float far,
                     int white_bkgd, float* out, float* coarse_out, void* workspace,
                     size_t workspace_bytes, const AonRenderOpts* opts, aon_stream_t stream);

/* ---- A8/A11  whole-image render from HOST buffers ----------------------------------------------
 * aon_render_rays with HOST rays in and HOST pixels out: H2D copies, the fused render, D2H copies on
 * `stream`, then SYNCHRONISES it.  rays_*_host [R,3]; out_host [R,5] = (r,g,b,acc,depth) of the FINE
 * level; coarse_out_host [R,5] may be NULL.  packed_* / folded_* / workspace are device pointers. */
int aon_render_image_host(int kind, int precision, const void* packed_coarse,
                          const void* packed_fine, const float* folded_coarse,
                          const float* folded_fine, const float* rays_o_host,
                          const float* rays_d_host, const float* viewdirs_host, int R, float near,
                          float far, int white_bkgd, float* out_host, float* coarse_out_host,
                          void* workspace, size_t workspace_bytes, const AonRenderOpts* opts,
                          aon_stream_t stream);

/* ---- training path, stage 1: per-element stages with hand-written adjoints ------------
 * The MLP contractions of training_step stay library GEMMs for now; these entry points replace every
 * elementwise / per-ray torch op around them and their autograd adjoints.
 *
 * aon_pos_enc / _backward: pos_enc(x, 0, max_deg) of helper.py:136-140 for x [n,3] -> out [n, 3+6*max_deg]
 * and its adjoint g_x [n,3] (needed by the auto-decoder, whose encoded position depends on the
 * deformation MLP, model_autodecoder.py:203-207). */
int aon_pos_enc(const float* x, long n, int max_deg, float* out, aon_stream_t stream);
int aon_pos_enc_backward(const float* x, const float* g_out, long n, int max_deg, float* g_x,
                         aon_stream_t stream);

/* aon_composite / _backward: activations (model.py:186-187; act_mode 1 = model_autodecoder.py:321-323)
 * + volumetric_rendering (helper.py:157-195) of MLP outputs raw_rgb [R,S,3], raw_sigma [R,S] ->
 * comp_rgb [R,3], acc [R], depth [R], weights [R,S] (may be NULL), trans [R,S] (exclusive transmittance;
 * may be NULL, needed by the backward).  The backward takes dL/dcomp_rgb (and optionally dL/dacc,
 * dL/ddepth; NULL = zero) and returns dL/draw_rgb [R,S,3], dL/draw_sigma [R,S]; t_vals carry no gradient
 * (the reference detaches the samples, helper.py:249). */
int aon_composite(const float* raw_rgb, const float* raw_sigma, const float* t_vals, long t_stride,
                  const float* dirs, int R, int S, int white_bkgd, int act_mode, float* comp_rgb,
                  float* acc, float* depth, float* weights, float* trans, aon_stream_t stream);
int aon_composite_backward(const float* raw_rgb, const float* raw_sigma, const float* t_vals,
                           long t_stride, const float* dirs, const float* weights, const float* trans,
                           const float* g_comp_rgb, const float* g_acc, const float* g_depth, int R,
                           int S, int white_bkgd, int act_mode, float* g_raw_rgb, float* g_raw_sigma,
                           aon_stream_t stream);

/* aon_adam_step: one torch.optim.Adam step (model.py:386-389: betas (0.9, 0.999), eps 1e-8, no weight
 * decay) over a FLAT parameter buffer; `step` counts from 1; lr is the value optimizer_step computed
 * (model.py:402-419); grads are multiplied by grad_scale first (1/world after a sum all-reduce). */
int aon_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n,
                  double lr, double beta1, double beta2, double eps, long step, double grad_scale,
                  aon_stream_t stream);
/* The same step with its seven step-dependent scalars (1-beta1, beta2, 1-beta2, eps, lr/bias_corr1, sqrt(bias_corr2),
 * grad_scale) read from DEVICE memory, for a training step captured in a CUDA graph (kernel arguments are baked into a
 * graph; the schedule is not): aon_adam_scalars fills the seven floats on the host exactly as aon_adam_step computes them,
 * the caller copies them to scalars7_dev before every replay. */
int aon_adam_scalars(double lr, double beta1, double beta2, double eps, long step, double grad_scale,
                     float* out7_host);
int aon_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n,
                      const float* scalars7_dev, aon_stream_t stream);

/* ---- training path, stage 3: the forward chain of one level in ONE launch ---------------------------------------------
 * Replaces, for training, cast_rays + pos_enc + NeRFMLP.forward of all R x S samples of a level (helper.py:25-26,136-140;
 * model.py:95-120; model_autodecoder.py:171-239) -- the same fused kernel as aon_render_level, in a mode that keeps each
 * layer's activations in shared memory / TMEM for the next layer and writes every layer output to HBM exactly ONCE, in the
 * packed plane layout the backward GEMMs of aon_gemm_tc read ([tile][feature/8][128][8] 16-bit hi (+ lo) planes scaled by 8,
 * plus a ReLU bit plane [rows][N/32]); no layer re-reads its input from HBM.  precision: AON_PREC_TC_F16X3 (hi + lo planes)
 * or AON_PREC_TC_F16 (hi only).  Row tile of (ray tile rt, sample s) = rt * S + s, row within the tile = ray % 128;
 * aon_train_tiles(R, S) = number of 128-row tiles = 2 * ceil(R / 256) * S (rows of rays >= R hold the last ray's values; the
 * backward must feed them zero gradients: aon_pack_rows_tiled does).  act_* / relu_bits are indexed by GEMM unit (vanilla:
 * pts_linears.0-7, bottleneck_layer, views_linear.0; auto-decoder: deformations_linear.0-3, pts_linears.0-7,
 * bottleneck_layer, views_linear.0-3); relu_bits entries may be NULL.  raw: [R*S,4] ray-major (rgb before the sigmoid, raw
 * density); warped (auto-decoder): [tiles*128,3] warped sample positions in tile order. */
typedef struct AonTrainDump {
  void* act_hi[28];
  void* act_lo[28];
  void* relu_bits[28];
  void* enc_hi;
  void* enc_lo;
  float* raw;
  float* warped;
} AonTrainDump;
int aon_train_tiles(int R, int S);
int aon_forward_train(int kind, int precision, const void* packed, const float* folded, const float* rays_o,
                      const float* rays_d, const float* viewdirs, const float* t_vals, long t_stride, int R, int S,
                      const AonTrainDump* dump, aon_stream_t stream);
/* fp32 -> packed planes in the tile order of aon_forward_train: packed row (tile rt * S + s, r) reads source row
 * (rt*128 + r) * S + s (per_ray = 0: src is [R*S, C]) or rt*128 + r (per_ray = 1: src is [R, C], replicated over the
 * samples); rows of rays >= R and columns >= C are zero. */
int aon_pack_rows_tiled(const float* src, long ld, int C, int R, int S, int per_ray, int c_pad, float scale,
                        void* hi, void* lo, aon_stream_t stream);
/* the inverse for fp32 matrices: dst[(ray * S + s), 0:C] = src[(tile, r), 0:C] for ray < R (src in tile order, row stride ld) */
int aon_unpack_rows_tiled(const float* src, long ld, int C, int R, int S, float* dst, aon_stream_t stream);

/* ---- training path, stage 2: tcgen05 GEMMs for the MLP's forward / dgrad / wgrad (csrc/gemm_tc.cu) ------------
 * Replaces the nn.Linear contractions of NeRFMLP.forward (model.py:99-118) and their autograd adjoints in
 * training_step (model.py:256-282).  Operands are 16-bit hi (+ lo) planes in two packed layouts:
 *   PK(rows, feat) = [rows/128][feat/8][128][8]   activations / gradients (rows = samples)
 *   PW(rows, k)    = [k/8][rows][8]               weights (rows = output rows of the GEMM, k = contraction)
 * AON_GEMM_NT:  D[128-row tile, N] = sum_seg A_seg[tile, a_off:+kext] . B_seg[b_row0:+N, b_off:+kext]^T   (forward, dgrad)
 *               A = PK tensors, B = PW tensors; up to AON_GEMM_MAX_SEG segments accumulate into one tile
 *               (skip / view concatenations, the two consumers of the last trunk activation).
 * AON_GEMM_TN:  D[128-feature tile of A, N] = sum_rows A[rows, a_off + 128 tile:+128]^T . B[rows, b_off:+N]    (wgrad)
 *               A, B = PK tensors read as MN-major operands; grid = (a_tiles, splits); every CTA writes an fp32 partial
 *               tile to partial[split][a_tiles*128][N]; aon_wgrad_reduce sums the splits in a fixed order.
 * Epilogues (NT): LINEAR  v = acc*inv_scale + bias, optional ReLU;   MASK  v = acc*inv_scale where the hi plane of the
 * forward activation mask_hi = PK(rows, mask_feat) is > 0 else 0 (ReLU adjoint).  v is written as fp32 row-major
 * out_f32[row, 0:n_valid] (row stride ldc) and/or as PK(rows, out_feat) planes at feature offset out_off, times out_scale;
 * colsum (optional) receives the per-tile column sums of v, summed over tiles by the caller in a fixed order.  A forward
 * GEMM can also emit the ReLU mask as a bit plane (relu_bits_out, 32 B per row for N = 256) that the dgrad GEMM reads
 * through mask_bits instead of re-reading a whole activation plane. */
#define AON_GEMM_MAX_SEG 2
#define AON_GEMM_NT 0
#define AON_GEMM_TN 1
#define AON_GEMM_EPI_LINEAR 0
#define AON_GEMM_EPI_MASK 1
#define AON_GEMM_EPI_PARTIAL 2
typedef struct AonGemm {
  int mode, epi, x3, nseg;     /* x3: 1 = hi+lo operands, 3 MMAs per K step (fp32-grade); 0 = hi planes only */
  int N, m_tiles;              /* output columns (multiple of 16, <= 256); number of 128-row tiles of the PK tensors */
  const void* a_hi[AON_GEMM_MAX_SEG];
  const void* a_lo[AON_GEMM_MAX_SEG];
  const void* b_hi[AON_GEMM_MAX_SEG];
  const void* b_lo[AON_GEMM_MAX_SEG];
  int a_feat[AON_GEMM_MAX_SEG], a_off[AON_GEMM_MAX_SEG], kext[AON_GEMM_MAX_SEG];
  int b_feat[AON_GEMM_MAX_SEG], b_off[AON_GEMM_MAX_SEG], b_row0[AON_GEMM_MAX_SEG];
  int a_tiles, splits, tiles_per_split, relu;
  float* partial;
  float inv_scale, out_scale;
  int n_valid, mask_feat, mask_off, out_feat, out_off, reserved;
  const float* bias;
  const void* mask_hi;
  float* out_f32;
  long ldc;
  void* out_hi;
  void* out_lo;
  float* colsum;               /* NT, optional: colsum[tile][N] = column sums of v over the tile's 128 rows (bias gradients) */
  uint32_t* relu_bits_out;     /* LINEAR + relu, optional (N % 32 == 0): bit c%32 of word [row][c/32] = (v[row][c] > 0) */
  const uint32_t* mask_bits;   /* MASK, optional alternative to mask_hi: the [rows][N/32] bit plane a forward GEMM wrote */
} AonGemm;
int aon_gemm_tc(const AonGemm* gemm, aon_stream_t stream);
/* sizeof(AonGemm) as this library was compiled: bindings in other languages check their mirror of the struct against it. */
size_t aon_gemm_struct_size(void);
/* fp32 [*, C] rows (row stride ld; packed row m reads source row m / row_div -- per-ray inputs broadcast to their
 * samples) -> PK(m_tiles*128, c_pad) hi (+ lo if non-NULL) times scale; rows >= M and columns >= C are zero. */
int aon_pack_rows(const float* src, long ld, int C, long M, int row_div, int m_tiles, int c_pad, float scale,
                  void* hi, void* lo, aon_stream_t stream);
/* nn.Linear weight [out, in] fp32 -> PW(r_pad, k_pad) hi (+ lo) times scale; transpose 0: rows = out, k = in
 * (forward); 1: rows = in, k = out (dgrad). */
int aon_pack_linear(const float* W, int out_features, int in_features, int transpose, int r_pad, int k_pad,
                    float scale, void* hi, void* lo, aon_stream_t stream);
/* number of rows of the `colsum` buffer an AON_GEMM_NT launch described by g writes (each row = N partial column sums, summed by
 * the caller in row order): one per row tile, or one per CTA when the persistent kernel takes the GEMM (large N = 128 / 256) */
int aon_gemm_colsum_rows(const AonGemm* g);
/* bias gradient from those partial rows: dst[c] = scale * sum_r partial[r][c] (rows summed in order: deterministic) */
int aon_colsum_finish(const float* partial, int rows, int N, float scale, float* dst, aon_stream_t stream);
/* dst[r, col_off + c] (flags & 1, transpose: dst[c, col_off + r]) = scale * sum_split partial[split][r][c], r < rows_valid,
 * c < cols_valid; flags & 2: added to dst instead of overwriting it (row-tile sub-batches of one backward pass) */
int aon_wgrad_reduce(const float* partial, int splits, int rows_pad, int N, float scale, float* dst, long ld,
                     int col_off, int rows_valid, int cols_valid, int flags, aon_stream_t stream);
/* column sums of a PK(m_tiles*128, feat) tensor (bias gradients): partial[split][feat], split = contiguous tile ranges */
int aon_colsum_packed(const void* hi, const void* lo, int feat, int m_tiles, int splits, float* partial,
                      aon_stream_t stream);

/* Number of kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
long aon_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* AON_H_ */
