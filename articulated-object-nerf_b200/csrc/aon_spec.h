// aon_spec.h -- static description of the two MLPs of the reference as "GEMM layers" + small heads.
//
// Reference layer shapes: models/vanilla_nerf/model.py:65-93 (vanilla) and
// models/vanilla_nerf/model_autodecoder.py:95-169 (auto-decoder); SURVEY.md Appendix B.
// A GEMM layer reads up to two on-chip sources, concatenated along K exactly as the reference
// concatenates its inputs:  [ X (previous hidden, K1 rows) ; AUX (encoding / view-encoding / raw
// position, padded to Kaux rows) ], and -- for the auto-decoder -- a block of latent columns that is
// folded into a per-call bias (aon_fold_latents) because the codes are constant over a call.
#pragma once
#include <stdint.h>

namespace aon {

enum Aux : int { AUX_NONE = 0, AUX_E = 1, AUX_V = 2, AUX_P = 3 };
// padded row counts of the aux sources
constexpr int KE = 64;  // pos_enc(xyz, 0, 10): 63 -> 64
constexpr int KV = 32;  // pos_enc(viewdir, 0, 4): 27 -> 32
constexpr int KP = 16;  // raw xyz: 3 -> 16

struct GemmLayer {
  int src;       // index into the state_dict-ordered layer list of the kind
  int N;         // out features (128 or 256)
  int K1;        // rows taken from X (0, 128, 256); original columns [0, K1)
  int aux;       // Aux source
  int aux_col0;  // first original column of the aux block
  int aux_cnt;   // number of real aux columns (63 / 27 / 3)
  int lat_col0;  // first original column of the folded latent block (or -1)
  int lat_cnt;   // number of latent columns
  int lat_off;   // offset into the latent vector [shape(128) | articulation(32) | appearance(128)]
  int relu;
};

struct Head {
  int src;  // state_dict-ordered layer index
  int N;    // 1 or 3
  int K;    // 256 or 128
};

#ifdef __CUDACC__
#define AON_HD __host__ __device__
#else
#define AON_HD
#endif
AON_HD constexpr int kauxOf(int aux) { return aux == AUX_E ? KE : aux == AUX_V ? KV : aux == AUX_P ? KP : 0; }

// ---------------- vanilla ----------------
constexpr int V_NUM_LAYERS = 12;
constexpr int V_SHAPES[V_NUM_LAYERS][2] = {{256, 63},  {256, 256}, {256, 256}, {256, 256},
                                           {256, 256}, {256, 319}, {256, 256}, {256, 256},
                                           {128, 283}, {256, 256}, {1, 256},   {3, 128}};
constexpr int V_NUM_GEMM = 10;
constexpr GemmLayer V_GEMM[V_NUM_GEMM] = {
    {0, 256, 0, AUX_E, 0, 63, -1, 0, 0, 1},     {1, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {2, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1}, {3, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {4, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1}, {5, 256, 256, AUX_E, 256, 63, -1, 0, 0, 1},
    {6, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1}, {7, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {9, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 0},  // bottleneck_layer, no activation
    {8, 128, 256, AUX_V, 256, 27, -1, 0, 0, 1},  // views_linear.0
};
constexpr Head V_HEAD_DENSITY = {10, 1, 256};
constexpr Head V_HEAD_RGB = {11, 3, 128};

// ---------------- auto-decoder ----------------
constexpr int A_NUM_LAYERS = 20;
constexpr int A_SHAPES[A_NUM_LAYERS][2] = {
    {128, 163}, {128, 128}, {128, 128}, {128, 128}, {3, 128},   {256, 191}, {256, 256},
    {256, 256}, {256, 256}, {256, 256}, {256, 447}, {256, 256}, {256, 256}, {128, 411},
    {128, 128}, {128, 128}, {128, 128}, {256, 256}, {1, 256},   {3, 128}};
constexpr int A_NUM_GEMM = 17;
constexpr GemmLayer A_GEMM[A_NUM_GEMM] = {
    {0, 128, 0, AUX_P, 0, 3, 3, 160, 0, 1},  // deformations_linear.0: [xyz | shape | art]
    {1, 128, 128, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {2, 128, 128, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {3, 128, 128, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {5, 256, 0, AUX_E, 0, 63, 63, 128, 0, 1},  // pts_linears.0: [enc | shape]
    {6, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {7, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {8, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {9, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {10, 256, 256, AUX_E, 256, 63, 319, 128, 0, 1},  // pts_linears.5: [h | enc | shape]
    {11, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {12, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {17, 256, 256, AUX_NONE, 0, 0, -1, 0, 0, 0},       // bottleneck_layer
    {13, 128, 256, AUX_V, 256, 27, 283, 128, 160, 1},  // views_linear.0: [bott | view | appearance]
    {14, 128, 128, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {15, 128, 128, AUX_NONE, 0, 0, -1, 0, 0, 1},
    {16, 128, 128, AUX_NONE, 0, 0, -1, 0, 0, 1},
};
constexpr Head A_HEAD_DEFORM = {4, 3, 128};
constexpr Head A_HEAD_DENSITY = {18, 1, 256};
constexpr Head A_HEAD_RGB = {19, 3, 128};
constexpr int A_LATENT_FLOATS = 288;  // shape 128 | articulation 32 | appearance 128
constexpr int A_FOLDED_FLOATS = 128 + 256 + 256 + 128;

constexpr int MAX_GEMM = 17;
constexpr int MAX_HEADS = 3;

// Offsets (in floats for fp32 packing; in bytes for tensor-core packing) of everything the render
// kernel reads; computed on the host by layout_*() and passed to kernels by value.
struct PackedLayout {
  int64_t w[MAX_GEMM];      // weight block of GEMM layer i
  int64_t bias[MAX_GEMM];   // fp32 bias [N] (float index)
  int64_t wlat[MAX_GEMM];   // fp32 latent block [lat_cnt][N] (float index) or -1
  int fold[MAX_GEMM];       // float offset into `folded` or -1
  int64_t head_w[MAX_HEADS];  // fp32 [N][K] (float index)
  int64_t head_b[MAX_HEADS];  // fp32 [4]
  int64_t total_bytes;
};

}  // namespace aon
