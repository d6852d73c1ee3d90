// render_tc.cu -- placeholder until the tcgen05 kernel lands.
#include "aon_common.cuh"
namespace aon {
PackedLayout layout_tc(int, int) { PackedLayout L; memset(&L, 0, sizeof(L)); return L; }
int render_level_tc(int, int, const void*, const float*, const float*, const float*, const float*, const float*,
                    long, int, int, int, float*, float*, float*, float*, cudaStream_t) {
  set_error("tensor-core precision modes are not built yet");
  return AON_E_UNSUPPORTED;
}
}
extern "C" int aon_pack_weights_tc(int, int, const float* const*, const float* const*, void*, size_t, cudaStream_t) {
  aon::set_error("tensor-core precision modes are not built yet");
  return AON_E_UNSUPPORTED;
}
