// placeholder until the tcgen05 attention kernel lands
#include "common.cuh"
int attn_tc_fwd(const void*, const void*, int64_t, int, const void*, int64_t, void*, float*, int, int, int, int, float, int, cudaStream_t) {
  pmv_set_error("attention: tcgen05 kernel not built yet");
  return PMV_ERR_UNSUPPORTED;
}
