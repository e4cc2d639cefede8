// tcgen05 / TMEM variant of the fused score kernel -- placeholder until the tensor-core path lands.
#include "common.h"
extern "C" size_t rt_score_bce_tc_ws_bytes(int, int, int) { return 0; }
extern "C" int rt_score_bce_tc(const float*, const float*, const float*, int, int, int, int, int, int,
                               const int32_t*, const int32_t*, float, double*, float*, float*, void*,
                               void*) {
  rt::set_error("rt_score_bce_fwd_bwd: variant 1 (tcgen05) is not built into this library yet");
  return 3;
}
