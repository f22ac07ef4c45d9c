// tcgen05 3x3 conv (placeholder until the tensor-core path lands).
#include "common.cuh"

extern "C" int dsin_pack_conv3x3_w(dsin_handle_t h, const float*, uint16_t*, uint16_t*, float*, int, int, void*) {
  return dsin_fail(h, DSIN_ERR_UNSUPPORTED, "%s: not built", __func__);
}
extern "C" int dsin_conv3x3_c128_tc(dsin_handle_t h, int, int, int, const uint16_t*, const uint16_t*,
                                    const uint16_t*, const uint16_t*, const float*, const float*, int,
                                    const uint16_t*, const uint16_t*, const uint16_t*, const uint16_t*,
                                    uint16_t*, uint16_t*, int, void*) {
  return dsin_fail(h, DSIN_ERR_UNSUPPORTED, "%s: not built", __func__);
}
