// tcgen05 SI-Finder scorer (placeholder until the tensor-core path lands).
#include "sif_common.cuh"

int64_t sif_tc_workspace_bytes(int n, int hh, int ww, int ph, int pw, int method) { return 0; }

int sif_tc_match(dsin_handle_t h, const float*, const float*, const float*, const float*, int, int, int, int,
                 int, int, unsigned long long*, void*, cudaStream_t) {
  return dsin_fail(h, DSIN_ERR_UNSUPPORTED, "%s: tensor-core SI-Finder not built", __func__);
}
