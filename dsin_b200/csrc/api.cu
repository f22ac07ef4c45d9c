// Handle management for libdsin_b200.
#include "common.cuh"

extern "C" {

int dsin_version(void) { return 100; }

int dsin_create(dsin_handle_t* out, int device) {
  if (!out) return DSIN_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count ||
      device >= DSIN_MAX_DEVICES)
    return DSIN_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DSIN_ERR_CUDA;
  if (prop.major != 10) return DSIN_ERR_UNSUPPORTED;  // sm_100a only: no fallback path exists
  dsin_handle_t h = new dsin_handle_s();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->launches = 0;
  h->ident128 = nullptr;
  h->err[0] = 0;
  *out = h;
  return DSIN_OK;
}

int dsin_destroy(dsin_handle_t h) {
  if (h && h->ident128) cudaFree(h->ident128);
  delete h;
  return DSIN_OK;
}

const char* dsin_last_error(dsin_handle_t h) { return h ? h->err : "null handle"; }

int64_t dsin_launch_count(dsin_handle_t h) { return h ? h->launches : -1; }

}  // extern "C"
