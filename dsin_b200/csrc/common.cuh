// Shared host/device helpers for libdsin_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/dsin_b200.h"

struct dsin_handle_s {
  int device;
  int sm_count;
  int64_t launches;
  void* ident128;  // 128x128 fp16 identity (device), lazily created by the CTA-pair conv (residual adds on the MMA)
  char err[512];
};

static inline int dsin_fail(dsin_handle_t h, int code, const char* fmt, const char* a = "",
                            long b = 0, long c = 0) {
  if (h) snprintf(h->err, sizeof(h->err), fmt, a, b, c);
  return code;
}

#define DSIN_MAX_DEVICES 64

// A handle belongs to one device; every entry point must be called with that device current (its launches,
// tensor maps and per-device kernel attributes all assume it).
static inline bool dsin_device_is_current(dsin_handle_t h) {
  int dev = -1;
  return h && cudaGetDevice(&dev) == cudaSuccess && dev == h->device;
}

#define DSIN_REQUIRE(h, cond, msg)                                        \
  do {                                                                    \
    if (!dsin_device_is_current(h))                                       \
      return dsin_fail((h), DSIN_ERR_ARG, "%s: the handle's device is not the current CUDA device", __func__); \
    if (!(cond)) return dsin_fail((h), DSIN_ERR_ARG, "%s: " msg, __func__); \
  } while (0)

#define DSIN_LAUNCHED(h)                                                                  \
  do {                                                                                    \
    (h)->launches++;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return dsin_fail((h), DSIN_ERR_CUDA, "CUDA launch error: %s", cudaGetErrorString(e__)); \
  } while (0)

// KITTI normalisation constants (src/autoencoder_imgcomp.py:160-170, src/AE.py:240-250):
// float32(mean), sqrt(float32(var) + 1e-10) evaluated in float32.
__host__ __device__ inline float dsin_mean(int c) {
  return c == 0 ? 93.70454143384742f : (c == 1 ? 98.28243432206516f : 94.84678088809876f);
}
__host__ __device__ inline float dsin_var(int c) {
  return c == 0 ? 5411.79935676f : (c == 1 ? 5758.60456747f : 5890.31451232f);
}
__device__ __forceinline__ float dsin_std(int c) { return __fsqrt_rn(__fadd_rn(dsin_var(c), 1e-10f)); }
// SI-Finder divisors (src/siFinder.py:62-64)
__host__ __device__ inline float dsin_sif_div(int c) {
  return c == 0 ? 73.56493292844912f : (c == 1 ? 75.88547006820752f : 76.74838442810665f);
}

static inline int dsin_same_pad_before(int n, int k, int s, int d) {
  int out = (n + s - 1) / s;
  int keff = (k - 1) * d + 1;
  int total = (out - 1) * s + keff - n;
  if (total < 0) total = 0;
  return total / 2;
}
