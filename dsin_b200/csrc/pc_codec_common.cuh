// Shared device code of the PC1 entropy coder: deterministic exp / frequency tables and the range encoder.
// The arithmetic is specified in oracle/pc_codec.c; csrc/pc_codec.cu (wavefront coder) and csrc/probclass.cu
// (full-volume frequency tables for the fast encoder) must agree with it bit for bit.
#pragma once
#include "common.cuh"

namespace pc1 {

constexpr int MAXL = 8;        // centres
constexpr uint32_t TOTAL_BITS = 16, TOTAL = 1u << TOTAL_BITS;

__device__ __forceinline__ float relu(float x) { return x > 0.0f ? x : 0.0f; }

__device__ __forceinline__ float exp_det(float x) {
  if (x < -80.0f) x = -80.0f;
  const float t = __fmul_rn(x, 1.4426950408889634f);
  const float n = floorf(t);
  const float f = __fsub_rn(t, n);
  float p = 1.5353362e-4f;
  p = __fmaf_rn(p, f, 1.3398874e-3f);
  p = __fmaf_rn(p, f, 9.6184370e-3f);
  p = __fmaf_rn(p, f, 5.5503324e-2f);
  p = __fmaf_rn(p, f, 2.4022648e-1f);
  p = __fmaf_rn(p, f, 6.9314720e-1f);
  p = __fmaf_rn(p, f, 1.0f);
  return __fmul_rn(p, __uint_as_float((uint32_t)((int)n + 127) << 23));
}

__device__ __forceinline__ void logits_to_freqs(const float* l, int L, uint32_t* f) {
  float m = l[0];
  int am = 0;
  for (int i = 1; i < L; ++i)
    if (l[i] > m) { m = l[i]; am = i; }
  float e[MAXL], Z = 0.0f;
  for (int i = 0; i < L; ++i) {
    e[i] = exp_det(__fsub_rn(l[i], m));
    Z = __fadd_rn(Z, e[i]);
  }
  const float scale = __fdiv_rn((float)(TOTAL - (uint32_t)L), Z);
  uint32_t sum = 0;
  for (int i = 0; i < L; ++i) {
    f[i] = 1u + (uint32_t)__fmul_rn(e[i], scale);
    sum += f[i];
  }
  f[am] += TOTAL - sum;
}

// ---------------------------------------------------------------- range coder (one thread per stream)
struct RcEnc {
  uint64_t low;
  uint32_t range;
  uint32_t cache;
  uint64_t cache_size;
  uint8_t* out;
  int64_t pos, cap;
  int overflow, skip_first;
  __device__ void init(uint8_t* o, int64_t c) {
    low = 0; range = 0xFFFFFFFFu; cache = 0; cache_size = 1; out = o; pos = 0; cap = c; overflow = 0; skip_first = 1;
  }
  __device__ void put(uint8_t b) {
    if (skip_first) { skip_first = 0; return; }
    if (pos < cap) out[pos] = b; else overflow = 1;
    pos++;
  }
  __device__ void shift_low() {
    if ((uint32_t)low < 0xFF000000u || (low >> 32) != 0) {
      const uint8_t carry = (uint8_t)(low >> 32);
      uint8_t c = (uint8_t)cache;
      do {
        put((uint8_t)(c + carry));
        c = 0xFF;
      } while (--cache_size != 0);
      cache = (uint32_t)((low >> 24) & 0xFF);
    }
    cache_size++;
    low = (low & 0x00FFFFFFull) << 8;
  }
  __device__ void encode(uint32_t cum, uint32_t freq) {
    const uint32_t r = range >> TOTAL_BITS;
    low += (uint64_t)r * cum;
    range = r * freq;
    while (range < (1u << 24)) { range <<= 8; shift_low(); }
  }
  __device__ void flush() {
    const uint64_t hi = low + range - 1;
    int k = 4;
    uint64_t v = 0;
    for (; k >= 0; --k) {
      v = hi & ~((1ull << (8 * k)) - 1);
      if (v >= low) break;
    }
    low = v;
    for (int i = 0; i < 5 - k; ++i) shift_low();
  }
};


}  // namespace pc1

// fast encoder, stage 1 (probclass.cu): the four layers over the whole volume in the coder's operation order,
// then per symbol (cumulative frequency << 16 | frequency) of the symbol that is there.
// weights: live-tap-major {w0[13][1][K], b0, w1[14][K][K], b1, w2[14][K][K], b2, w3[14][K][L], b3}
int pc1_symbol_tables(dsin_handle_t h, const float* qhard_nchw, const int64_t* symbols, int n, int c, int hh, int ww,
                      const float* centers, int L, const float* const* wb, uint32_t* packed, void* workspace,
                      cudaStream_t st);
int64_t pc1_symbol_tables_workspace(int n, int c, int hh, int ww);
