// Layout conversion + normalisation kernels (HBM-bound, trivially small).
#include "common.cuh"

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int c,
                                    int hw, int normalize) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*hw pixels
  if (idx >= (int64_t)n * hw) return;
  int img = (int)(idx / hw);
  int pix = (int)(idx % hw);
  for (int ch = 0; ch < c; ++ch) {
    float v = x[((int64_t)img * c + ch) * hw + pix];  // coalesced over pix
    if (normalize) v = __fdiv_rn(__fsub_rn(v, dsin_mean(ch)), dsin_std(ch));
    y[idx * c + ch] = v;
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int c,
                                    int hw) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * hw) return;
  int img = (int)(idx / hw);
  int pix = (int)(idx % hw);
  for (int ch = 0; ch < c; ++ch) y[((int64_t)img * c + ch) * hw + pix] = x[idx * c + ch];
}

__global__ void concat_normalize_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                        float* __restrict__ out, int64_t npix) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix) return;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    out[idx * 6 + ch] = __fdiv_rn(__fsub_rn(a[idx * 3 + ch], dsin_mean(ch)), dsin_std(ch));
    out[idx * 6 + 3 + ch] = __fdiv_rn(__fsub_rn(b[idx * 3 + ch], dsin_mean(ch)), dsin_std(ch));
  }
}

// concat([normalize(x_dec), normalize(y_syn)]) written as a 32-channel split-fp16 tensor (channels 6..31 = 0)
// so that the first SI-Net layer can run on the 32-channel tensor-core kernel.
__global__ void concat_normalize_split32_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                __half* __restrict__ hi, __half* __restrict__ lo, int64_t npix) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * 4) return;  // 4 x 16-byte pieces (8 channels) per pixel
  const int64_t pix = idx >> 2;
  const int piece = (int)(idx & 3);
  __half h[8], l[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) h[e] = l[e] = __float2half_rn(0.f);
  if (piece == 0) {
#pragma unroll
    for (int ch = 0; ch < 6; ++ch) {
      const float raw = ch < 3 ? a[pix * 3 + ch] : b[pix * 3 + ch - 3];
      const float v = __fdiv_rn(__fsub_rn(raw, dsin_mean(ch % 3)), dsin_std(ch % 3));
      h[ch] = __float2half_rn(v);
      l[ch] = __float2half_rn(v - __half2float(h[ch]));
    }
  }
  reinterpret_cast<uint4*>(hi)[idx] = *reinterpret_cast<const uint4*>(h);
  if (lo) reinterpret_cast<uint4*>(lo)[idx] = *reinterpret_cast<const uint4*>(l);
}

// NCHW image -> normalised, space-to-depth(2) NHWC split-fp16 with 32 channels: out[n, a, b, (sy*2+sx)*3 + c] =
// norm(x[n, c, 2a+sy, 2b+sx]) (channels 12..31 = 0).  Lets the 5x5 stride-2 stem conv run as a 3x3 stride-1
// tensor-core conv (src/autoencoder_imgcomp.py:136-144,223).
__global__ void nchw_to_s2d_split32_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                           __half* __restrict__ lo, int n, int hh, int ww) {
  const int h2 = hh / 2, w2 = ww / 2;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (img, a, b, piece) with 4 pieces of 8 ch
  if (idx >= (int64_t)n * h2 * w2 * 4) return;
  const int piece = (int)(idx & 3);
  int64_t t = idx >> 2;
  const int b = (int)(t % w2);
  t /= w2;
  const int a = (int)(t % h2);
  const int img = (int)(t / h2);
  __half h[8], l[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ch = piece * 8 + e;
    float v = 0.f;
    if (ch < 12) {
      const int c = ch % 3, s = ch / 3, sy = s >> 1, sx = s & 1;
      const float raw = x[(((int64_t)img * 3 + c) * hh + 2 * a + sy) * ww + 2 * b + sx];
      v = __fdiv_rn(__fsub_rn(raw, dsin_mean(c)), dsin_std(c));
    }
    h[e] = __float2half_rn(v);
    l[e] = __float2half_rn(v - __half2float(h[e]));
  }
  reinterpret_cast<uint4*>(hi)[idx] = *reinterpret_cast<const uint4*>(h);
  if (lo) reinterpret_cast<uint4*>(lo)[idx] = *reinterpret_cast<const uint4*>(l);
}

__global__ void f32_to_split_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                    __half* __restrict__ lo, int64_t count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float v = x[i];
  __half h = __float2half_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2half_rn(v - __half2float(h));
}

__global__ void split_to_f32_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                    float* __restrict__ y, int64_t count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  y[i] = lo ? __half2float(hi[i]) + __half2float(lo[i]) : __half2float(hi[i]);
}

extern "C" {

int dsin_nchw_to_nhwc(dsin_handle_t h, const float* x, float* y, int n, int c, int hh, int ww,
                      int normalize, void* stream) {
  DSIN_REQUIRE(h, x && y && n > 0 && c > 0 && hh > 0 && ww > 0, "bad argument");
  DSIN_REQUIRE(h, !normalize || c == 3, "normalisation needs 3 channels");
  int64_t tot = (int64_t)n * hh * ww;
  nchw_to_nhwc_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, c, hh * ww,
                                                                                   normalize);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

int dsin_nhwc_to_nchw(dsin_handle_t h, const float* x, float* y, int n, int c, int hh, int ww,
                      void* stream) {
  DSIN_REQUIRE(h, x && y && n > 0 && c > 0 && hh > 0 && ww > 0, "bad argument");
  int64_t tot = (int64_t)n * hh * ww;
  nhwc_to_nchw_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, c, hh * ww);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

int dsin_concat_normalize(dsin_handle_t h, const float* a, const float* b, float* out, int n, int hh,
                          int ww, void* stream) {
  DSIN_REQUIRE(h, a && b && out && n > 0, "bad argument");
  int64_t tot = (int64_t)n * hh * ww;
  concat_normalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, tot);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

int dsin_concat_normalize_split32(dsin_handle_t h, const float* a, const float* b, uint16_t* hi, uint16_t* lo, int n,
                                  int hh, int ww, void* stream) {
  DSIN_REQUIRE(h, a && b && hi && n > 0, "bad argument");
  int64_t tot = (int64_t)n * hh * ww;
  concat_normalize_split32_kernel<<<(unsigned)((tot * 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      a, b, (__half*)hi, (__half*)lo, tot);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

int dsin_nchw_to_s2d_split32(dsin_handle_t h, const float* x_nchw, uint16_t* hi, uint16_t* lo, int n, int hh, int ww,
                             void* stream) {
  DSIN_REQUIRE(h, x_nchw && hi && n > 0 && hh % 2 == 0 && ww % 2 == 0, "bad argument");
  int64_t tot = (int64_t)n * (hh / 2) * (ww / 2) * 4;
  nchw_to_s2d_split32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x_nchw, (__half*)hi, (__half*)lo, n, hh, ww);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

int dsin_f32_to_split(dsin_handle_t h, const float* x, uint16_t* hi, uint16_t* lo, int64_t count,
                      void* stream) {
  DSIN_REQUIRE(h, x && hi && count > 0, "bad argument");
  f32_to_split_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, (__half*)hi, (__half*)lo, count);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

int dsin_split_to_f32(dsin_handle_t h, const uint16_t* hi, const uint16_t* lo, float* y, int64_t count,
                      void* stream) {
  DSIN_REQUIRE(h, y && hi && count > 0, "bad argument");
  split_to_f32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)hi, (const __half*)lo, y, count);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

}  // extern "C"
