// K3: heatmap mask + 6-centre scalar quantiser, fused pointwise (HBM-bound, ~1.3 MB/image).
// Follows src/autoencoder_imgcomp.py:173-201 and src/quantizer_imgcomp.py:73-95 literally:
// symbols = first argmax of softmax(-1e7 * d), qbar = qsoft + (qhard - qsoft).
#include "common.cuh"

#define DSIN_MAX_CENTERS 16

__global__ void heatmap_quantize_kernel(const float* __restrict__ z33, const float* __restrict__ centers,
                                        int L, int64_t npix, int hw, int c, float* __restrict__ qbar_nhwc,
                                        float* __restrict__ qbar_nchw, int64_t* __restrict__ sym_nchw,
                                        float* __restrict__ qhard_nchw, float* __restrict__ z_nchw,
                                        float* __restrict__ heat_nchw) {
  __shared__ float s_c[DSIN_MAX_CENTERS];
  if (threadIdx.x < L) s_c[threadIdx.x] = centers[threadIdx.x];
  __syncthreads();
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over npix * c, channel fastest
  if (idx >= npix * c) return;
  int64_t pix = idx / c;
  int ch = (int)(idx % c);
  const float* zp = z33 + pix * (c + 1);
  float z0 = zp[0];
  // sigmoid(z0) * C
  float hm = __fmul_rn(__fdiv_rn(1.f, __fadd_rn(1.f, expf(-z0))), (float)c);
  float h3 = fmaxf(fminf(__fsub_rn(hm, (float)ch), 1.f), 0.f);
  float z = __fmul_rn(h3, zp[1 + ch]);

  float d[DSIN_MAX_CENTERS];
  float m_soft = -INFINITY, m_hard = -INFINITY;
#pragma unroll
  for (int j = 0; j < DSIN_MAX_CENTERS; ++j) {
    if (j < L) {
      float a = fabsf(__fsub_rn(z, s_c[j]));
      d[j] = __fmul_rn(a, a);
      m_soft = fmaxf(m_soft, -d[j]);
      m_hard = fmaxf(m_hard, __fmul_rn(-1e7f, d[j]));
    }
  }
  float s_soft = 0.f, s_hard = 0.f;
  float e_soft[DSIN_MAX_CENTERS], e_hard[DSIN_MAX_CENTERS];
#pragma unroll
  for (int j = 0; j < DSIN_MAX_CENTERS; ++j) {
    if (j < L) {
      e_soft[j] = expf(__fsub_rn(-d[j], m_soft));
      e_hard[j] = expf(__fsub_rn(__fmul_rn(-1e7f, d[j]), m_hard));
      s_soft = __fadd_rn(s_soft, e_soft[j]);
      s_hard = __fadd_rn(s_hard, e_hard[j]);
    }
  }
  float qsoft = 0.f, best = -1.f;
  int sym = 0;
#pragma unroll
  for (int j = 0; j < DSIN_MAX_CENTERS; ++j) {
    if (j < L) {
      qsoft = __fadd_rn(qsoft, __fmul_rn(__fdiv_rn(e_soft[j], s_soft), s_c[j]));
      float ph = __fdiv_rn(e_hard[j], s_hard);
      if (ph > best) {  // strict: first maximal index
        best = ph;
        sym = j;
      }
    }
  }
  float qhard = s_c[sym];
  float qbar = __fadd_rn(qsoft, __fsub_rn(qhard, qsoft));
  if (qbar_nhwc) qbar_nhwc[idx] = qbar;
  int64_t img = pix / hw;
  int64_t p_in = pix % hw;
  int64_t o = (img * c + ch) * hw + p_in;
  if (qbar_nchw) qbar_nchw[o] = qbar;
  if (sym_nchw) sym_nchw[o] = (int64_t)sym;
  if (qhard_nchw) qhard_nchw[o] = qhard;  // EncoderOutput.qhard / .z / .heatmap (src/autoencoder_imgcomp.py:239-245)
  if (z_nchw) z_nchw[o] = z;
  if (heat_nchw) heat_nchw[o] = h3;
}

extern "C" int dsin_heatmap_quantize(dsin_handle_t h, const float* z33, const float* centers, int L, int n,
                                     int hh, int ww, int c, float* qbar_nhwc, float* qbar_nchw,
                                     int64_t* symbols_nchw, float* qhard_nchw, float* z_nchw,
                                     float* heatmap_nchw, void* stream) {
  DSIN_REQUIRE(h, z33 && centers && n > 0 && hh > 0 && ww > 0 && c > 0, "bad argument");
  DSIN_REQUIRE(h, L >= 1 && L <= DSIN_MAX_CENTERS, "1 <= L <= 16");
  int64_t npix = (int64_t)n * hh * ww;
  int64_t tot = npix * c;
  heatmap_quantize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      z33, centers, L, npix, hh * ww, c, qbar_nhwc, qbar_nchw, symbols_nchw, qhard_nchw, z_nchw, heatmap_nchw);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}
