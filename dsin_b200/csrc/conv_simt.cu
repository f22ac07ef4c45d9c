// Generic fp32 implicit-GEMM convolution (CUDA cores) with fused folded-BN/bias, activation,
// residual adds and de-normalisation.  Used for the layers that are not (yet) on the tcgen05
// path: 5x5 stride-2 convs (h1, h2, to_bn), stride-2 transposed convs (from_bn, h12, h13) and the
// dilated SI-Net convs.  M = output pixels, N = cout, K = taps*cin (flattened).
#include "common.cuh"

struct ConvP {
  const float* x;
  const float* w;
  const float* scale;
  const float* shift;
  const float* res1;
  const float* res2;
  float* y;
  int n, h, w_, cin, cout, kh, kw, stride, dil, transposed, act, post;
  int oh, ow, pad_t, pad_l, K;
  int64_t M;
};

template <int TM, int TN>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvP p) {
  constexpr int TK = 16;
  constexpr int RM = TM / 16, RN = TN / 16;
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ int s_oy[TM], s_ox[TM], s_img[TM];

  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int co0 = blockIdx.y * TN;

  for (int m = tid; m < TM; m += 256) {
    int64_t mg = m0 + m;
    if (mg < p.M) {
      int ox = (int)(mg % p.ow);
      int64_t t = mg / p.ow;
      s_ox[m] = ox;
      s_oy[m] = (int)(t % p.oh);
      s_img[m] = (int)(t / p.oh);
    } else {
      s_img[m] = -1;
      s_oy[m] = 0;
      s_ox[m] = 0;
    }
  }
  __syncthreads();

  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

  const int kk_a = tid % TK;  // this thread's k within a chunk for A loads
  for (int k0 = 0; k0 < p.K; k0 += TK) {
    // ---- A tile: TM pixels x TK ks ----
    {
      int kg = k0 + kk_a;
      bool kvalid = kg < p.K;
      int tap = kvalid ? kg / p.cin : 0;
      int ci = kvalid ? kg % p.cin : 0;
      int ky = tap / p.kw, kx = tap % p.kw;
#pragma unroll
      for (int i = 0; i < (TM * TK) / 256; ++i) {
        int m = tid / TK + i * (256 / TK);
        float v = 0.f;
        int img = s_img[m];
        if (kvalid && img >= 0) {
          int iy, ix;
          bool ok;
          if (p.transposed) {
            int iy2 = s_oy[m] + p.pad_t - ky, ix2 = s_ox[m] + p.pad_l - kx;
            ok = iy2 >= 0 && ix2 >= 0 && !(iy2 & 1) && !(ix2 & 1);
            iy = iy2 >> 1;
            ix = ix2 >> 1;
            ok = ok && iy < p.h && ix < p.w_;
          } else {
            iy = s_oy[m] * p.stride + ky * p.dil - p.pad_t;
            ix = s_ox[m] * p.stride + kx * p.dil - p.pad_l;
            ok = iy >= 0 && ix >= 0 && iy < p.h && ix < p.w_;
          }
          if (ok) v = __ldg(p.x + (((int64_t)img * p.h + iy) * p.w_ + ix) * p.cin + ci);
        }
        As[kk_a][m] = v;
      }
    }
    // ---- B tile: TK ks x TN couts ----
#pragma unroll
    for (int i = 0; i < (TK * TN + 255) / 256; ++i) {
      int e = tid + i * 256;
      if (e < TK * TN) {
        int co = e % TN, kk = e / TN;
        int kg = k0 + kk, cg = co0 + co;
        Bs[kk][co] = (kg < p.K && cg < p.cout) ? __ldg(p.w + (int64_t)kg * p.cout + cg) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[RM], b[RN];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = As[kk][ty * RM + i];
#pragma unroll
      for (int j = 0; j < RN; ++j) b[j] = Bs[kk][tx * RN + j];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    int m = ty * RM + i;
    int64_t mg = m0 + m;
    if (mg >= p.M) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      int co = co0 + tx * RN + j;
      if (co >= p.cout) continue;
      float v = acc[i][j];
      float sc = p.scale ? p.scale[co] : 1.f;
      float sh = p.shift ? p.shift[co] : 0.f;
      v = __fadd_rn(__fmul_rn(v, sc), sh);
      if (p.act == DSIN_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == DSIN_ACT_LRELU02) v = fmaxf(__fmul_rn(v, 0.2f), v);
      int64_t o = mg * p.cout + co;
      if (p.res1) v = __fadd_rn(v, p.res1[o]);
      if (p.res2) v = __fadd_rn(v, p.res2[o]);
      if (p.post != DSIN_POST_NONE) {
        v = __fadd_rn(__fmul_rn(v, dsin_std(co)), dsin_mean(co));
        if (p.post == DSIN_POST_DENORM_CLIP) v = fminf(fmaxf(v, 0.f), 255.f);
      }
      p.y[o] = v;
    }
  }
}

extern "C" int dsin_conv2d(dsin_handle_t h, const dsin_conv_desc_t* d, const float* x, const float* w,
                           const float* scale, const float* shift, const float* res1,
                           const float* res2, float* y, void* stream) {
  DSIN_REQUIRE(h, d && x && w && y, "null pointer");
  DSIN_REQUIRE(h, d->n > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0 && d->kh > 0 && d->kw > 0,
               "bad shape");
  DSIN_REQUIRE(h, d->stride == 1 || d->stride == 2, "stride must be 1 or 2");
  DSIN_REQUIRE(h, d->dilation >= 1 && (d->dilation == 1 || d->stride == 1), "bad dilation");
  DSIN_REQUIRE(h, !d->transposed || (d->stride == 2 && d->dilation == 1), "transposed conv is stride 2");
  DSIN_REQUIRE(h, d->post == DSIN_POST_NONE || d->cout == 3, "denormalisation needs cout == 3");
  DSIN_REQUIRE(h, d->dilation_x == 0 || d->dilation_x == d->dilation, "anisotropic dilation: tensor-core path only");
  ConvP p;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.res1 = res1; p.res2 = res2; p.y = y;
  p.n = d->n; p.h = d->h; p.w_ = d->w; p.cin = d->cin; p.cout = d->cout; p.kh = d->kh; p.kw = d->kw;
  p.stride = d->stride; p.dil = d->dilation; p.transposed = d->transposed; p.act = d->act; p.post = d->post;
  if (d->transposed) {
    p.oh = 2 * d->h; p.ow = 2 * d->w;
    p.pad_t = dsin_same_pad_before(p.oh, d->kh, 2, 1);
    p.pad_l = dsin_same_pad_before(p.ow, d->kw, 2, 1);
  } else {
    p.oh = (d->h + d->stride - 1) / d->stride; p.ow = (d->w + d->stride - 1) / d->stride;
    p.pad_t = dsin_same_pad_before(d->h, d->kh, d->stride, d->dilation);
    p.pad_l = dsin_same_pad_before(d->w, d->kw, d->stride, d->dilation);
  }
  p.K = d->kh * d->kw * d->cin;
  p.M = (int64_t)d->n * p.oh * p.ow;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->cout > 32) {
    dim3 grid((unsigned)((p.M + 63) / 64), (d->cout + 63) / 64);
    conv_simt_kernel<64, 64><<<grid, 256, 0, st>>>(p);
  } else if (d->cout > 16) {
    dim3 grid((unsigned)((p.M + 127) / 128), 1);
    conv_simt_kernel<128, 32><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((unsigned)((p.M + 255) / 256), 1);
    conv_simt_kernel<256, 16><<<grid, 256, 0, st>>>(p);
  }
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}
