// K4: 3-D masked-conv probability model -> per-symbol bits and per-image bit sums.
// src/probclass_imgcomp.py:63-106 (bitcost), :150-176 (masks), :185-196 (residual), :214-221
// (_ResShallow._logits), :258-260 (conv3d + bias + activation), :268-292 (padding).
// Volume axes: D = bottleneck channel, H, W; activations are channels-last (N,D,H,W,K).
// v1: one direct-conv kernel per layer (CUDA cores), weights staged in shared memory,
// dead (masked) taps skipped; layer 0 reads qbar with the centres[0] padding applied on the fly;
// the last layer fuses ReLU + log-sum-exp cross entropy + log2(e) and the per-image fp64 sum.
#include "common.cuh"
#include "conv_tc.cuh"
#include "pc_codec_common.cuh"

struct PcP {
  const float* in;     // layer input (N,Di,Hi,Wi,CIN) or qbar_nchw for the first layer
  const float* w;      // [2][3][3][CIN][COUT]
  const float* b;      // [COUT]
  const float* skip;   // residual source (N,Ds,Hs,Ws,COUT) or null (cropped [2:,2:-2,2:-2])
  float* out;          // (N,Do,Ho,Wo,COUT) or null
  const int64_t* sym;  // last layer only
  float* bits;         // last layer only, may be null (N,C,H,W)
  double* bits_sum;    // last layer only (N)
  int n, Di, Hi, Wi;   // input volume dims (for the first layer: padded dims)
  int c, hh, ww;       // bottleneck dims (first layer addressing)
  float pad_value;
  const float* pad_ptr; // when set, the pad value is read from device memory (centres[0])
  unsigned live_mask;  // bit t set = tap t (d*9+h*3+w) has non-zero mask
  int relu;
  int packed_taps;     // weights hold the live taps only, in tap order ([nlive][CIN][COUT])
  uint32_t* packed;    // EXACT last layer: (cumulative frequency << 16 | frequency) of the symbol at each position
  int L;
};

// EXACT = the operation order of the PC1 entropy coder (oracle/pc_codec.c): acc = bias, then the fmaf chain over
// live taps and input channels, ReLU as a compare-select; the last layer then emits the coder's frequencies.
template <int CIN, int COUT, bool FIRST, bool LAST, bool EXACT = false>
__global__ void __launch_bounds__(128) pc_conv3d_kernel(PcP p) {
  extern __shared__ float s_w[];  // [18 or live][CIN][COUT] + bias[COUT]
  const int ntaps_w = p.packed_taps ? __popc(p.live_mask) : 18;
  float* s_b = s_w + ntaps_w * CIN * COUT;
  for (int i = threadIdx.x; i < ntaps_w * CIN * COUT; i += blockDim.x) s_w[i] = p.w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_b[i] = p.b[i];
  __syncthreads();

  const int Do = p.Di - 1, Ho = p.Hi - 2, Wo = p.Wi - 2;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)p.n * Do * Ho * Wo;
  bool active = idx < total;
  int wq = 0, hq = 0, dq = 0, img = 0;
  if (active) {
    wq = (int)(idx % Wo);
    int64_t t = idx / Wo;
    hq = (int)(t % Ho);
    t /= Ho;
    dq = (int)(t % Do);
    img = (int)(t / Do);
  }
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = EXACT ? s_b[o] : 0.f;
  if (active) {
#pragma unroll
    for (int t = 0; t < 18; ++t) {
      if (!((p.live_mask >> t) & 1u)) continue;
      const int tw = p.packed_taps ? __popc(p.live_mask & ((1u << t) - 1u)) : t;  // slot of tap t in the weights
      int dd = t / 9, dh = (t / 3) % 3, dw = t % 3;
      int id = dq + dd, ih = hq + dh, iw = wq + dw;
      if (FIRST) {
        // padded volume (C+4, H+8, W+8), value = qbar or pad_value
        int cc = id - 4, yy = ih - 4, xx = iw - 4;
        float v = p.pad_ptr ? __ldg(p.pad_ptr) : p.pad_value;
        if (cc >= 0 && yy >= 0 && yy < p.hh && xx >= 0 && xx < p.ww)
          v = __ldg(p.in + (((int64_t)img * p.c + cc) * p.hh + yy) * p.ww + xx);
        const float* wt = s_w + tw * COUT;
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = fmaf(v, wt[o], acc[o]);
      } else {
        const float* ip = p.in + ((((int64_t)img * p.Di + id) * p.Hi + ih) * p.Wi + iw) * CIN;
        const float* wt = s_w + tw * CIN * COUT;
#pragma unroll 4
        for (int ci = 0; ci < CIN; ++ci) {
          float v = __ldg(ip + ci);
#pragma unroll
          for (int o = 0; o < COUT; ++o) acc[o] = fmaf(v, wt[ci * COUT + o], acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      float v = EXACT ? acc[o] : __fadd_rn(acc[o], s_b[o]);
      if (p.relu) v = EXACT ? pc1::relu(v) : fmaxf(v, 0.f);
      acc[o] = v;
    }
    if (p.skip) {
      // residual_input[..., 2:, 2:-2, 2:-2, :] -- skip volume is (Do+2, Ho+4, Wo+4)
      const float* sp = p.skip + ((((int64_t)img * (Do + 2) + dq + 2) * (Ho + 4) + hq + 2) * (Wo + 4) + wq + 2) * COUT;
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[o] = __fadd_rn(acc[o], sp[o]);
    }
  }
  if (!LAST) {
    if (active) {
      float* op = p.out + idx * COUT;
#pragma unroll
      for (int o = 0; o < COUT; ++o) op[o] = acc[o];
    }
  } else if (EXACT) {
    if (active) {
      uint32_t f[pc1::MAXL];
      pc1::logits_to_freqs(acc, COUT, f);
      const int64_t o_nchw = (((int64_t)img * Do + dq) * Ho + hq) * Wo + wq;
      const int sy = (int)p.sym[o_nchw];
      uint32_t cum = 0;
#pragma unroll
      for (int o = 0; o < COUT; ++o) cum += o < sy ? f[o] : 0u;
      uint32_t fs = f[0];
#pragma unroll
      for (int o = 1; o < COUT; ++o) fs = o == sy ? f[o] : fs;
      p.packed[o_nchw] = cum << 16 | fs;  // cum <= 65535, 1 <= f <= 65531
    }
  } else {
    // softmax cross entropy with the target symbol, in bits
    double mybits = 0.0;
    if (active) {
      float m = acc[0];
#pragma unroll
      for (int o = 1; o < COUT; ++o) m = fmaxf(m, acc[o]);
      float s = 0.f;
#pragma unroll
      for (int o = 0; o < COUT; ++o) s = __fadd_rn(s, expf(__fsub_rn(acc[o], m)));
      float lse = __fadd_rn(m, logf(s));
      int64_t o_nchw = (((int64_t)img * Do + dq) * Ho + hq) * Wo + wq;
      int sy = (int)p.sym[o_nchw];
      float picked = 0.f;
#pragma unroll
      for (int o = 0; o < COUT; ++o) picked = (o == sy) ? acc[o] : picked;
      float bit = __fmul_rn(__fsub_rn(lse, picked), 1.4426950408889634f);
      if (p.bits) p.bits[o_nchw] = bit;
      mybits = (double)bit;
    }
    // block reduction; a block never straddles images when Do*Ho*Wo % blockDim != 0 -> handle generally
    __shared__ double s_red[128];
    __shared__ int s_img[128];
    s_red[threadIdx.x] = mybits;
    s_img[threadIdx.x] = active ? img : -1;
    __syncthreads();
    if (threadIdx.x == 0) {
      int cur = -1;
      double sum = 0.0;
      for (int i = 0; i < (int)blockDim.x; ++i) {
        if (s_img[i] < 0) continue;
        if (s_img[i] != cur) {
          if (cur >= 0) atomicAdd(p.bits_sum + cur, sum);
          cur = s_img[i];
          sum = 0.0;
        }
        sum += s_red[i];
      }
      if (cur >= 0) atomicAdd(p.bits_sum + cur, sum);
    }
  }
}

static unsigned pc_live_mask(bool first) {
  // tap index t = d*9 + h*3 + w ; depth slice d=1 is the "current" slice
  unsigned m = 0;
  for (int d = 0; d < 2; ++d)
    for (int hh = 0; hh < 3; ++hh)
      for (int ww = 0; ww < 3; ++ww) {
        bool live = true;
        if (d == 1) {
          if (hh > 1) live = false;
          if (hh == 1 && (first ? ww >= 1 : ww > 1)) live = false;
        }
        if (live) m |= 1u << (d * 9 + hh * 3 + ww);
      }
  return m;
}

extern "C" int64_t dsin_probclass_workspace_bytes(int n, int c, int hh, int ww, int k) {
  int64_t v0 = (int64_t)(c + 3) * (hh + 6) * (ww + 6);
  int64_t v1 = (int64_t)(c + 2) * (hh + 4) * (ww + 4);
  int64_t v2 = (int64_t)(c + 1) * (hh + 2) * (ww + 2);
  return (v0 + v1 + v2) * n * k * (int64_t)sizeof(float) + 1024;
}

extern "C" int dsin_probclass_bits(dsin_handle_t h, const float* qbar, const int64_t* symbols, int n, int c,
                                   int hh, int ww, int k, int L, float pad_value, const float* w0,
                                   const float* b0, const float* w1, const float* b1, const float* w2,
                                   const float* b2, const float* w3, const float* b3, float* bits_nchw,
                                   double* bits_sum, void* workspace, void* stream) {
  DSIN_REQUIRE(h, qbar && symbols && bits_sum && workspace, "null pointer");
  DSIN_REQUIRE(h, k == 24 && L == 6, "only arch res_shallow with k=24, L=6 is built");
  DSIN_REQUIRE(h, w0 && b0 && w1 && b1 && w2 && b2 && w3 && b3, "null weights");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t v0 = (int64_t)n * (c + 3) * (hh + 6) * (ww + 6);
  int64_t v1 = (int64_t)n * (c + 2) * (hh + 4) * (ww + 4);
  int64_t v2 = (int64_t)n * (c + 1) * (hh + 2) * (ww + 2);
  float* a0 = (float*)workspace;
  float* a1 = a0 + v0 * k;
  float* a2 = a1 + v1 * k;
  if (cudaMemsetAsync(bits_sum, 0, sizeof(double) * n, st) != cudaSuccess)
    return dsin_fail(h, DSIN_ERR_CUDA, "%s: memset failed", __func__);
  const unsigned first = pc_live_mask(true), other = pc_live_mask(false);
  PcP p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.c = c; p.hh = hh; p.ww = ww; p.pad_value = pad_value;
  // layer 0: padded (c+4, hh+8, ww+8, 1) -> (c+3, hh+6, ww+6, 24), ReLU
  p.in = qbar; p.w = w0; p.b = b0; p.skip = nullptr; p.out = a0; p.Di = c + 4; p.Hi = hh + 8; p.Wi = ww + 8;
  p.live_mask = first; p.relu = 1;
  {
    size_t smem = (18 * 1 * 24 + 24) * sizeof(float);
    pc_conv3d_kernel<1, 24, true, false><<<(unsigned)((v0 + 127) / 128), 128, smem, st>>>(p);
    DSIN_LAUNCHED(h);
  }
  size_t smem24 = (18 * 24 * 24 + 24) * sizeof(float);
  // res1/conv1: ReLU
  p.in = a0; p.w = w1; p.b = b1; p.out = a1; p.Di = c + 3; p.Hi = hh + 6; p.Wi = ww + 6; p.live_mask = other;
  pc_conv3d_kernel<24, 24, false, false><<<(unsigned)((v1 + 127) / 128), 128, smem24, st>>>(p);
  DSIN_LAUNCHED(h);
  // res1/conv2: no activation, + cropped skip from a0
  p.in = a1; p.w = w2; p.b = b2; p.out = a2; p.skip = a0; p.Di = c + 2; p.Hi = hh + 4; p.Wi = ww + 4; p.relu = 0;
  pc_conv3d_kernel<24, 24, false, false><<<(unsigned)((v2 + 127) / 128), 128, smem24, st>>>(p);
  DSIN_LAUNCHED(h);
  // conv2: 24 -> 6, ReLU, fused cross entropy
  p.in = a2; p.w = w3; p.b = b3; p.out = nullptr; p.skip = nullptr; p.Di = c + 1; p.Hi = hh + 2; p.Wi = ww + 2;
  p.relu = 1; p.sym = symbols; p.bits = bits_nchw; p.bits_sum = bits_sum;
  {
    int64_t v3 = (int64_t)n * c * hh * ww;
    size_t smem = (18 * 24 * 6 + 6) * sizeof(float);
    pc_conv3d_kernel<24, 6, false, true><<<(unsigned)((v3 + 127) / 128), 128, smem, st>>>(p);
    DSIN_LAUNCHED(h);
  }
  return DSIN_OK;
}


// ---------------------------------------------------------------------------------------------
// tensor-core variant: the two 24->24 layers (88 % of the model's FLOPs) run on the generic tcgen05
// conv (3-D VALID mode, channels padded 24 -> 32, split fp16); the 1->24 stem and the 24->6 layer with
// its fused cross entropy stay on CUDA cores.
// ---------------------------------------------------------------------------------------------
// Stem of the tensor-core path: layer 0 (1 -> 24, the arithmetic of pc_conv3d_kernel<1, 24, true, false>) with
// both of its consumers' formats written at once -- fp32 (the skip of res1/conv2) and 32-channel split fp16 (the
// operand of res1/conv1).  A block owns 128 consecutive voxels and writes them as contiguous 16-byte pieces.
__global__ void __launch_bounds__(128) pc_stem_split_kernel(PcP p, __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float s_w[18 * 24 + 24];
  __shared__ __align__(16) float s_out[128 * 24];
  for (int i = threadIdx.x; i < 18 * 24 + 24; i += blockDim.x) s_w[i] = i < 18 * 24 ? p.w[i] : p.b[i - 18 * 24];
  __syncthreads();
  const float* s_b = s_w + 18 * 24;
  const int Do = p.Di - 1, Ho = p.Hi - 2, Wo = p.Wi - 2;
  const int64_t total = (int64_t)p.n * Do * Ho * Wo;
  const int64_t v0 = (int64_t)blockIdx.x * 128;
  const int64_t idx = v0 + threadIdx.x;
  if (idx < total) {
    const int wq = (int)(idx % Wo);
    int64_t t = idx / Wo;
    const int hq = (int)(t % Ho);
    t /= Ho;
    const int dq = (int)(t % Do), img = (int)(t / Do);
    float acc[24];
#pragma unroll
    for (int o = 0; o < 24; ++o) acc[o] = 0.f;
#pragma unroll
    for (int tp = 0; tp < 18; ++tp) {
      if (!((p.live_mask >> tp) & 1u)) continue;
      const int cc = dq + tp / 9 - 4, yy = hq + (tp / 3) % 3 - 4, xx = wq + tp % 3 - 4;
      float v = p.pad_value;
      if (cc >= 0 && yy >= 0 && yy < p.hh && xx >= 0 && xx < p.ww)
        v = __ldg(p.in + (((int64_t)img * p.c + cc) * p.hh + yy) * p.ww + xx);
      const float* wt = s_w + tp * 24;
#pragma unroll
      for (int o = 0; o < 24; ++o) acc[o] = fmaf(v, wt[o], acc[o]);
    }
#pragma unroll
    for (int o = 0; o < 24; ++o) s_out[threadIdx.x * 24 + o] = fmaxf(__fadd_rn(acc[o], s_b[o]), 0.f);
  }
  __syncthreads();
  const int nv = (int)min((int64_t)128, total - v0);
  float4* o4 = reinterpret_cast<float4*>(p.out + v0 * 24);
  for (int e = threadIdx.x; e < nv * 6; e += blockDim.x) o4[e] = reinterpret_cast<const float4*>(s_out)[e];
  uint4* h4 = reinterpret_cast<uint4*>(hi + v0 * 32);
  uint4* l4 = reinterpret_cast<uint4*>(lo + v0 * 32);
  for (int e = threadIdx.x; e < nv * 4; e += blockDim.x) {  // 8 channels per piece; channels 24..31 are zero padding
    const int vox = e / 4, c8 = (e % 4) * 8;
    uint4 uh = make_uint4(0u, 0u, 0u, 0u), ul = make_uint4(0u, 0u, 0u, 0u);
    if (c8 < 24) {
      __half2* hh2 = reinterpret_cast<__half2*>(&uh);
      __half2* ll2 = reinterpret_cast<__half2*>(&ul);
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        const float f0 = s_out[vox * 24 + c8 + 2 * k2], f1 = s_out[vox * 24 + c8 + 2 * k2 + 1];
        const __half h0 = __float2half_rn(f0), h1 = __float2half_rn(f1);
        hh2[k2] = __halves2half2(h0, h1);
        ll2[k2] = __halves2half2(__float2half_rn(f0 - __half2float(h0)), __float2half_rn(f1 - __half2float(h1)));
      }
    }
    h4[e] = uh;
    l4[e] = ul;
  }
}

// cross entropy of 6 ReLU'd logits against the target symbol, in bits, + per-image fp64 sums
__global__ void pc_cross_entropy_kernel(const float* __restrict__ logits, const int64_t* __restrict__ sym,
                                        float* __restrict__ bits, double* __restrict__ bits_sum, int64_t per_img,
                                        int n) {
  const int img = blockIdx.y;
  __shared__ double s_red[256];
  double acc = 0.0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < per_img; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = (int64_t)img * per_img + v;
    const float* l = logits + g * 6;
    float a[6];
#pragma unroll
    for (int o = 0; o < 6; ++o) a[o] = l[o];
    float m = a[0];
#pragma unroll
    for (int o = 1; o < 6; ++o) m = fmaxf(m, a[o]);
    float s = 0.f;
#pragma unroll
    for (int o = 0; o < 6; ++o) s = __fadd_rn(s, expf(__fsub_rn(a[o], m)));
    const float lse = __fadd_rn(m, logf(s));
    const int sy = (int)sym[g];
    float picked = 0.f;
#pragma unroll
    for (int o = 0; o < 6; ++o) picked = (o == sy) ? a[o] : picked;
    const float bit = __fmul_rn(__fsub_rn(lse, picked), 1.4426950408889634f);
    if (bits) bits[g] = bit;
    acc += (double)bit;
  }
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = blockDim.x / 2; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) s_red[threadIdx.x] += s_red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(bits_sum + img, s_red[0]);
}

static void pc_live_taps(ConvTc3dArgs& a) {
  a.ntaps = 0;
  for (int d = 0; d < 2; ++d)
    for (int hh = 0; hh < 3; ++hh)
      for (int ww = 0; ww < 3; ++ww) {
        if (d == 1 && (hh > 1 || (hh == 1 && ww > 1))) continue;  // other_mask
        a.tap_d[a.ntaps] = (short)d; a.tap_h[a.ntaps] = (short)hh; a.tap_w[a.ntaps] = (short)ww;
        a.tap_wi[a.ntaps] = (short)(d * 9 + hh * 3 + ww);
        a.ntaps++;
      }
  a.wtaps = 18;
}

extern "C" int64_t dsin_probclass_tc_workspace_bytes(int n, int c, int hh, int ww) {
  int64_t v0 = (int64_t)n * (c + 3) * (hh + 6) * (ww + 6);
  int64_t v1 = (int64_t)n * (c + 2) * (hh + 4) * (ww + 4);
  int64_t v2 = (int64_t)n * (c + 1) * (hh + 2) * (ww + 2);
  int64_t v3 = (int64_t)n * c * hh * ww;
  return v0 * 24 * 4 + v0 * 32 * 2 * 2 + v1 * 32 * 2 * 2 + v2 * 32 * 2 * 2 + v3 * 6 * 4 + 8192;
}

extern "C" int dsin_probclass_bits_tc(dsin_handle_t h, const float* qbar, const int64_t* symbols, int n, int c,
                                      int hh, int ww, float pad_value, const float* w0, const float* b0,
                                      const uint16_t* w1_hi, const uint16_t* w1_lo, const float* scale1,
                                      const float* shift1, const uint16_t* w2_hi, const uint16_t* w2_lo,
                                      const float* scale2, const float* shift2, const uint16_t* w3_hi,
                                      const uint16_t* w3_lo, const float* scale3, const float* shift3, int terms,
                                      float* bits_nchw, double* bits_sum, void* workspace, void* stream) {
  DSIN_REQUIRE(h, qbar && symbols && bits_sum && workspace, "null pointer");
  DSIN_REQUIRE(h, w0 && b0 && w1_hi && w1_lo && scale1 && shift1 && w2_hi && w2_lo && scale2 && shift2 && w3_hi &&
                      w3_lo && scale3 && shift3,
               "null weights");
  DSIN_REQUIRE(h, hh >= 8 && ww >= 16, "volume smaller than one tile");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t v0 = (int64_t)n * (c + 3) * (hh + 6) * (ww + 6);
  int64_t v1 = (int64_t)n * (c + 2) * (hh + 4) * (ww + 4);
  int64_t v2 = (int64_t)n * (c + 1) * (hh + 2) * (ww + 2);
  uint8_t* base = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float* a0 = (float*)base;                       base += (v0 * 24 * 4 + 255) / 256 * 256;
  __half* x1h = (__half*)base;                    base += (v0 * 32 * 2 + 255) / 256 * 256;
  __half* x1l = (__half*)base;                    base += (v0 * 32 * 2 + 255) / 256 * 256;
  __half* y1h = (__half*)base;                    base += (v1 * 32 * 2 + 255) / 256 * 256;
  __half* y1l = (__half*)base;                    base += (v1 * 32 * 2 + 255) / 256 * 256;
  __half* y2h = (__half*)base;                    base += (v2 * 32 * 2 + 255) / 256 * 256;
  __half* y2l = (__half*)base;                    base += (v2 * 32 * 2 + 255) / 256 * 256;
  float* logits = (float*)base;
  if (cudaMemsetAsync(bits_sum, 0, sizeof(double) * n, st) != cudaSuccess)
    return dsin_fail(h, DSIN_ERR_CUDA, "%s: memset failed", __func__);
  PcP p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.c = c; p.hh = hh; p.ww = ww; p.pad_value = pad_value;
  // layer 0 (CUDA cores): padded (c+4, hh+8, ww+8, 1) -> (c+3, hh+6, ww+6, 24), ReLU
  p.in = qbar; p.w = w0; p.b = b0; p.out = a0; p.Di = c + 4; p.Hi = hh + 8; p.Wi = ww + 8;
  p.live_mask = pc_live_mask(true); p.relu = 1;
  pc_stem_split_kernel<<<(unsigned)((v0 + 127) / 128), 128, 0, st>>>(p, x1h, x1l);
  DSIN_LAUNCHED(h);
  // res1/conv1 (tcgen05): ReLU, 32-channel (24 + 8 zero) split output
  ConvTc3dArgs a;
  memset(&a, 0, sizeof(a));
  pc_live_taps(a);
  a.vols = n; a.D = c + 3; a.H = hh + 6; a.W = ww + 6; a.kd = 2; a.kh = 3; a.kw = 3;
  a.cin = 32; a.cout = 32; a.terms = terms; a.act = DSIN_ACT_RELU;
  a.x_hi = (const uint16_t*)x1h; a.x_lo = (const uint16_t*)x1l; a.w_hi = w1_hi; a.w_lo = w1_lo;
  a.scale = scale1; a.shift = shift1; a.y_hi = (uint16_t*)y1h; a.y_lo = (uint16_t*)y1l;
  int rc = conv_tc_valid3d(h, a, st);
  if (rc != DSIN_OK) return rc;
  // res1/conv2 (tcgen05): no activation, + skip a0[2:, 2:-2, 2:-2] (fp32, 24 ch), 32-channel split output
  a.D = c + 2; a.H = hh + 4; a.W = ww + 4; a.cout = 32; a.act = DSIN_ACT_NONE;
  a.x_hi = (const uint16_t*)y1h; a.x_lo = (const uint16_t*)y1l; a.w_hi = w2_hi; a.w_lo = w2_lo;
  a.scale = scale2; a.shift = shift2; a.y_hi = (uint16_t*)y2h; a.y_lo = (uint16_t*)y2l; a.y_f32 = nullptr;
  a.r1f = a0; a.r1_d = c + 3; a.r1_oh = hh + 6; a.r1_ow = ww + 6; a.r1_dz = 2; a.r1_dy = 2; a.r1_dx = 2; a.r1_c = 24;
  rc = conv_tc_valid3d(h, a, st);
  if (rc != DSIN_OK) return rc;
  // conv2 (tcgen05): 24 -> 6, ReLU, fp32 logits; then the cross entropy kernel
  a.D = c + 1; a.H = hh + 2; a.W = ww + 2; a.cout = 6; a.act = DSIN_ACT_RELU;
  a.x_hi = (const uint16_t*)y2h; a.x_lo = (const uint16_t*)y2l; a.w_hi = w3_hi; a.w_lo = w3_lo;
  a.scale = scale3; a.shift = shift3; a.y_hi = nullptr; a.y_lo = nullptr; a.y_f32 = logits;
  a.r1f = nullptr; a.r1_c = 0;
  rc = conv_tc_valid3d(h, a, st);
  if (rc != DSIN_OK) return rc;
  {
    const int64_t per_img = (int64_t)c * hh * ww;
    int blocks = (int)((per_img + 255) / 256);
    if (blocks > 256) blocks = 256;
    pc_cross_entropy_kernel<<<dim3(blocks, n), 256, 0, st>>>(logits, symbols, bits_nchw, bits_sum, per_img, n);
    DSIN_LAUNCHED(h);
  }
  return DSIN_OK;
}


// ---------------------------------------------------------------------------------------------
// PC1 fast encoder, stage 1: with every symbol known, the context model needs no sequential schedule -- run the
// four layers over the whole volume (coder operation order) and keep only each symbol's own interval.
// ---------------------------------------------------------------------------------------------
int64_t pc1_symbol_tables_workspace(int n, int c, int hh, int ww) {
  return dsin_probclass_workspace_bytes(n, c, hh, ww, 24);
}

int pc1_symbol_tables(dsin_handle_t h, const float* qhard_nchw, const int64_t* symbols, int n, int c, int hh, int ww,
                      const float* centers, int L, const float* const* wb, uint32_t* packed, void* workspace,
                      cudaStream_t st) {
  DSIN_REQUIRE(h, L == 6, "the full-volume tables are built for 6 centres");
  const int k = 24;
  int64_t v0 = (int64_t)n * (c + 3) * (hh + 6) * (ww + 6);
  int64_t v1 = (int64_t)n * (c + 2) * (hh + 4) * (ww + 4);
  int64_t v2 = (int64_t)n * (c + 1) * (hh + 2) * (ww + 2);
  int64_t v3 = (int64_t)n * c * hh * ww;
  float* a0 = (float*)workspace;
  float* a1 = a0 + v0 * k;
  float* a2 = a1 + v1 * k;
  const unsigned first = pc_live_mask(true), other = pc_live_mask(false);
  PcP p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.c = c; p.hh = hh; p.ww = ww; p.pad_ptr = centers; p.packed_taps = 1; p.L = L;  // pad = centres[0]
  p.in = qhard_nchw; p.w = wb[0]; p.b = wb[1]; p.out = a0; p.Di = c + 4; p.Hi = hh + 8; p.Wi = ww + 8;
  p.live_mask = first; p.relu = 1;
  pc_conv3d_kernel<1, 24, true, false, true><<<(unsigned)((v0 + 127) / 128), 128, (13 * 24 + 24) * sizeof(float), st>>>(p);
  DSIN_LAUNCHED(h);
  const size_t smem24 = (14 * 24 * 24 + 24) * sizeof(float);
  p.in = a0; p.w = wb[2]; p.b = wb[3]; p.out = a1; p.Di = c + 3; p.Hi = hh + 6; p.Wi = ww + 6; p.live_mask = other;
  pc_conv3d_kernel<24, 24, false, false, true><<<(unsigned)((v1 + 127) / 128), 128, smem24, st>>>(p);
  DSIN_LAUNCHED(h);
  p.in = a1; p.w = wb[4]; p.b = wb[5]; p.out = a2; p.skip = a0; p.Di = c + 2; p.Hi = hh + 4; p.Wi = ww + 4; p.relu = 0;
  pc_conv3d_kernel<24, 24, false, false, true><<<(unsigned)((v2 + 127) / 128), 128, smem24, st>>>(p);
  DSIN_LAUNCHED(h);
  p.in = a2; p.w = wb[6]; p.b = wb[7]; p.out = nullptr; p.skip = nullptr; p.Di = c + 1; p.Hi = hh + 2; p.Wi = ww + 2;
  p.relu = 1; p.sym = symbols; p.packed = packed;
  pc_conv3d_kernel<24, 6, false, true, true><<<(unsigned)((v3 + 127) / 128), 128, (14 * 24 * 6 + 6) * sizeof(float), st>>>(p);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}
