// K1 on tensor cores: 3x3 stride-1 128->128 convolution as a tcgen05 implicit GEMM.
//   replaces slim.conv2d + batch_norm + ReLU + skip adds of the 64 trunk layers
//   (src/autoencoder_imgcomp.py:229-234,257-262,275-288).
//
// GEMM view per output tile: D[128 pixels x 128 couts] = sum over 9 taps x 2 channel halves of
//   A_tap[128 px x 64 ci] * B_tap[64 ci x 128 co].
// A comes straight from the NHWC fp16 activation by a 4-D TMA box (64 ch, 16 w, 8 h, 1 n) whose
// (w,h) origin is shifted by the tap; out-of-bounds box elements are zero-filled by TMA, which IS the
// TF 'SAME' zero padding.  B is the per-tap [co][ci] weight slab (K-major).  Both land in shared
// memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes; accumulators live in TMEM
// (2 x 128 columns, double buffered) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Precision ("terms"): activations and weights are carried as split fp16 pairs (v = hi + lo).
//   terms = 3:  hi*hi + hi*lo + lo*hi  -> ~22-bit operands, fp32-class result (parity mode)
//   terms = 1:  hi*hi only             -> fp16 operands (fast mode)
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2-5 = epilogue (TMEM -> registers -> scale/shift/act/residual -> split fp16 -> global).
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int BW = 16, BH = 8;          // spatial tile: 8 rows x 16 cols = 128 GEMM rows
constexpr int TILE_BYTES = 128 * 128;   // 128 rows x 64 fp16 = 16 KB
constexpr int NUM_KB = 18;              // 9 taps x 2 halves of the 128 input channels

struct TcP {
  const float* scale;
  const float* shift;
  const __half *r1h, *r1l, *r2h, *r2l;
  __half *yh, *yl;
  int n, H, W, act;
  int tiles_w, tiles_h, total_tiles;
};

template <int TERMS>
struct Cfg {
  static constexpr int kStages = TERMS == 3 ? 3 : 6;
  static constexpr int kStageBytes = (TERMS == 3 ? 4 : 2) * TILE_BYTES;
  static constexpr int kSmem = kStages * kStageBytes + 1024 /*align slack*/ + 2048 /*barriers, scale, shift*/;
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <int TERMS>
__global__ void __launch_bounds__(192, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                  const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl, TcP p) {
  constexpr int S = Cfg<TERMS>::kStages;
  constexpr int STAGE = Cfg<TERMS>::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* tiles = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * STAGE);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_scale = reinterpret_cast<float*>(tmem_ptr + 2);
  float* s_shift = s_scale + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
    prefetch_tmap(&tm_xh);
    prefetch_tmap(&tm_wh);
    if (TERMS == 3) {
      prefetch_tmap(&tm_xl);
      prefetch_tmap(&tm_wl);
    }
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int tw = tile % p.tiles_w, t2 = tile / p.tiles_w;
        const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
        const int ox0 = tw * BW, oy0 = th * BH;
        for (int kb = 0; kb < NUM_KB; ++kb) {
          const int tap = kb >> 1, cc = kb & 1;
          const int ky = tap / 3, kx = tap - 3 * ky;
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* st = tiles + stage * STAGE;
          mbar_expect_tx(&full[stage], STAGE);
          tma_load_4d(st, &tm_xh, &full[stage], cc * 64, ox0 + kx - 1, oy0 + ky - 1, n);
          tma_load_2d(st + TILE_BYTES, &tm_wh, &full[stage], cc * 64, tap * 128);
          if (TERMS == 3) {
            tma_load_4d(st + 2 * TILE_BYTES, &tm_xl, &full[stage], cc * 64, ox0 + kx - 1, oy0 + ky - 1, n);
            tma_load_2d(st + 3 * TILE_BYTES, &tm_wl, &full[stage], cc * 64, tap * 128);
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, 128, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&tempty[acc], aphase ^ 1u);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 128u;
        for (int kb = 0; kb < NUM_KB; ++kb) {
          mbar_wait(&full[stage], phase);
          fence_after_sync();
          const uint32_t sa = smem_u32(tiles + stage * STAGE);
          const uint64_t a_hi = make_smem_desc(sa, 16, 1024, LAYOUT_SW128);
          const uint64_t b_hi = make_smem_desc(sa + TILE_BYTES, 16, 1024, LAYOUT_SW128);
          const uint64_t a_lo = make_smem_desc(sa + 2 * TILE_BYTES, 16, 1024, LAYOUT_SW128);
          const uint64_t b_lo = make_smem_desc(sa + 3 * TILE_BYTES, 16, 1024, LAYOUT_SW128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x K=16 per 64-channel block; +32 B per step
            umma_f16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) ? 1u : 0u);
            if (TERMS == 3) {
              umma_f16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
              umma_f16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
            }
          }
          umma_commit(&empty[stage]);  // frees this smem stage once the MMAs above retire
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..5
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int hl = row >> 4, wl = row & 15;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const int tw = tile % p.tiles_w, t2 = tile / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      const int oy = th * BH + hl, ox = tw * BW + wl;
      const bool valid = oy < p.H && ox < p.W;
      const size_t pix = ((size_t)n * p.H + oy) * p.W + ox;
      mbar_wait(&tfull[acc], aphase);
      fence_after_sync();
#pragma unroll 1
      for (int chunk = 0; chunk < 4; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128 + chunk * 32), v);
        tmem_ld_wait();
        if (valid) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = chunk * 32 + j;
            float t = __fadd_rn(__fmul_rn(__uint_as_float(v[j]), s_scale[c]), s_shift[c]);
            f[j] = p.act == DSIN_ACT_RELU ? fmaxf(t, 0.f) : t;
          }
          const size_t off = pix * 128 + chunk * 32;
          if (p.r1h) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float a[8], b[8];
              unpack8(__ldg(reinterpret_cast<const uint4*>(p.r1h + off) + g), a);
              if (p.r1l) {
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.r1l + off) + g), b);
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = __fadd_rn(a[e], b[e]);
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) f[g * 8 + e] = __fadd_rn(f[g * 8 + e], a[e]);
            }
          }
          if (p.r2h) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float a[8], b[8];
              unpack8(__ldg(reinterpret_cast<const uint4*>(p.r2h + off) + g), a);
              if (p.r2l) {
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.r2l + off) + g), b);
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = __fadd_rn(a[e], b[e]);
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) f[g * 8 + e] = __fadd_rn(f[g * 8 + e], a[e]);
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 uh, ul;
            __half2* hh = reinterpret_cast<__half2*>(&uh);
            __half2* ll = reinterpret_cast<__half2*>(&ul);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x0 = f[g * 8 + 2 * e], x1 = f[g * 8 + 2 * e + 1];
              __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
              hh[e] = __halves2half2(h0, h1);
              ll[e] = __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
            }
            reinterpret_cast<uint4*>(p.yh + off)[g] = uh;
            if (p.yl) reinterpret_cast<uint4*>(p.yl + off)[g] = ul;
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

// weights [3][3][cin][cout] fp32 -> [tap][cout][cin] split fp16 with a per-cout power-of-two scale that
// moves the row's largest |w| into [8,16) so that the lo part stays in fp16's normal range.
__global__ void pack_w3x3_kernel(const float* __restrict__ w, __half* __restrict__ w_hi, __half* __restrict__ w_lo,
                                 float* __restrict__ wscale, int cin, int cout) {
  const int co = blockIdx.x;
  __shared__ float s_max[128];
  float m = 0.f;
  for (int i = threadIdx.x; i < 9 * cin; i += blockDim.x) m = fmaxf(m, fabsf(w[(size_t)i * cout + co]));
  s_max[threadIdx.x] = m;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) s_max[threadIdx.x] = fmaxf(s_max[threadIdx.x], s_max[threadIdx.x + s]);
    __syncthreads();
  }
  m = s_max[0];
  int ex = 0;
  float sc = 1.f;
  if (m > 0.f) {
    frexpf(m, &ex);           // m = f * 2^ex, f in [0.5,1)
    sc = ldexpf(1.f, 4 - ex); // m*sc in [8,16)
  }
  if (threadIdx.x == 0) wscale[co] = sc;
  for (int i = threadIdx.x; i < 9 * cin; i += blockDim.x) {
    int tap = i / cin, ci = i - tap * cin;
    float v = w[(size_t)i * cout + co] * sc;
    __half hi = __float2half_rn(v);
    size_t o = ((size_t)tap * cout + co) * cin + ci;
    w_hi[o] = hi;
    w_lo[o] = __float2half_rn(v - __half2float(hi));
  }
}

template <int TERMS>
int launch_tc(dsin_handle_t h, const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& wh,
              const CUtensorMap& wl, const TcP& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv3x3_tc_kernel<TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg<TERMS>::kSmem) != cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: cannot raise dynamic shared memory", __func__);
    configured = true;
  }
  int grid = p.total_tiles < h->sm_count ? p.total_tiles : h->sm_count;
  conv3x3_tc_kernel<TERMS><<<grid, 192, Cfg<TERMS>::kSmem, st>>>(xh, xl, wh, wl, p);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

}  // namespace

extern "C" int dsin_pack_conv3x3_w(dsin_handle_t h, const float* w_hwio, uint16_t* w_hi, uint16_t* w_lo,
                                   float* wscale, int cin, int cout, void* stream) {
  DSIN_REQUIRE(h, w_hwio && w_hi && w_lo && wscale, "null pointer");
  DSIN_REQUIRE(h, cin == 128 && cout == 128, "only 128 -> 128 channels are built");
  pack_w3x3_kernel<<<cout, 128, 0, (cudaStream_t)stream>>>(w_hwio, (__half*)w_hi, (__half*)w_lo, wscale, cin, cout);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

extern "C" int dsin_conv3x3_c128_tc(dsin_handle_t h, int n, int hh, int ww, const uint16_t* x_hi,
                                    const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo,
                                    const float* scale, const float* shift, int act, const uint16_t* res1_hi,
                                    const uint16_t* res1_lo, const uint16_t* res2_hi, const uint16_t* res2_lo,
                                    uint16_t* y_hi, uint16_t* y_lo, int terms, void* stream) {
  DSIN_REQUIRE(h, x_hi && w_hi && scale && shift && y_hi, "null pointer");
  DSIN_REQUIRE(h, terms == 1 || terms == 3, "terms must be 1 or 3");
  DSIN_REQUIRE(h, terms == 1 || (x_lo && w_lo && y_lo), "terms == 3 needs the lo planes");
  DSIN_REQUIRE(h, n > 0 && hh >= BH && ww >= BW, "image smaller than one 8x16 tile");
  DSIN_REQUIRE(h, act == DSIN_ACT_NONE || act == DSIN_ACT_RELU, "activation must be none or relu");
  CUtensorMap xh, xl, wh, wl;
  const uint64_t xd[4] = {128, (uint64_t)ww, (uint64_t)hh, (uint64_t)n};
  const uint64_t xs[3] = {256, (uint64_t)ww * 256, (uint64_t)hh * ww * 256};
  const uint32_t xb[4] = {64, BW, BH, 1};
  const uint64_t wd[2] = {128, 9 * 128};
  const uint64_t wsb[1] = {256};
  const uint32_t wb[2] = {64, 128};
  bool ok = encode_tmap(&xh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x_hi, xd, xs, xb, CU_TENSOR_MAP_SWIZZLE_128B) &&
            encode_tmap(&xl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x_lo ? x_lo : x_hi, xd, xs, xb,
                        CU_TENSOR_MAP_SWIZZLE_128B) &&
            encode_tmap(&wh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w_hi, wd, wsb, wb, CU_TENSOR_MAP_SWIZZLE_128B) &&
            encode_tmap(&wl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w_lo ? w_lo : w_hi, wd, wsb, wb,
                        CU_TENSOR_MAP_SWIZZLE_128B);
  if (!ok) return dsin_fail(h, DSIN_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", __func__);
  TcP p;
  p.scale = scale; p.shift = shift;
  p.r1h = (const __half*)res1_hi; p.r1l = (const __half*)res1_lo;
  p.r2h = (const __half*)res2_hi; p.r2l = (const __half*)res2_lo;
  p.yh = (__half*)y_hi; p.yl = (__half*)y_lo;
  p.n = n; p.H = hh; p.W = ww; p.act = act;
  p.tiles_w = (ww + BW - 1) / BW; p.tiles_h = (hh + BH - 1) / BH;
  p.total_tiles = n * p.tiles_w * p.tiles_h;
  cudaStream_t st = (cudaStream_t)stream;
  return terms == 3 ? launch_tc<3>(h, xh, xl, wh, wl, p, st) : launch_tc<1>(h, xh, xl, wh, wl, p, st);
}
