// K1, second generation: 3x3 stride-1 128->128 convolution with a HALO-RESIDENT activation tile.
//   same arithmetic and epilogue as conv_tc_kernel (src/autoencoder_imgcomp.py:229-234,257-262,275-288).
//
// conv_tc_kernel re-fetches the 128-pixel activation tile for each of the 9 taps (576 KB of A per tile in
// 3-term mode); the kernel was bound by TMA bytes in flight, not by the tensor pipe (fp16 1-term mode
// reached 39 % of peak, 3-term 75 %).  Here one TMA box per 64-channel chunk brings the tile WITH its
// 1-pixel halo -- (18 rows) x (10 cols) x 64 ch, 128-byte swizzled -- and the 9 taps are 9 UMMA
// descriptors into that single buffer: the pixel tile is 16 rows x 8 cols, so GEMM row r = (h = r/8,
// w = r%8) and every 8-row group of the operand is 8 consecutive halo pixels (8 x 128 B contiguous);
// consecutive groups are one halo row apart, a UNIFORM stride of 10 x 128 = 1280 B = the descriptor's SBO.
// Tap (ky,kx) just moves the start address by (10*ky + kx) x 128 B; because that is not 1024-aligned the
// descriptor's base_offset field carries (start >> 7) & 7 so the hardware's 128-B swizzle phase matches
// the one TMA used when writing.  A traffic drops 6.3x; only the weight slabs stream per tap.
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int TH = 16, TW = 8;                 // pixel tile (rows x cols)
constexpr int HALO_H = TH + 2, HALO_W = TW + 2;
constexpr int A_PLANE = 23 * 1024;             // 18*10*128 = 23040 B, rounded to 1 KB
constexpr int B_TILE = 128 * 128;              // [128 co][64 ci] fp16

struct HP {
  const float* scale;
  const float* shift;
  const __half *r1h, *r1l, *r2h, *r2l;
  __half *yh, *yl;
  int n, H, W, act;
  int tiles_w, tiles_h, total_tiles;
};

template <int TERMS>
struct HCfg {
  static constexpr int kASlot = (TERMS == 3 ? 2 : 1) * A_PLANE;
  static constexpr int kASlots = 2;
  static constexpr int kBStage = (TERMS == 3 ? 2 : 1) * B_TILE;
  static constexpr int kBStages = TERMS == 3 ? 4 : 8;
  static constexpr int kSmem = kASlots * kASlot + kBStages * kBStage + 1024 + 2048;
};

__device__ __forceinline__ uint64_t desc_sw128_off(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = make_smem_desc(saddr, 16, sbo_bytes, LAYOUT_SW128);
  d |= (uint64_t)((saddr >> 7) & 7u) << 49;  // base offset: swizzle phase of a non-1024-aligned start
  return d;
}

__device__ __forceinline__ void unpack8h(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ void add_res16(float* f, const __half* rh, const __half* rl, size_t off) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float a[8], b[8];
    unpack8h(__ldg(reinterpret_cast<const uint4*>(rh + off) + g), a);
    if (rl) {
      unpack8h(__ldg(reinterpret_cast<const uint4*>(rl + off) + g), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = __fadd_rn(a[e], b[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) f[g * 8 + e] = __fadd_rn(f[g * 8 + e], a[e]);
  }
}

template <int TERMS>
__global__ void __launch_bounds__(192, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                    const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
                    const __grid_constant__ HP p) {
  using C = HCfg<TERMS>;
  constexpr int SB = C::kBStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* a_buf = smem;
  uint8_t* b_buf = smem + C::kASlots * C::kASlot;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_buf + SB * C::kBStage);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + SB;
  uint64_t* tfull = b_empty + SB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_scale = reinterpret_cast<float*>(tmem_ptr + 2);
  float* s_shift = s_scale + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    for (int i = 0; i < SB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
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
  constexpr uint32_t kABytes = (TERMS == 3 ? 2u : 1u) * (uint32_t)(HALO_H * HALO_W * 128);
  constexpr uint32_t kBBytes = (TERMS == 3 ? 2u : 1u) * (uint32_t)B_TILE;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int aslot = 0, bstage = 0;
      uint32_t aphase = 0, bphase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int tw = tile % p.tiles_w, t2 = tile / p.tiles_w;
        const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
        const int ox0 = tw * TW - 1, oy0 = th * TH - 1;
        for (int cc = 0; cc < 2; ++cc) {
          mbar_wait(&a_empty[aslot], aphase ^ 1u);
          uint8_t* as = a_buf + aslot * C::kASlot;
          mbar_expect_tx(&a_full[aslot], kABytes);
          tma_load_4d(as, &tm_xh, &a_full[aslot], cc * 64, ox0, oy0, n);
          if (TERMS == 3) tma_load_4d(as + A_PLANE, &tm_xl, &a_full[aslot], cc * 64, ox0, oy0, n);
          if (++aslot == 2) {
            aslot = 0;
            aphase ^= 1u;
          }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bstage], bphase ^ 1u);
            uint8_t* bs = b_buf + bstage * C::kBStage;
            mbar_expect_tx(&b_full[bstage], kBBytes);
            tma_load_2d(bs, &tm_wh, &b_full[bstage], cc * 64, tap * 128);
            if (TERMS == 3) tma_load_2d(bs + B_TILE, &tm_wl, &b_full[bstage], cc * 64, tap * 128);
            if (++bstage == SB) {
              bstage = 0;
              bphase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, 128, 0);
      int aslot = 0, bstage = 0, it = 0;
      uint32_t aphase = 0, bphase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 128u;
        for (int cc = 0; cc < 2; ++cc) {
          mbar_wait(&a_full[aslot], aphase);
          fence_after_sync();
          const uint32_t sa = smem_u32(a_buf + aslot * C::kASlot);
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_full[bstage], bphase);
            fence_after_sync();
            const int ky = tap / 3, kx = tap - 3 * ky;
            const uint32_t a_off = (uint32_t)(ky * HALO_W + kx) * 128u;
            const uint32_t sb = smem_u32(b_buf + bstage * C::kBStage);
            const uint64_t b_hi = make_smem_desc(sb, 16, 1024, LAYOUT_SW128);
            const uint64_t b_lo = make_smem_desc(sb + B_TILE, 16, 1024, LAYOUT_SW128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a_hi = desc_sw128_off(sa + a_off + k * 32, HALO_W * 128);
              umma_f16(d_tmem, a_hi, b_hi + 2 * k, idesc, (cc | tap | k) ? 1u : 0u);
              if (TERMS == 3) {
                const uint64_t a_lo = desc_sw128_off(sa + A_PLANE + a_off + k * 32, HALO_W * 128);
                umma_f16(d_tmem, a_hi, b_lo + 2 * k, idesc, 1u);
                umma_f16(d_tmem, a_lo, b_hi + 2 * k, idesc, 1u);
              }
            }
            umma_commit(&b_empty[bstage]);
            if (++bstage == SB) {
              bstage = 0;
              bphase ^= 1u;
            }
          }
          umma_commit(&a_empty[aslot]);
          if (++aslot == 2) {
            aslot = 0;
            aphase ^= 1u;
          }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..5
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int hl = row >> 3, wl = row & 7;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int tw = tile % p.tiles_w, t2 = tile / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      const int oy = th * TH + hl, ox = tw * TW + wl;
      const bool valid = oy < p.H && ox < p.W;
      const size_t pix = ((size_t)n * p.H + oy) * p.W + ox;
      mbar_wait(&tfull[acc], (uint32_t)(it >> 1) & 1u);
      fence_after_sync();
#pragma unroll 1
      for (int chunk = 0; chunk < 8; ++chunk) {
        const int c0 = chunk * 16;
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128 + c0), v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t = __fadd_rn(__fmul_rn(__uint_as_float(v[j]), s_scale[c0 + j]), s_shift[c0 + j]);
            f[j] = p.act == DSIN_ACT_RELU ? fmaxf(t, 0.f) : t;
          }
          const size_t off = pix * 128 + c0;
          if (p.r1h) add_res16(f, p.r1h, p.r1l, off);
          if (p.r2h) add_res16(f, p.r2h, p.r2l, off);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
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

template <int TERMS>
int launch_halo(dsin_handle_t h, const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& wh,
                const CUtensorMap& wl, const HP& p, cudaStream_t st) {
  using C = HCfg<TERMS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv3x3_halo_kernel<TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem) !=
        cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: cannot raise dynamic shared memory", __func__);
    configured = true;
  }
  int grid = p.total_tiles < h->sm_count ? p.total_tiles : h->sm_count;
  conv3x3_halo_kernel<TERMS><<<grid, 192, C::kSmem, st>>>(xh, xl, wh, wl, p);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

}  // namespace

// Internal entry (declared in conv_tc.cuh): same contract as dsin_conv3x3_c128_tc.
int conv3x3_c128_halo(dsin_handle_t h, int n, int hh, int ww, const uint16_t* x_hi, const uint16_t* x_lo,
                      const uint16_t* w_hi, const uint16_t* w_lo, const float* scale, const float* shift, int act,
                      const uint16_t* r1h, const uint16_t* r1l, const uint16_t* r2h, const uint16_t* r2l,
                      uint16_t* y_hi, uint16_t* y_lo, int terms, cudaStream_t st) {
  CUtensorMap xh, xl, wh, wl;
  const uint64_t xd[4] = {128, (uint64_t)ww, (uint64_t)hh, (uint64_t)n};
  const uint64_t xs[3] = {256, (uint64_t)ww * 256, (uint64_t)hh * ww * 256};
  const uint32_t xb[4] = {64, HALO_W, HALO_H, 1};
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
  HP p;
  p.scale = scale; p.shift = shift;
  p.r1h = (const __half*)r1h; p.r1l = (const __half*)r1l; p.r2h = (const __half*)r2h; p.r2l = (const __half*)r2l;
  p.yh = (__half*)y_hi; p.yl = (__half*)y_lo;
  p.n = n; p.H = hh; p.W = ww; p.act = act;
  p.tiles_w = (ww + TW - 1) / TW; p.tiles_h = (hh + TH - 1) / TH;
  p.total_tiles = n * p.tiles_w * p.tiles_h;
  return terms == 3 ? launch_halo<3>(h, xh, xl, wh, wl, p, st) : launch_halo<1>(h, xh, xl, wh, wl, p, st);
}
