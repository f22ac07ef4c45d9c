// K8 / K4 on a HALO tile: the 32-channel 3x3 layers of the SI-Net with dilation 1, 2, 4 (src/siNet.py:31-34,39) and the
// (2,3,3) masked 3-D convolutions of the probability model (src/probclass_imgcomp.py:185-196,214-261) as tcgen05
// implicit GEMMs whose activation tile is loaded ONCE with its halo and whose filter stays in shared memory.
//
// These layers are small-N GEMMs (N = 32 or 16 output channels, K = 32 per tap): per 128-pixel tile the MMAs take
// ~1-2 k cycles, while the tap-streaming kernel (conv_tc.cu) moved 9-18 activation tiles per output tile from L2 --
// measured L2-bound (profiles/r2_v1_ncu_small_kernels.txt: SI-Net layer 0.33-0.39 ms at batch 8 with L2 at 56-65 %
// and DRAM at 28 %; probclass layers 0.35-0.47 ms).  Here
//   * the (16 + span_y) x (8 + span_x) [x 2 depth slices] input pixels a 16 x 8 output tile reads arrive as ONE TMA box
//     per plane, 64-byte rows, 64-byte swizzle; every tap is a shared-memory descriptor into that tile (start address
//     + ((dz * HH + dy) * HW + dx) * 64 B, 8-pixel row groups HW * 64 B apart) -- as in conv_ws.cu, TMA and the UMMA
//     descriptor agree on the swizzle phase because both take it from the absolute shared-memory address;
//   * the whole filter (<= 18 taps x 32 x 32 split fp16 = 72 KB) is loaded once per CTA;
//   * the hi*hi products and the 2^-11-times-smaller hi*lo + lo*hi products have separate TMEM accumulators
//     (round-toward-zero accumulation, see conv_h3.cu) that sit side by side, and a tap's hi and lo weight slabs sit
//     side by side in shared memory: ONE MMA of N = 2 x NPAD computes a_hi x [w_hi ; w_lo] into both accumulators,
//     a second of N = NPAD adds a_lo x w_hi -- two MMAs instead of three per k-step (with N this small an MMA costs
//     its 4 KB activation-operand read, not its arithmetic); the epilogue adds the two accumulators in fp32;
//   * split-fp16 results leave through per-warp swizzled staging blocks and TMA stores (which also clip partial tiles).
#include "tc_common.cuh"
#include "conv_tc.cuh"

using namespace tc;

namespace {

constexpr int TR = 16, TC = 8;  // output tile: 16 rows x 8 pixels = 128 GEMM rows, m = r * 8 + c
constexpr int NTHREADS = 320;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue

template <int NPAD, int TERMS>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_h32_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
                const __grid_constant__ CUtensorMap tm_yh, const __grid_constant__ CUtensorMap tm_yl,
                const __grid_constant__ ConvH32Args p) {
  constexpr int PLANES = TERMS == 3 ? 2 : 1;
  constexpr int W_SLAB = PLANES * NPAD * 64;        // one tap: [hi: NPAD couts x 32 cin fp16][lo: same], 64-byte rows
  constexpr int ACC = (TERMS == 3 ? 2 : 1) * NPAD;  // accumulator columns per buffer (large terms, small terms)
  constexpr int TMEM_COLS = 2 * ACC <= 32 ? 32 : (2 * ACC <= 64 ? 64 : 128);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* wsm = smem;                                    // [plane][tap][NPAD rows][64 B]
  uint8_t* a_tiles = wsm + p.w_region;                    // na x [plane][halo box]
  uint8_t* stg = a_tiles + p.na * PLANES * p.a_plane;     // 2 buffers x 4 quarters x [hi 2 KB, lo 2 KB]
  uint64_t* full_a = reinterpret_cast<uint64_t*>(stg + 8 * 4096);
  uint64_t* empty_a = full_a + 4;
  uint64_t* tfull = empty_a + 4;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wfull + 1);
  float* s_scale = reinterpret_cast<float*>(tmem_ptr + 2);
  float* s_shift = s_scale + 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full_a[i], 1);
      mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    mbar_init(wfull, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_xh);
    prefetch_tmap(&tm_wh);
    if (TERMS == 3) {
      prefetch_tmap(&tm_xl);
      prefetch_tmap(&tm_wl);
    }
    if (p.yh) prefetch_tmap(&tm_yh);
  }
  if (threadIdx.x < 32) {
    s_scale[threadIdx.x] = (int)threadIdx.x < p.cout ? p.scale[threadIdx.x] : 0.f;
    s_shift[threadIdx.x] = (int)threadIdx.x < p.cout ? p.shift[threadIdx.x] : 0.f;
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t a_bytes = (uint32_t)(p.hd * p.hh * p.hw * 64);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: the filter once, then halo boxes
    if (elect_one()) {
      mbar_expect_tx(wfull, (uint32_t)(p.ntaps * W_SLAB));
      for (int t = 0; t < p.ntaps; ++t) {
        tma_load_2d(wsm + t * W_SLAB, &tm_wh, wfull, 0, p.tw[t] * NPAD);
        if (TERMS == 3) tma_load_2d(wsm + t * W_SLAB + NPAD * 64, &tm_wl, wfull, 0, p.tw[t] * NPAD);
      }
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w, t2 = tile / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      const int n_in = p.dout ? (n / p.dout) * p.din + (n % p.dout) : n;
      mbar_wait(&empty_a[stage], phase ^ 1u);
      if (elect_one()) {
        uint8_t* st = a_tiles + stage * PLANES * p.a_plane;
        mbar_expect_tx(&full_a[stage], PLANES * a_bytes);
        tma_load_4d(st, &tm_xh, &full_a[stage], 0, tw * TC + p.ox, th * TR + p.oy, n_in);
        if (TERMS == 3) tma_load_4d(st + p.a_plane, &tm_xl, &full_a[stage], 0, tw * TC + p.ox, th * TR + p.oy, n_in);
      }
      __syncwarp();
      if (++stage == p.na) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(128, NPAD, 0);
    constexpr uint32_t idesc2 = make_idesc_f16(128, 2 * NPAD, 0);  // a_hi x [w_hi ; w_lo] -> [large | small] accumulators
    const uint32_t w_base = smem_u32(wsm);
    mbar_wait(wfull, 0);
    fence_after_sync();
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
      fence_after_sync();
      const uint32_t d_main = tmem_base + (uint32_t)(acc * ACC);
      const uint32_t d_lo = d_main + NPAD;
      mbar_wait(&full_a[stage], phase);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t a_base = smem_u32(a_tiles + stage * PLANES * p.a_plane);
        const uint32_t sbo = (uint32_t)(p.hw * 64);
        for (int t = 0; t < p.ntaps; ++t) {
          const uint32_t a_off = (uint32_t)(((p.tz[t] * p.hh + p.ty[t]) * p.hw + p.tx[t]) * 64);
          const uint64_t a_hi = make_smem_desc(a_base + a_off, 16, sbo, LAYOUT_SW64);
          const uint64_t a_lo = make_smem_desc(a_base + p.a_plane + a_off, 16, sbo, LAYOUT_SW64);
          const uint64_t b_hi = make_smem_desc(w_base + (uint32_t)(t * W_SLAB), 16, 512, LAYOUT_SW64);  // hi rows, then lo
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t first = (t | k) ? 1u : 0u;
            if (TERMS == 3) {
              umma_f16(d_main, a_hi + 2 * k, b_hi + 2 * k, idesc2, first);  // columns [0, NPAD): hi*hi, [NPAD, 2 NPAD): hi*lo
              umma_f16(d_lo, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);        // lo*hi into the small-term accumulator
            } else {
              umma_f16(d_main, a_hi + 2 * k, b_hi + 2 * k, idesc, first);
            }
          }
        }
        umma_commit(&empty_a[stage]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++stage == p.na) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..9: two per TMEM lane quarter
    // Warp (q, half) turns 16 of the quarter's NPAD columns (NPAD = 16: only half 0 has columns) into output; the two
    // warps of a quarter share one staging block (pixel p at p * 64 B) and meet at named barrier 1 + q; half 0 issues
    // the quarter's TMA stores.  (Measured with 4 warps x NPAD columns: the kernel was bound by this epilogue, not by
    // its MMAs or its memory traffic -- profiles/r2_v4_ncu_conv_h32.txt.)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool has_cols = half * 16 < NPAD;
    const int row = q * 32 + lane;
    const int rl = row >> 3, cl = row & 7;
    const uint32_t stg_q = smem_u32(stg + q * 4096);  // + (it & 1) * 16 KB: tiles alternate between two staging buffers
    const uint32_t row_off = (uint32_t)lane * 64u;
    const int sw = (lane >> 1) & 3;  // 64-byte swizzle: 16-byte piece j of pixel p sits at ((j ^ ((p >> 1) & 3)) * 16)
    const bool split_out = p.yh != nullptr;
    const int c0 = half * 16;
    // activation without per-element branches: t = max(max(t, t * slope), floor)
    const float slope = p.act == DSIN_ACT_LRELU02 ? 0.2f : 1.f;
    const float floor_v = p.act == DSIN_ACT_RELU ? 0.f : -INFINITY;
    float sc[16], sh[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      sc[j] = has_cols ? s_scale[c0 + j] : 0.f;
      sh[j] = has_cols ? s_shift[c0 + j] : 0.f;
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int tw = tile % p.tiles_w, t2 = tile / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      const int oy = th * TR + rl, ox = tw * TC + cl;
      const bool valid = oy < p.OH && ox < p.OW;
      const size_t pix = ((size_t)n * p.OH + oy) * p.OW + ox;
      const uint32_t stg_hi = stg_q + (uint32_t)(acc * 4 * 4096);
      const uint32_t stg_lo = stg_hi + 2048;
      if (split_out) {
        // the stores of the tile before the previous one (issued by half 0) have finished reading this staging buffer,
        // the previous tile's may still be in flight
        if (half == 0 && lane == 0) tma_store_wait_read1();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      }
      mbar_wait(&tfull[acc], (uint32_t)(it >> 1) & 1u);
      fence_after_sync();
      if (has_cols) {
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC + c0);
        uint32_t v[16], vl[16];
        tmem_ld_32x16(lane_base, v);
        if (TERMS == 3) tmem_ld_32x16(lane_base + (uint32_t)NPAD, vl);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(v[j]);
          if (TERMS == 3) a = __fadd_rn(a, __uint_as_float(vl[j]));  // large + small product terms, round to nearest
          const float t = __fadd_rn(__fmul_rn(a, sc[j]), sh[j]);
          f[j] = fmaxf(fmaxf(t, __fmul_rn(t, slope)), floor_v);
        }
        if (p.r1f && valid) {  // cropped fp32 skip of the probability model's residual block
          const size_t rpix = (((size_t)(n / p.dout) * p.r1_d + (n % p.dout) + p.r1_dz) * p.r1_oh + oy + p.r1_dy) *
                                  p.r1_ow + ox + p.r1_dx;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < p.r1_c) f[j] = __fadd_rn(f[j], __ldg(p.r1f + rpix * p.r1_c + c0 + j));
        }
        if (!split_out) {
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.cout) p.yf[pix * p.cout + c0 + j] = f[j];
          }
        } else {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 uh, ul;
            __half2* hh2 = reinterpret_cast<__half2*>(&uh);
            __half2* ll2 = reinterpret_cast<__half2*>(&ul);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x0 = f[g * 8 + 2 * e], x1 = f[g * 8 + 2 * e + 1];
              const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
              hh2[e] = __halves2half2(h0, h1);
              ll2[e] = __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
            }
            const uint32_t so = row_off + (uint32_t)(((2 * half + g) ^ sw) << 4);
            sts16(stg_hi + so, uh);
            if (p.yl) sts16(stg_lo + so, ul);
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (split_out) {
        fence_proxy_async();  // this thread's generic-proxy writes to the staging block -> visible to the TMA stores
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        if (half == 0 && lane == 0) {  // the stores clip rows / pixels past the output
          tma_store_4d_issue(&tm_yh, stg_hi, 0, tw * TC, th * TR + q * 4, n);
          if (p.yl) tma_store_4d_issue(&tm_yl, stg_lo, 0, tw * TC, th * TR + q * 4, n);
          tma_store_commit();  // one group per tile
        }
      }
    }
    if (split_out) {
      if (half == 0 && lane == 0) tma_store_wait_all();
      __syncwarp();
    }
  }

  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int NPAD, int TERMS>
int launch_h32(dsin_handle_t h, const CUtensorMap* m, const ConvH32Args& p, int smem, cudaStream_t st) {
  static bool configured[DSIN_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!configured[h->device]) {
    if (cudaFuncSetAttribute(conv_h32_kernel<NPAD, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: cannot raise dynamic shared memory", __func__);
    configured[h->device] = true;
  }
  const int grid = p.total_tiles < h->sm_count ? p.total_tiles : h->sm_count;
  conv_h32_kernel<NPAD, TERMS><<<grid, NTHREADS, smem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], p);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

}  // namespace

// x: (32 ch, W, H, ND) channels-last split fp16; w packed [wtaps][npad][32] (dsin_pack_conv_w_tc); see ConvH32Args.
int conv_h32_launch(dsin_handle_t h, const __half* x_hi, const __half* x_lo, const __half* w_hi, const __half* w_lo,
                    int W, int H, int ND, int wtaps, const ConvH32Args& a, cudaStream_t st) {
  const int npad = (a.cout + 15) / 16 * 16;
  if (npad != 16 && npad != 32) return dsin_fail(h, DSIN_ERR_UNSUPPORTED, "%s: cout must be <= 32", __func__);
  if (a.ntaps < 1 || a.ntaps > 18) return dsin_fail(h, DSIN_ERR_UNSUPPORTED, "%s: 1..18 taps", __func__);
  const int planes = a.terms == 3 ? 2 : 1;
  ConvH32Args p = a;
  p.tiles_w = (a.OW + TC - 1) / TC;
  p.tiles_h = (a.OH + TR - 1) / TR;
  p.total_tiles = a.n_out * p.tiles_w * p.tiles_h;
  p.w_plane = 0;
  p.w_region = (planes * a.ntaps * npad * 64 + 1023) / 1024 * 1024;
  p.a_plane = (a.hd * a.hh * a.hw * 64 + 1023) / 1024 * 1024;
  const int fixed = p.w_region + 8 * 4096 + 1024 /*barriers, scale/shift*/ + 1024 /*alignment*/;
  int na = (227 * 1024 - fixed) / (planes * p.a_plane);
  if (na > 4) na = 4;
  if (na < 2) return dsin_fail(h, DSIN_ERR_UNSUPPORTED, "%s: halo tile too large for two pipeline stages", __func__);
  p.na = na;
  const int smem = fixed + na * planes * p.a_plane;
  CUtensorMap m[6];
  const CUtensorMapDataType f16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_64B;
  const uint64_t xd[4] = {32, (uint64_t)W, (uint64_t)H, (uint64_t)ND};
  const uint64_t xs[3] = {64, (uint64_t)W * 64, (uint64_t)H * W * 64};
  const uint32_t xb[4] = {32, (uint32_t)a.hw, (uint32_t)a.hh, (uint32_t)a.hd};
  const uint64_t wd[2] = {32, (uint64_t)wtaps * npad};
  const uint64_t wsb[1] = {64};
  const uint32_t wb[2] = {32, (uint32_t)npad};
  const uint64_t yd[4] = {32, (uint64_t)a.OW, (uint64_t)a.OH, (uint64_t)a.n_out};
  const uint64_t ys[3] = {64, (uint64_t)a.OW * 64, (uint64_t)a.OH * a.OW * 64};
  const uint32_t yb[4] = {32, TC, 4, 1};
  bool ok = encode_tmap(&m[0], f16, 4, x_hi, xd, xs, xb, sw) &&
            encode_tmap(&m[1], f16, 4, x_lo ? x_lo : x_hi, xd, xs, xb, sw) &&
            encode_tmap(&m[2], f16, 2, w_hi, wd, wsb, wb, sw) &&
            encode_tmap(&m[3], f16, 2, w_lo ? w_lo : w_hi, wd, wsb, wb, sw);
  if (a.yh) {
    if (a.cout != 32) return dsin_fail(h, DSIN_ERR_UNSUPPORTED, "%s: split output needs 32 channels", __func__);
    ok = ok && encode_tmap(&m[4], f16, 4, a.yh, yd, ys, yb, sw) &&
         encode_tmap(&m[5], f16, 4, a.yl ? a.yl : a.yh, yd, ys, yb, sw);
  } else {
    m[4] = m[0];
    m[5] = m[0];
  }
  if (!ok) return dsin_fail(h, DSIN_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", __func__);
  if (npad == 32) return a.terms == 3 ? launch_h32<32, 3>(h, m, p, smem, st) : launch_h32<32, 1>(h, m, p, smem, st);
  return a.terms == 3 ? launch_h32<16, 3>(h, m, p, smem, st) : launch_h32<16, 1>(h, m, p, smem, st);
}
