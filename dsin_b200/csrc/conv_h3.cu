// K1, fp32-class form ("terms = 3"): the 3x3 128->128 trunk convolution on CTA pairs with a HALO-RESIDENT
// split-fp16 activation tile, streamed split-fp16 weights, SEPARATE accumulators for the large and the small
// product terms, and a coalesced staged epilogue.  Same arithmetic as conv_tc2_kernel<3>
// (src/autoencoder_imgcomp.py:229-234,257-262,275-288): y = act(conv(x) * scale + shift) + r1 + r2, every tensor a
// pair of NHWC fp16 planes (value = hi + lo), products hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM.
//
// What changed against conv_tc2.cu, and why:
//  * The activation tile (hi and lo plane) is loaded ONCE per 64-channel chunk with its halo -- 16 x 8 output pixels
//    read 18 x 10 input pixels -- and the nine taps are nine shared-memory descriptors into it (start address moved by
//    (ky * 10 + kx) * 128 B, 8-pixel row groups 1280 B apart; TMA and the UMMA descriptor both take the 128-byte
//    swizzle phase from the absolute shared-memory address, see conv_ws.cu).  Only the weight slabs (16 KB per tap
//    and chunk and CTA) still stream.  L2 -> shared-memory bytes per tile: 380 KB instead of 864 KB -- under the
//    1000 W cap bytes moved are clock.
//  * tcgen05 accumulates with round-toward-zero: every MMA that adds into an accumulator of magnitude |acc| loses up
//    to one ulp(|acc|), biased toward zero.  With all three product terms in one accumulator that is 216 roundings per
//    output (measured: z of the encoder 2.8e-5 rms off the float64 oracle, 28x the fp32 CPU oracle's own error, bias
//    -7.6e-6 relative).  The hi*hi products (72 MMAs) now have an accumulator of their own; the 2^-11-times-smaller
//    hi*lo and lo*hi products (144 MMAs) go into a second one whose roundings are ~2^-11 of an ulp of the result; the
//    epilogue adds the two in fp32 (round to nearest).
//  * Residual tensors no longer ride the tensor pipe (identity MMAs, a third accumulator): every epilogue warp brings
//    its 32 pixels x 64 channels of each residual plane in with cp.async (4 pixels x 128 contiguous bytes per
//    instruction) while the tile's MMAs are still running.  Results leave through per-warp swizzled staging blocks
//    and TMA stores (one per plane, warp and tile; partial tiles are clipped by the store).
#include "tc_common.cuh"
#include "conv_tc.cuh"

using namespace tc;

namespace {

constexpr int TR = 16, TC = 8;               // output tile: 16 rows x 8 pixels = 128 GEMM rows, m = r * 8 + c
constexpr int HALO_R = TR + 2, PITCH = TC + 2;
constexpr int A_BYTES = HALO_R * PITCH * 128;             // 23 040 B: one plane of one 64-channel chunk of the halo tile
constexpr int A_PLANE = (A_BYTES + 1023) / 1024 * 1024;   // planes start 1024-byte aligned
constexpr int A_STAGE = 2 * A_PLANE;                      // hi + lo
constexpr int NA = 2;
constexpr int B_PLANE = 64 * 128;                         // 64 couts x 64 cin fp16
constexpr int B_STAGE = 2 * B_PLANE;                      // hi + lo slab of one (tap, chunk)
constexpr int NB = 4;
constexpr int STG_PLANE = 32 * 128;                       // one epilogue warp: 32 pixels x 64 channels fp16
constexpr int STG_WARP = 2 * STG_PLANE;                   // hi + lo
constexpr int SMEM_BYTES = NA * A_STAGE + NB * B_STAGE + 8 * STG_WARP + 2048 + 1024;
constexpr int NTHREADS = 352;  // warp 0: weight producer, 1: MMA issuer, 2..9: epilogue, 10: activation producer

__device__ __forceinline__ void add_split8(float* f, const uint4& h, const uint4& l, bool has_lo) {
  const __half2* hh = reinterpret_cast<const __half2*>(&h);
  const __half2* ll = reinterpret_cast<const __half2*>(&l);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 a = __half22float2(hh[i]);
    if (has_lo) {
      const float2 b = __half22float2(ll[i]);
      a.x = __fadd_rn(a.x, b.x);  // exact: hi + lo carries at most 22 significant bits
      a.y = __fadd_rn(a.y, b.y);
    }
    f[2 * i] = __fadd_rn(f[2 * i], a.x);
    f[2 * i + 1] = __fadd_rn(f[2 * i + 1], a.y);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
conv_h3_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
               const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
               const __grid_constant__ CUtensorMap tm_yh, const __grid_constant__ CUtensorMap tm_yl,
               const __grid_constant__ ConvH3Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* a_tiles = smem;                       // NA x (hi plane, lo plane) halo chunks
  uint8_t* b_tiles = a_tiles + NA * A_STAGE;     // NB x (hi slab, lo slab)
  uint8_t* stg = b_tiles + NB * B_STAGE;         // 8 x (hi block, lo block)
  uint64_t* full_a = reinterpret_cast<uint64_t*>(stg + 8 * STG_WARP);
  uint64_t* empty_a = full_a + NA;
  uint64_t* full_b = empty_a + NA;
  uint64_t* empty_b = full_b + NB;
  uint64_t* tfull = empty_b + NB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_scale = reinterpret_cast<float*>(tmem_ptr + 2);
  float* s_shift = s_scale + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) {
      mbar_init(&full_a[i], 1);
      mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(&full_b[i], 1);
      mbar_init(&empty_b[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);  // 8 epilogue warps x 2 CTAs (the leader's copy is the one used)
    }
    fence_barrier_init();
    prefetch_tmap(&tm_xh);
    prefetch_tmap(&tm_xl);
    prefetch_tmap(&tm_wh);
    prefetch_tmap(&tm_wl);
    prefetch_tmap(&tm_yh);
    prefetch_tmap(&tm_yl);
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  if (warp == 1) tmem_alloc2(tmem_ptr, 512);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  // everything above is independent of the previous layer; its output (our input, our residuals) is read from here on
  pdl_wait();
  pdl_launch_dependents();
  const int pairs = (p.total_tiles + 1) / 2;
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 10) {
    // ------------------------------------------------------------ activation producer (both CTAs): halo chunks
    int stage = 0;
    uint32_t phase = 0;
    for (int pi = cid; pi < pairs; pi += nclusters) {
      int tile = 2 * pi + (int)rank;
      if (tile >= p.total_tiles) tile = p.total_tiles - 1;  // odd tail: recompute a valid tile, never stored
      const int tw = tile % p.tiles_w, t2 = tile / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      int nt = 2 * (pi + nclusters) + (int)rank;  // this CTA's next tile: pull its boxes into L2 now
      const bool has_next = nt < p.total_tiles;
      if (!has_next) nt = tile;
      const int ntw = nt % p.tiles_w, nt2 = nt / p.tiles_w;
      const int nth = nt2 % p.tiles_h, nn = nt2 / p.tiles_h;
      for (int cc = 0; cc < 2; ++cc) {
        mbar_wait(&empty_a[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* st = a_tiles + stage * A_STAGE;
          if (leader) mbar_expect_tx(&full_a[stage], 4u * (uint32_t)A_BYTES);  // 2 planes x 2 CTAs
          tma2_load_4d(st, &tm_xh, &full_a[stage], cc * 64, tw * TC - 1, th * TR - 1, n);
          tma2_load_4d(st + A_PLANE, &tm_xl, &full_a[stage], cc * 64, tw * TC - 1, th * TR - 1, n);
          if (has_next) {
            tma_prefetch_4d(&tm_xh, cc * 64, ntw * TC - 1, nth * TR - 1, nn);
            tma_prefetch_4d(&tm_xl, cc * 64, ntw * TC - 1, nth * TR - 1, nn);
          }
        }
        __syncwarp();
        if (++stage == NA) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------ weight producer (both CTAs): this CTA's 64 couts
    int stage = 0;
    uint32_t phase = 0;
    for (int pi = cid; pi < pairs; pi += nclusters)
      for (int kb = 0; kb < 18; ++kb) {  // kb = chunk * 9 + tap
        const int cc = kb / 9, tap = kb - cc * 9;
        mbar_wait(&empty_b[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* st = b_tiles + stage * B_STAGE;
          if (leader) mbar_expect_tx(&full_b[stage], 2u * (uint32_t)B_STAGE);
          tma2_load_2d(st, &tm_wh, &full_b[stage], cc * 64, tap * 128 + (int)rank * 64);
          tma2_load_2d(st + B_PLANE, &tm_wl, &full_b[stage], cc * 64, tap * 128 + (int)rank * 64);
        }
        __syncwarp();
        if (++stage == NB) {
          stage = 0;
          phase ^= 1u;
        }
      }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(256, 128, 0);
      int sa = 0, sb = 0, it = 0;
      uint32_t pa = 0, pb = 0;
      for (int pi = cid; pi < pairs; pi += nclusters, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        fence_after_sync();
        const uint32_t d_main = tmem_base + (uint32_t)acc * 256u;  // hi*hi products
        const uint32_t d_lo = d_main + 128u;                       // hi*lo + lo*hi products
        for (int cc = 0; cc < 2; ++cc) {
          mbar_wait(&full_a[sa], pa);
          fence_after_sync();
          const uint32_t a_base = smem_u32(a_tiles + sa * A_STAGE);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&full_b[sb], pb);
            fence_after_sync();
            if (elect_one()) {
              const int ky = tap / 3, kx = tap - ky * 3;
              const uint32_t a_off = (uint32_t)((ky * PITCH + kx) * 128);
              const uint32_t b_base = smem_u32(b_tiles + sb * B_STAGE);
              const uint64_t a_hi = make_smem_desc(a_base + a_off, 16, PITCH * 128, LAYOUT_SW128);
              const uint64_t a_lo = make_smem_desc(a_base + A_PLANE + a_off, 16, PITCH * 128, LAYOUT_SW128);
              const uint64_t b_hi = make_smem_desc(b_base, 16, 1024, LAYOUT_SW128);
              const uint64_t b_lo = make_smem_desc(b_base + B_PLANE, 16, 1024, LAYOUT_SW128);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t first = (cc | tap | k) ? 1u : 0u;
                umma2_f16(d_main, a_hi + 2 * k, b_hi + 2 * k, idesc, first);
                umma2_f16(d_lo, a_hi + 2 * k, b_lo + 2 * k, idesc, first);
                umma2_f16(d_lo, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
              }
              umma2_commit(&empty_b[sb]);  // frees the weight stage in BOTH CTAs
            }
            __syncwarp();
            if (++sb == NB) {
              sb = 0;
              pb ^= 1u;
            }
          }
          if (elect_one()) umma2_commit(&empty_a[sa]);  // the halo chunk is consumed
          __syncwarp();
          if (++sa == NA) {
            sa = 0;
            pa ^= 1u;
          }
        }
        if (elect_one()) umma2_commit(&tfull[acc]);  // accumulators complete in both CTAs
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..9 (both CTAs, own TMEM)
    // Two warps per TMEM lane quarter, 64 output channels each; thread = pixel m = q * 32 + lane; the warp's 32
    // pixels are rows 4q..4q+3 of the tile.  Staging blocks (hi, lo): pixel p at p * 128 B, 16-byte piece j at
    // ((j ^ (p & 7)) * 16) = the 128-byte-swizzle layout of a (64 ch, 8 px, 4 rows) TMA box.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t stg_hi = smem_u32(stg + (warp - 2) * STG_WARP);
    const uint32_t stg_lo = stg_hi + STG_PLANE;
    const uint32_t row_off = (uint32_t)lane * 128u;
    const int sw = lane & 7;
    int it = 0;
    for (int pi = cid; pi < pairs; pi += nclusters, ++it) {
      const int acc = it & 1;
      const int tile = 2 * pi + (int)rank;
      const bool tvalid = tile < p.total_tiles;
      const int tcl = tvalid ? tile : p.total_tiles - 1;
      const int tw = tcl % p.tiles_w, t2 = tcl / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      const int oy0 = th * TR + q * 4, ox0 = tw * TC;
      if (lane == 0) tma_store_wait_read();  // the previous tile's stores have finished READING the blocks
      __syncwarp();
      // sum of the residual tensors for this thread's pixel and 64 channels (fp32), fetched during the tile's MMAs
      float rs[64];
      const bool any_res = p.r1h != nullptr || p.r2h != nullptr;
      if (any_res) {
#pragma unroll
        for (int j = 0; j < 64; ++j) rs[j] = 0.f;
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const __half* rh = rr == 0 ? p.r1h : p.r2h;
        const __half* rl = rr == 0 ? p.r1l : p.r2l;
        if (rh == nullptr) continue;  // warp-uniform
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pp = 4 * i + (lane >> 3), j = lane & 7;  // pixel of the block, 16-byte piece
          const int oy = oy0 + (pp >> 3), ox = ox0 + (pp & 7);
          const bool ok = tvalid && oy < p.OH && ox < p.OW;
          const size_t goff = ok ? ((((size_t)n * p.OH + oy) * p.OW + ox) * 128 + half * 64 + j * 8) : 0;
          const uint32_t so = (uint32_t)(pp * 128 + ((j ^ (pp & 7)) << 4));
          cp_async16(stg_hi + so, rh + goff, ok);
          if (rl) cp_async16(stg_lo + so, rl + goff, ok);
        }
        cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t so = row_off + (uint32_t)((j ^ sw) << 4);
          const uint4 vh = lds16(stg_hi + so);
          uint4 vl = make_uint4(0u, 0u, 0u, 0u);
          if (rl) vl = lds16(stg_lo + so);
          add_split8(rs + 8 * j, vh, vl, rl != nullptr);
        }
        __syncwarp();  // everyone has read before the blocks are overwritten
      }
      mbar_wait(&tfull[acc], (uint32_t)(it >> 1) & 1u);
      fence_after_sync();
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + half * 64);
#pragma unroll
      for (int chunk = 0; chunk < 4; ++chunk) {
        const int c0 = half * 64 + chunk * 16;
        uint32_t vm[16], vl[16];
        tmem_ld_32x16(lane_base + (uint32_t)(chunk * 16), vm);
        tmem_ld_32x16(lane_base + 128u + (uint32_t)(chunk * 16), vl);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = __fadd_rn(__uint_as_float(vm[j]), __uint_as_float(vl[j]));  // large + small terms, RN
          float t = __fadd_rn(__fmul_rn(a, s_scale[c0 + j]), s_shift[c0 + j]);
          t = p.act == DSIN_ACT_RELU ? fmaxf(t, 0.f) : t;
          f[j] = any_res ? __fadd_rn(t, rs[chunk * 16 + j]) : t;
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 uh, ul;
          __half2* hh = reinterpret_cast<__half2*>(&uh);
          __half2* ll = reinterpret_cast<__half2*>(&ul);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x0 = f[g * 8 + 2 * e], x1 = f[g * 8 + 2 * e + 1];
            const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
            hh[e] = __halves2half2(h0, h1);
            ll[e] = __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
          }
          const uint32_t so = row_off + (uint32_t)(((2 * chunk + g) ^ sw) << 4);
          sts16(stg_hi + so, uh);
          sts16(stg_lo + so, ul);
        }
      }
      // accumulators drained: hand them back before the (asynchronous) stores
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[acc], 0);  // the leader's accumulator-empty barrier
      fence_proxy_async();  // generic-proxy writes to the staging blocks -> visible to the TMA stores
      __syncwarp();
      if (lane == 0 && tvalid) {  // the stores clip rows / pixels past the image
        tma_store_4d(&tm_yh, stg_hi, half * 64, ox0, oy0, n);
        tma_store_4d(&tm_yl, stg_lo, half * 64, ox0, oy0, n);
      }
    }
    if (lane == 0) tma_store_wait_all();  // global writes complete before the kernel ends
    __syncwarp();
  }

  __syncthreads();
  cluster_sync_all();  // no CTA of the pair may exit (or free TMEM) while the other can still signal it
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace

int conv_h3_launch(dsin_handle_t h, const __half* x_hi, const __half* x_lo, const __half* w_hi, const __half* w_lo,
                   const ConvH3Args& a, cudaStream_t st) {
  CUtensorMap txh, txl, twh, twl, tyh, tyl;
  const uint64_t xd[4] = {128, (uint64_t)a.OW, (uint64_t)a.OH, (uint64_t)a.n};
  const uint64_t xs[3] = {256, (uint64_t)a.OW * 256, (uint64_t)a.OH * a.OW * 256};
  const uint32_t xb[4] = {64, PITCH, HALO_R, 1};
  const uint32_t yb[4] = {64, TC, 4, 1};  // one epilogue warp's block: 64 channels x 8 pixels x 4 rows
  const uint64_t wd[2] = {128, 9 * 128};
  const uint64_t wsb[1] = {256};
  const uint32_t wb[2] = {64, 64};
  const CUtensorMapDataType f16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if (!encode_tmap(&txh, f16, 4, x_hi, xd, xs, xb, sw) || !encode_tmap(&txl, f16, 4, x_lo, xd, xs, xb, sw) ||
      !encode_tmap(&twh, f16, 2, w_hi, wd, wsb, wb, sw) || !encode_tmap(&twl, f16, 2, w_lo, wd, wsb, wb, sw) ||
      !encode_tmap(&tyh, f16, 4, a.yh, xd, xs, yb, sw) || !encode_tmap(&tyl, f16, 4, a.yl, xd, xs, yb, sw))
    return dsin_fail(h, DSIN_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", __func__);
  static bool configured[DSIN_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!configured[h->device]) {
    if (cudaFuncSetAttribute(conv_h3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: cannot raise dynamic shared memory", __func__);
    configured[h->device] = true;
  }
  ConvH3Args p = a;
  p.tiles_w = (a.OW + TC - 1) / TC;
  p.tiles_h = (a.OH + TR - 1) / TR;
  p.total_tiles = a.n * p.tiles_w * p.tiles_h;
  const int pairs = (p.total_tiles + 1) / 2;
  int clusters = h->sm_count / 2;
  if (clusters > pairs) clusters = pairs;
  if (launch_pdl(conv_h3_kernel, dim3(2 * clusters), dim3(NTHREADS), SMEM_BYTES, st, txh, txl, twh, twl, tyh, tyl, p) !=
      cudaSuccess)
    return dsin_fail(h, DSIN_ERR_CUDA, "%s: launch failed", __func__);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}
