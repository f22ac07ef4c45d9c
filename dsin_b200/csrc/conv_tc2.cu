// K1 on CTA PAIRS: the 128-channel convolutions (the 64 trunk layers and h2) with tcgen05 cta_group::2.
//   same arithmetic and epilogue as conv_tc_kernel (src/autoencoder_imgcomp.py:224,229-234,257-262,275-288).
//
// Why: with one CTA per tile, a 128x128x16 MMA reads 8 KB of operands from shared memory every 64 cycles
// (128 B/cycle = the whole shared-memory bandwidth of the SM) while TMA is refilling the stages; the
// tensor pipe measured 57-70 % active with L2 and DRAM far from saturated (profiles/r1_v5_*).  A CTA pair
// (2-CTA cluster on one TPC) issues ONE M=256 MMA for two pixel tiles: each CTA keeps its own 128-pixel
// A tile, but the weight slab B is split -- each CTA loads and holds 64 of the 128 couts -- so per SM the
// operand reads drop to 6 KB per MMA and the TMA fill to 48 KB per k-block (4 stages instead of 3).
//
// Protocol (leader = cluster rank 0): both CTAs' producers issue their TMA loads against the LEADER's
// full barrier (cta_group::2 TMA); the leader's MMA warp issues tcgen05.mma.cta_group::2 and commits with
// multicast to the per-CTA empty / accumulator-full barriers; every epilogue warp of both CTAs arrives on
// the leader's accumulator-empty barrier.  Accumulators stay in each CTA's own TMEM (double buffered).
//
// Residual inputs (the skip adds of src/autoencoder_imgcomp.py:275-288) are added ON THE TENSOR PIPE: each
// residual plane tile is TMA-loaded as an A operand and multiplied by a 128x128 identity slab into a second
// accumulator (TMEM columns +128; products with 1.0 are exact, fp32 accumulation), 2 extra pipeline stages
// per residual tensor.  The epilogue then reads both accumulators from TMEM and never touches global memory
// for inputs: the previous per-thread 16-byte residual loads at a 256-byte stride kept the tensor pipe only
// 58 % busy on those layers (profiles/r1_v7_ncu_full_conv_tc2_b8.txt).
#include <vector>

#include "tc_common.cuh"
#include "conv_tc.cuh"

using namespace tc;

namespace {

constexpr int BW = 16, BH = 8;
constexpr int A_TILE = 128 * 128;  // 128 px x 64 ch fp16
constexpr int B_HALF = 64 * 128;   // 64 couts x 64 ch fp16

template <int TERMS>
struct Cfg2 {
  // TERMS == 1 stages keep room for the lo plane of a residual tile
  static constexpr int kStage = TERMS == 3 ? 2 * (A_TILE + B_HALF) : 2 * A_TILE + B_HALF;
  static constexpr int kStages = TERMS == 3 ? 4 : 5;
  static constexpr int kSmem = kStages * kStage + 1024 + 2048;
};

template <int TERMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
                const __grid_constant__ ConvTc2Res rm, const __grid_constant__ ConvTc2Args p) {
  using C = Cfg2<TERMS>;
  constexpr int S = C::kStages;
  constexpr int STAGE = C::kStage;
  constexpr int OFF_B = A_TILE, OFF_ALO = A_TILE + B_HALF, OFF_BLO = 2 * A_TILE + B_HALF;
  constexpr int OFF_ALO_RES = A_TILE + B_HALF;  // lo plane of a residual (TERMS == 1 stages are padded to hold it)
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
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);  // 8 epilogue warps x 2 CTAs (only the leader's copy is used)
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
  if (warp == 1) tmem_alloc2(tmem_ptr, 512);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int num_kb = p.ntaps * p.nchunks;
  const int pairs = (p.total_tiles + 1) / 2;
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  constexpr uint32_t kBytesCta = (TERMS == 3 ? 2u : 1u) * (uint32_t)(A_TILE + B_HALF);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int pi = cid; pi < pairs; pi += nclusters) {
      int tile = 2 * pi + (int)rank;
      if (tile >= p.total_tiles) tile = p.total_tiles - 1;  // odd tail: recompute a valid tile, never stored
      const int tw = tile % p.tiles_w, t2 = tile / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      const int x0 = tw * BW * p.in_step, y0 = th * BH * p.in_step;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / p.nchunks, cc = kb - tap * p.nchunks;
        mbar_wait(&empty[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* st = tiles + stage * STAGE;
          if (leader) mbar_expect_tx(&full[stage], 2u * kBytesCta);
          const int ax = x0 + p.dx[tap], ay = y0 + p.dy[tap], wrow = p.wi[tap] * 128 + (int)rank * 64;
          tma2_load_4d(st, &tm_xh, &full[stage], cc * 64, ax, ay, n);
          tma2_load_2d(st + OFF_B, &tm_wh, &full[stage], cc * 64, wrow);
          if (TERMS == 3) {
            tma2_load_4d(st + OFF_ALO, &tm_xl, &full[stage], cc * 64, ax, ay, n);
            tma2_load_2d(st + OFF_BLO, &tm_wl, &full[stage], cc * 64, wrow);
          }
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
      // residual tensors ride the same pipeline: plane tiles as A operands, an identity slab as B
      for (int r = 0; r < rm.nres; ++r)
        for (int cc = 0; cc < 2; ++cc) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (elect_one()) {
            uint8_t* st = tiles + stage * STAGE;
            const uint32_t bytes = (uint32_t)(A_TILE + B_HALF) + (rm.has_lo[r] ? (uint32_t)A_TILE : 0u);
            if (leader) mbar_expect_tx(&full[stage], 2u * bytes);
            tma2_load_4d(st, &rm.plane[2 * r], &full[stage], cc * 64, tw * BW, th * BH, n);
            tma2_load_2d(st + OFF_B, &rm.ident, &full[stage], cc * 64, (int)rank * 64);
            if (rm.has_lo[r]) tma2_load_4d(st + OFF_ALO_RES, &rm.plane[2 * r + 1], &full[stage], cc * 64, tw * BW, th * BH, n);
          }
          __syncwarp();
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(256, 128, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int pi = cid; pi < pairs; pi += nclusters, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        // Without residual inputs the second accumulator (columns +128) takes the small product terms hi*lo + lo*hi:
        // tcgen05 accumulates with round-toward-zero, and 2 x 4 x num_kb small-term MMAs must neither pay nor add
        // roundings at the magnitude of the hi*hi sum (h2: 25 taps -> 300 MMAs in one accumulator otherwise).
        const uint32_t d_small = (TERMS == 3 && rm.nres == 0) ? d_tmem + 128u : d_tmem;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          fence_after_sync();
          if (elect_one()) {
            const uint32_t sa = smem_u32(tiles + stage * STAGE);
            const uint64_t a_hi = make_smem_desc(sa, 16, 1024, LAYOUT_SW128);
            const uint64_t b_hi = make_smem_desc(sa + OFF_B, 16, 1024, LAYOUT_SW128);
            const uint64_t a_lo = make_smem_desc(sa + OFF_ALO, 16, 1024, LAYOUT_SW128);
            const uint64_t b_lo = make_smem_desc(sa + OFF_BLO, 16, 1024, LAYOUT_SW128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma2_f16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) ? 1u : 0u);
              if (TERMS == 3) {
                umma2_f16(d_small, a_hi + 2 * k, b_lo + 2 * k, idesc, (d_small == d_tmem || (kb | k)) ? 1u : 0u);
                umma2_f16(d_small, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
              }
            }
            umma2_commit(&empty[stage]);  // frees this stage in BOTH CTAs
          }
          __syncwarp();
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // residuals: D2 (columns +128) = sum over planes of  plane x identity  (exact products, fp32 accumulation)
        for (int r = 0; r < rm.nres; ++r)
          for (int cc = 0; cc < 2; ++cc) {
            mbar_wait(&full[stage], phase);
            fence_after_sync();
            if (elect_one()) {
              const uint32_t sa = smem_u32(tiles + stage * STAGE);
              const uint64_t a_hi = make_smem_desc(sa, 16, 1024, LAYOUT_SW128);
              const uint64_t b_id = make_smem_desc(sa + OFF_B, 16, 1024, LAYOUT_SW128);
              const uint64_t a_lo = make_smem_desc(sa + OFF_ALO_RES, 16, 1024, LAYOUT_SW128);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma2_f16(d_tmem + 128u, a_hi + 2 * k, b_id + 2 * k, idesc, (r | cc | k) ? 1u : 0u);
                if (rm.has_lo[r]) umma2_f16(d_tmem + 128u, a_lo + 2 * k, b_id + 2 * k, idesc, 1u);
              }
              umma2_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == S) {
              stage = 0;
              phase ^= 1u;
            }
          }
        if (elect_one()) umma2_commit(&tfull[acc]);  // accumulators complete in both CTAs
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..9 (both CTAs, own TMEM)
    // Two warps per TMEM lane quarter, 64 columns each.  Residual sums arrive in the second accumulator
    // (columns +128) -- no global loads here.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int hl = row >> 4, wl = row & 15;
    const bool has_res = rm.nres > 0;
    const bool has_small = TERMS == 3 && !has_res;  // columns +128 hold the hi*lo + lo*hi products
    int it = 0;
    for (int pi = cid; pi < pairs; pi += nclusters, ++it) {
      const int acc = it & 1;
      const int tile = 2 * pi + (int)rank;
      const bool tvalid = tile < p.total_tiles;
      const int tcl = tvalid ? tile : p.total_tiles - 1;
      const int tw = tcl % p.tiles_w, t2 = tcl / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      const int oy = th * BH + hl, ox = tw * BW + wl;
      const bool valid = tvalid && oy < p.OH && ox < p.OW;
      const size_t pix = ((size_t)n * p.OH + oy) * p.OW + ox;
      mbar_wait(&tfull[acc], (uint32_t)(it >> 1) & 1u);
      fence_after_sync();
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll 1
      for (int chunk = half * 4; chunk < half * 4 + 4; ++chunk) {
        const int c0 = chunk * 16;
        uint32_t v[16], r[16];
        tmem_ld_32x16(lane_base + (uint32_t)c0, v);
        if (has_res || has_small) tmem_ld_32x16(lane_base + 128u + (uint32_t)c0, r);
        tmem_ld_wait();
        if (!valid) continue;
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(v[j]);
          if (has_small) a = __fadd_rn(a, __uint_as_float(r[j]));  // large + small product terms, round to nearest
          float t = __fadd_rn(__fmul_rn(a, s_scale[c0 + j]), s_shift[c0 + j]);
          t = p.act == DSIN_ACT_RELU ? fmaxf(t, 0.f) : t;
          f[j] = has_res ? __fadd_rn(t, __uint_as_float(r[j])) : t;
        }
        const size_t off = pix * 128 + c0;
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
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[acc], 0);  // the leader's accumulator-empty barrier
    }
  }

  __syncthreads();
  cluster_sync_all();  // no CTA of the pair may exit (or free TMEM) while the other can still signal it
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc2(tmem_base, 512);
  }
}

template <int TERMS>
int launch2(dsin_handle_t h, const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& wh,
            const CUtensorMap& wl, const ConvTc2Res& rm, const ConvTc2Args& p, cudaStream_t st) {
  using C = Cfg2<TERMS>;
  static bool configured[DSIN_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!configured[h->device]) {
    if (cudaFuncSetAttribute(conv_tc2_kernel<TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem) !=
        cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: cannot raise dynamic shared memory", __func__);
    configured[h->device] = true;
  }
  const int pairs = (p.total_tiles + 1) / 2;
  int clusters = h->sm_count / 2;
  if (clusters > pairs) clusters = pairs;
  conv_tc2_kernel<TERMS><<<2 * clusters, 320, C::kSmem, st>>>(xh, xl, wh, wl, rm, p);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}

}  // namespace

int conv_tc2_launch(dsin_handle_t h, int terms, const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& wh,
                    const CUtensorMap& wl, const ConvTc2Args& p, cudaStream_t st) {
  ConvTc2Res rm;
  memset(&rm, 0, sizeof(rm));
  const __half* hi[2] = {p.r1h, p.r2h};
  const __half* lo[2] = {p.r1l, p.r2l};
  if (hi[0] || hi[1]) {
    if (!h->ident128) {  // one-time: 128x128 fp16 identity, the B operand that adds a residual tile on the MMA
      std::vector<__half> eye(128 * 128, __float2half(0.f));
      for (int i = 0; i < 128; ++i) eye[i * 128 + i] = __float2half(1.f);
      if (cudaMalloc(&h->ident128, eye.size() * sizeof(__half)) != cudaSuccess ||
          cudaMemcpy(h->ident128, eye.data(), eye.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess)
        return dsin_fail(h, DSIN_ERR_CUDA, "%s: cannot create the identity slab", __func__);
    }
    const uint64_t id_d[2] = {128, 128};
    const uint64_t id_s[1] = {256};
    const uint32_t id_b[2] = {64, 64};
    bool ok = encode_tmap(&rm.ident, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, h->ident128, id_d, id_s, id_b,
                          CU_TENSOR_MAP_SWIZZLE_128B);
    const uint64_t d4[4] = {128, (uint64_t)p.OW, (uint64_t)p.OH, (uint64_t)p.n};
    const uint64_t s3[3] = {256, (uint64_t)p.OW * 256, (uint64_t)p.OH * p.OW * 256};
    const uint32_t b4[4] = {64, BW, BH, 1};
    for (int i = 0; i < 2; ++i) {
      if (!hi[i]) continue;
      const int r = rm.nres++;
      rm.has_lo[r] = lo[i] != nullptr;
      ok = ok && encode_tmap(&rm.plane[2 * r], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, hi[i], d4, s3, b4,
                             CU_TENSOR_MAP_SWIZZLE_128B);
      ok = ok && encode_tmap(&rm.plane[2 * r + 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, lo[i] ? lo[i] : hi[i], d4, s3,
                             b4, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (!ok) return dsin_fail(h, DSIN_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", __func__);
  }
  return terms == 3 ? launch2<3>(h, xh, xl, wh, wl, rm, p, st) : launch2<1>(h, xh, xl, wh, wl, rm, p, st);
}
