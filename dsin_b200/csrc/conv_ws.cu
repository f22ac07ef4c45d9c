// K1, fp16-operand form ("terms = 1"): the 3x3 128->128 trunk convolution as a WEIGHT-STATIONARY tcgen05
// kernel with a halo-resident activation tile.  Same arithmetic and epilogue as conv_tc2_kernel<1>
// (src/autoencoder_imgcomp.py:229-234,257-262,275-288): y = act(conv(x) * scale + shift) + r1 + r2 on NHWC fp16.
//
// Why a second kernel.  With one MMA per product the tap-streaming kernel (conv_tc2.cu) is no longer bound by the
// tensor pipe but by the L2 -> shared-memory fill: per 128-pixel tile it re-fetches the activation tile for each
// of the 9 taps (288 KB) and the weight slabs again for every tile (144 KB), 94 B/cycle/SM against 4608 MMA cycles
// -- 2.4x more than the L2 delivers to 148 SMs at once.  Here
//   * the CTA keeps ITS HALF OF THE WHOLE FILTER in shared memory for the lifetime of the launch:
//     9 taps x 64 couts x 128 cin fp16 = 144 KB, loaded once (CTA pair, cta_group::2: each CTA supplies 64 of the
//     128 couts of the N = 128 operand);
//   * the activation tile is loaded ONCE with its halo -- 16 x 8 output pixels read 18 x 10 input pixels -- and
//     the nine taps are nine shared-memory descriptors into that one tile: the tile is stored at a 16-pixel row
//     pitch (2048 B, so every 8-pixel row group starts at a fixed phase of the 128-byte swizzle), the tap's row
//     shift is a multiple of the pitch, and its column shift (dx+1) x 128 B goes into the descriptor start
//     address together with the descriptor's 3-bit base-offset field (the swizzle phase of the first row).
//     L2 -> shared traffic per tile: 2 x 36 KB instead of 432 KB.
//   * residual inputs are NOT fed through the tensor pipe (conv_tc2's identity MMAs would cost 11-22 % here):
//     every epilogue thread prefetches its pixel's residual channels into registers while the tile's MMAs are
//     still running, so the loads are off the critical path.
// TMA out-of-bounds zero fill provides the SAME padding as before (negative / past-the-end box origins).
#include "tc_common.cuh"
#include "conv_tc.cuh"

using namespace tc;

namespace {

constexpr int TR = 16, TC = 8;              // output tile: 16 rows x 8 pixels = 128 GEMM rows, m = r * 8 + c
constexpr int HALO_R = TR + 2;              // 18 input rows
constexpr int PITCH = 16;                   // pixels per stored row (10 are used)
constexpr int A_STAGE = HALO_R * PITCH * 128;  // 36 864 B: one 64-channel chunk of the halo tile
constexpr int W_SLAB = 64 * 128;            // one (tap, chunk) slab of this CTA's 64 couts
constexpr int W_BYTES = 18 * W_SLAB;        // 147 456 B
constexpr int NSTAGE = 2;
constexpr int SMEM_BYTES = W_BYTES + NSTAGE * A_STAGE + 2048 + 1024;

__device__ __forceinline__ uint4 ld_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void add_half8(float* f, const uint4& u) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = __fadd_rn(f[2 * i], t.x);
    f[2 * i + 1] = __fadd_rn(f[2 * i + 1], t.y);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
conv_ws_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ ConvWsArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* wsm = smem;                       // [tap 9][chunk 2][64 couts][128 B], 128-byte swizzle
  uint8_t* tiles = smem + W_BYTES;           // NSTAGE halo chunks
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + NSTAGE * A_STAGE);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wfull + 1);
  float* s_scale = reinterpret_cast<float*>(tmem_ptr + 2);
  float* s_shift = s_scale + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);  // 8 epilogue warps x 2 CTAs (the leader's copy is the one used)
    }
    mbar_init(wfull, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_w);
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  if (warp == 1) tmem_alloc2(tmem_ptr, 256);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int pairs = (p.total_tiles + 1) / 2;
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {  // this CTA's half of the filter, once
      if (leader) mbar_expect_tx(wfull, 2u * (uint32_t)W_BYTES);
      for (int s = 0; s < 18; ++s)
        tma2_load_2d(wsm + s * W_SLAB, &tm_w, wfull, (s & 1) * 64, (s >> 1) * 128 + (int)rank * 64);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int pi = cid; pi < pairs; pi += nclusters) {
      int tile = 2 * pi + (int)rank;
      if (tile >= p.total_tiles) tile = p.total_tiles - 1;  // odd tail: recompute a valid tile, never stored
      const int tw = tile % p.tiles_w, t2 = tile / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      for (int cc = 0; cc < 2; ++cc) {
        mbar_wait(&empty[stage], phase ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full[stage], 2u * (uint32_t)A_STAGE);
          tma2_load_4d(tiles + stage * A_STAGE, &tm_x, &full[stage], cc * 64, tw * TC - 1, th * TR - 1, n);
        }
        __syncwarp();
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(256, 128, 0);
      const uint32_t w_base = smem_u32(wsm);
      mbar_wait(wfull, 0);
      fence_after_sync();
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int pi = cid; pi < pairs; pi += nclusters, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 128u;
        for (int cc = 0; cc < 2; ++cc) {
          mbar_wait(&full[stage], phase);
          fence_after_sync();
          if (elect_one()) {
            const uint32_t sa = smem_u32(tiles + stage * A_STAGE);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int ky = tap / 3, kx = tap % 3;  // input pixel (r + ky, c + kx) of the halo tile
              // row group g = output row r: 8 pixels x 128 B, groups PITCH x 128 B apart; the first row of every
              // group sits kx rows into its 1024-byte swizzle period -> base offset kx
              const uint64_t a_desc = make_smem_desc(sa + (uint32_t)((ky * PITCH + kx) * 128), 16, PITCH * 128,
                                                     LAYOUT_SW128) |
                                      ((uint64_t)(p.base_offset_mode ? kx : 0) << 49);
              const uint64_t b_desc = make_smem_desc(w_base + (uint32_t)((tap * 2 + cc) * W_SLAB), 16, 1024, LAYOUT_SW128);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma2_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (cc | tap | k) ? 1u : 0u);
            }
            umma2_commit(&empty[stage]);  // frees this stage in BOTH CTAs
          }
          __syncwarp();
          if (++stage == NSTAGE) {
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
    // Two warps per TMEM lane quarter, 64 output channels each; thread = pixel m = q * 32 + lane = (r, c).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int rl = row >> 3, cl = row & 7;
    int it = 0;
    for (int pi = cid; pi < pairs; pi += nclusters, ++it) {
      const int acc = it & 1;
      const int tile = 2 * pi + (int)rank;
      const bool tvalid = tile < p.total_tiles;
      const int tcl = tvalid ? tile : p.total_tiles - 1;
      const int tw = tcl % p.tiles_w, t2 = tcl / p.tiles_w;
      const int th = t2 % p.tiles_h, n = t2 / p.tiles_h;
      const int oy = th * TR + rl, ox = tw * TC + cl;
      const bool valid = tvalid && oy < p.OH && ox < p.OW;
      const size_t off = (((size_t)n * p.OH + oy) * p.OW + ox) * 128 + half * 64;
      // residual channels of this pixel, fetched while the tile's MMAs are in flight
      uint4 R1[8], R2[8];
      if (p.r1 && valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) R1[j] = ld_nc16(p.r1 + off + j * 8);
      }
      if (p.r2 && valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) R2[j] = ld_nc16(p.r2 + off + j * 8);
      }
      mbar_wait(&tfull[acc], (uint32_t)(it >> 1) & 1u);
      fence_after_sync();
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128 + half * 64);
#pragma unroll
      for (int chunk = 0; chunk < 4; ++chunk) {
        const int c0 = half * 64 + chunk * 16;
        uint32_t v[16];
        tmem_ld_32x16(lane_base + (uint32_t)(chunk * 16), v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t = __fadd_rn(__fmul_rn(__uint_as_float(v[j]), s_scale[c0 + j]), s_shift[c0 + j]);
            f[j] = p.act == DSIN_ACT_RELU ? fmaxf(t, 0.f) : t;
          }
          if (p.r1) {
            add_half8(f, R1[2 * chunk]);
            add_half8(f + 8, R1[2 * chunk + 1]);
          }
          if (p.r2) {
            add_half8(f, R2[2 * chunk]);
            add_half8(f + 8, R2[2 * chunk + 1]);
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 uh;
            __half2* hh = reinterpret_cast<__half2*>(&uh);
#pragma unroll
            for (int e = 0; e < 4; ++e) hh[e] = __floats2half2_rn(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1]);
            reinterpret_cast<uint4*>(p.y + off + chunk * 16)[g] = uh;
          }
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
    tmem_dealloc2(tmem_base, 256);
  }
}

}  // namespace

int conv_ws_launch(dsin_handle_t h, const __half* x, const __half* w_packed, const ConvWsArgs& a, cudaStream_t st) {
  CUtensorMap tx, tw;
  const uint64_t xd[4] = {128, (uint64_t)a.OW, (uint64_t)a.OH, (uint64_t)a.n};
  const uint64_t xs[3] = {256, (uint64_t)a.OW * 256, (uint64_t)a.OH * a.OW * 256};
  const uint32_t xb[4] = {64, PITCH, HALO_R, 1};
  const uint64_t wd[2] = {128, 9 * 128};
  const uint64_t wsb[1] = {256};
  const uint32_t wb[2] = {64, 64};
  if (!encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, xd, xs, xb, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !encode_tmap(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w_packed, wd, wsb, wb, CU_TENSOR_MAP_SWIZZLE_128B))
    return dsin_fail(h, DSIN_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", __func__);
  static bool configured[DSIN_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!configured[h->device]) {
    if (cudaFuncSetAttribute(conv_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: cannot raise dynamic shared memory", __func__);
    configured[h->device] = true;
  }
  ConvWsArgs p = a;
  p.tiles_w = (a.OW + TC - 1) / TC;
  p.tiles_h = (a.OH + TR - 1) / TR;
  p.total_tiles = a.n * p.tiles_w * p.tiles_h;
  const int pairs = (p.total_tiles + 1) / 2;
  int clusters = h->sm_count / 2;
  if (clusters > pairs) clusters = pairs;
  conv_ws_kernel<<<2 * clusters, 320, SMEM_BYTES, st>>>(tx, tw, p);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}
