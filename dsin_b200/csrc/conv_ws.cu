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
//     the nine taps are nine shared-memory descriptors into that one tile: output row r of the tile is one 8-row
//     group of the MMA's M dimension (8 pixels x 128 B), groups one stored row (10 pixels, 1280 B) apart (SBO), and
//     a tap only moves
//     the descriptor's start address by (ky x pitch + kx) x 128 B.  This works because both TMA and the UMMA
//     descriptor derive the 128-byte-swizzle phase from the ABSOLUTE shared-memory address bits [7:9] (measured:
//     tools/conv_ws_debug.py -- with start addresses that are only 128-byte aligned the operand rows come out
//     right with the descriptor's base-offset field left 0, and wrong with it set).
//     L2 -> shared traffic per tile: 2 x 22.5 KB instead of 432 KB; the next tile's boxes are prefetched into L2 one tile ahead.
//   * the epilogue is COALESCED.  A TMEM row is a pixel, so a warp's natural 16-byte accesses land in 32
//     different 128-byte lines (pixels are 256 B apart): 32 L1TEX wavefronts per instruction, ~2000 cycles of
//     stores and ~2000 per residual tensor per tile against 4608 MMA cycles -- the first version of this kernel
//     was bound by exactly that (profiles/r2_v2_ncu_conv_ws.txt: tensor pipe 59 % / 40 % with residuals).  Now
//     every epilogue warp owns a 4 KB swizzled staging block (its 32 pixels x 64 channels): results go
//     registers -> staging -> ONE TMA store per warp and tile (which also clips partial tiles), and residual
//     tensors come in through cp.async with 4 pixels x 128 contiguous bytes per instruction, are read back
//     conflict-free, and are fetched while the tile's MMAs are still running.
// TMA out-of-bounds zero fill provides the SAME padding as before (negative / past-the-end box origins).
#include "tc_common.cuh"
#include "conv_tc.cuh"

using namespace tc;

namespace {

constexpr int TR = 16, TC = 8;              // output tile: 16 rows x 8 pixels = 128 GEMM rows, m = r * 8 + c
constexpr int HALO_R = TR + 2;              // 18 input rows
constexpr int PITCH = TC + 2;               // 10 input pixels per stored row, stored densely (1280-byte rows)
constexpr int A_BYTES = HALO_R * PITCH * 128;  // 23 040 B: one 64-channel chunk of the halo tile
constexpr int A_STAGE = (A_BYTES + 1023) / 1024 * 1024;  // stage bases stay 1024-byte aligned
constexpr int W_SLAB = 64 * 128;            // one (tap, chunk) slab of this CTA's 64 couts
constexpr int W_BYTES = 18 * W_SLAB;        // 147 456 B
constexpr int NSTAGE = 2;
constexpr int STG_WARP = 32 * 128;          // one epilogue warp's staging block: 32 pixels x 64 channels fp16
constexpr int STG_BYTES = 8 * STG_WARP;
constexpr int SMEM_BYTES = W_BYTES + NSTAGE * A_STAGE + STG_BYTES + 2048 + 1024;

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
               const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ ConvWsArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* wsm = smem;                       // [tap 9][chunk 2][64 couts][128 B], 128-byte swizzle
  uint8_t* tiles = smem + W_BYTES;           // NSTAGE halo chunks
  uint8_t* stg = tiles + NSTAGE * A_STAGE;   // 8 x 4 KB epilogue staging blocks (1024-byte aligned)
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + STG_BYTES);
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
    prefetch_tmap(&tm_y);
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
  if (warp == 0) {  // this CTA's half of the filter, once -- it does not depend on the previous layer either
    if (elect_one()) {
      if (leader) mbar_expect_tx(wfull, 2u * (uint32_t)W_BYTES);
      for (int s = 0; s < 18; ++s)
        tma2_load_2d(wsm + s * W_SLAB, &tm_w, wfull, (s & 1) * 64, (s >> 1) * 128 + (int)rank * 64);
    }
    __syncwarp();
  }
  // everything above is independent of the previous layer; its output (our input, our residuals) is read from here on
  pdl_wait();
  pdl_launch_dependents();
  const int pairs = (p.total_tiles + 1) / 2;
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
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
        mbar_wait(&empty[stage], phase ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full[stage], 2u * (uint32_t)A_BYTES);
          tma2_load_4d(tiles + stage * A_STAGE, &tm_x, &full[stage], cc * 64, tw * TC - 1, th * TR - 1, n);
          if (has_next) tma_prefetch_4d(&tm_x, cc * 64, ntw * TC - 1, nth * TR - 1, nn);
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
              // row group g = output row r: 8 pixels x 128 B, groups PITCH x 128 B apart
              const uint64_t a_desc = make_smem_desc(sa + (uint32_t)((ky * PITCH + kx) * 128), 16, PITCH * 128,
                                                     LAYOUT_SW128);
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
    // Two warps per TMEM lane quarter, 64 output channels each; thread = pixel m = q * 32 + lane = (r, c), i.e. the
    // warp's 32 pixels are rows 4q..4q+3 of the tile.  Staging block: pixel p at p * 128 B, its 16-byte piece j at
    // ((j ^ (p & 7)) * 16) -- the 128-byte-swizzle layout of a (64 ch, 8 px, 4 rows) TMA box.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t my_stg = smem_u32(stg + (warp - 2) * STG_WARP);
    const uint32_t my_row = my_stg + (uint32_t)lane * 128u;   // this thread's pixel in the staging block
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
      // the previous tile's TMA store must have finished READING the staging block
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
      // residual channels of the warp's pixels: cp.async, 4 pixels x 128 contiguous bytes per instruction, fetched
      // while the tile's MMAs are still in flight; then every thread reads its own pixel back (conflict-free)
      uint4 R1[8], R2[8];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const __half* rsrc = rr == 0 ? p.r1 : p.r2;
        if (rsrc == nullptr) continue;  // warp-uniform
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pp = 4 * i + (lane >> 3), j = lane & 7;  // pixel of the block, 16-byte piece
          const int oy = oy0 + (pp >> 3), ox = ox0 + (pp & 7);
          const bool ok = tvalid && oy < p.OH && ox < p.OW;
          const size_t goff = ok ? ((((size_t)n * p.OH + oy) * p.OW + ox) * 128 + half * 64 + j * 8) : 0;
          cp_async16(my_stg + (uint32_t)(pp * 128 + ((j ^ (pp & 7)) << 4)), rsrc + goff, ok);
        }
        cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 v = lds16(my_row + (uint32_t)((j ^ sw) << 4));
          if (rr == 0) R1[j] = v; else R2[j] = v;
        }
        __syncwarp();  // everyone has read before the block is overwritten
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
          sts16(my_row + (uint32_t)(((2 * chunk + g) ^ sw) << 4), uh);
        }
      }
      // accumulator drained: hand it back before the (asynchronous) store
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[acc], 0);  // the leader's accumulator-empty barrier
      fence_proxy_async();  // generic-proxy writes to the staging block -> visible to the TMA store
      __syncwarp();
      if (lane == 0 && tvalid) tma_store_4d(&tm_y, my_stg, half * 64, ox0, oy0, n);  // clips rows / pixels past the image
    }
    if (lane == 0) tma_store_wait_all();  // global writes complete before the kernel ends
    __syncwarp();
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
  CUtensorMap tx, tw, ty;
  const uint64_t xd[4] = {128, (uint64_t)a.OW, (uint64_t)a.OH, (uint64_t)a.n};
  const uint64_t xs[3] = {256, (uint64_t)a.OW * 256, (uint64_t)a.OH * a.OW * 256};
  const uint32_t xb[4] = {64, PITCH, HALO_R, 1};
  const uint32_t yb[4] = {64, TC, 4, 1};  // one epilogue warp's block: 64 channels x 8 pixels x 4 rows
  const uint64_t wd[2] = {128, 9 * 128};
  const uint64_t wsb[1] = {256};
  const uint32_t wb[2] = {64, 64};
  if (!encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, xd, xs, xb, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !encode_tmap(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w_packed, wd, wsb, wb, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !encode_tmap(&ty, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.y, xd, xs, yb, CU_TENSOR_MAP_SWIZZLE_128B))
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
  if (launch_pdl(conv_ws_kernel, dim3(2 * clusters), dim3(320), SMEM_BYTES, st, tx, tw, ty, p) != cudaSuccess)
    return dsin_fail(h, DSIN_ERR_CUDA, "%s: launch failed", __func__);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}
