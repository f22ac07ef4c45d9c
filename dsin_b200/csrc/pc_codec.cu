// PC1 entropy coder on the GPU: real bitstreams from the probability model.
//
// The reference only ships the pieces of a coder (src/probclass_imgcomp.py:361-482: one symbol at a time, each with
// the frequencies the context model predicts from the symbols already coded).  The bit-exact definition this kernel
// implements -- operation order of the fp32 network, frequency quantisation, range coder, symbol order -- is written
// down in oracle/pc_codec.c; the two must produce identical bytes.
//
// Parallel schedule.  The receptive field of the four masked (2,3,3) convolutions (probclass_imgcomp.py:150-176,
// 214-261) reaches one step back per layer, so position p = (d, h, w) of every layer only needs positions whose
// wavefront time 25 d + 5 h + w is smaller.  One CTA owns one stream = the depth slices d == stream (mod nstreams)
// and walks a slice in steps of u = 5 h + w (<= 33 positions per step).  Inside the CTA the step is a two-stage
// software pipeline (see pc_codec_kernel): a coder warp range-codes step u while nine bulk warps pre-accumulate
// the taps of step u + 1 that are already old enough, then finish the chains and build the frequency tables.
// Slice d may run u + 6 steps behind slice d - 1 (another CTA of the same image): a per-slice progress counter in
// global memory (release / acquire) is the only inter-CTA synchronisation, so the slices of an image form a
// software pipeline across its CTAs.  All CTAs of a launch must be co-resident (cooperative launch).
// The encoder does not need the schedule at all (every symbol is known): run_codec() takes the full-volume path
// (probclass.cu pc1_symbol_tables + pc_symbol_coder_kernel); the wavefront kernel in encode mode gives the same
// bytes and is kept as a cross-check (dsin_pc_encode_wavefront).
//
// Determinism.  Every activation is an fmaf chain in the order of oracle/pc_codec.c (bias; live taps in raster
// order; input channels ascending); exp is a fixed polynomial; only correctly rounded IEEE operations are used.
#include "pc_codec_common.cuh"

namespace {

using namespace pc1;

constexpr int K = 24;          // hidden channels
constexpr int KP = 28;         // padded channel stride of the transposed weights in shared memory (conflict-free LDS.128)
constexpr int NT0 = 13, NT = 14;
constexpr int MAXPOS = 33;
constexpr int BULK = 288;      // nine bulk warps: 33 positions x 8 channel triples = 264 working threads
constexpr int THREADS = 320;   // + the coder warp
constexpr int MAXH = 64;       // rows (with halo) of the shared-memory rings: symbol rows <= 58
constexpr int PROGRESS_INIT = -1000000;

__constant__ int c_taps[NT][3] = {{-1, -1, -1}, {-1, -1, 0}, {-1, -1, 1}, {-1, 0, -1}, {-1, 0, 0}, {-1, 0, 1}, {-1, 1, -1},
                                  {-1, 1, 0},   {-1, 1, 1},  {0, -1, -1}, {0, -1, 0}, {0, -1, 1}, {0, 0, -1}, {0, 0, 0}};

struct CodecArgs {
  int n, C, H, W, L, nstreams, decode, reset_status;
  const float* centers;
  const float *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3;
  float* q;        // [n][C][H+8][W+8]
  float* a0;       // [n][C][H+6][W+6][K]
  float* a1;       // [n][C][H+4][W+4][K]
  float* a2;       // [n][C][H+2][W+2][K]
  int* progress;   // [n][C]
  int64_t* sym;    // [n][C][H][W]   (input when encoding, output when decoding)
  uint8_t* bytes;  // [n][nstreams][cap]
  int64_t cap;
  int64_t* sizes;  // [n][nstreams]
  int* status;     // != 0: a stream overflowed its capacity
};

constexpr int WIN = 128;  // bytes of the stream staged per step (a step consumes <= 2 bytes per symbol)

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// NC fmaf chains (output channels) of one position advanced over one tap: x = the 24 input channels of the tap
// (shared memory), w0 = the tap's weights of chain 0 ([KP] floats), the other chains follow at chain_stride floats.
// Within a chain the order is the coder's: input channels ascending.
template <int NC>
__device__ __forceinline__ void chain_tap(float (&acc)[NC], const float* __restrict__ w0, int chain_stride,
                                          const float* __restrict__ x) {
  const float4* x4 = reinterpret_cast<const float4*>(x);
#pragma unroll
  for (int g = 0; g < K / 4; ++g) {
    const float4 xv = x4[g];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 wv = reinterpret_cast<const float4*>(w0 + c * chain_stride)[g];
      acc[c] = __fmaf_rn(xv.x, wv.x, acc[c]);
      acc[c] = __fmaf_rn(xv.y, wv.y, acc[c]);
      acc[c] = __fmaf_rn(xv.z, wv.z, acc[c]);
      acc[c] = __fmaf_rn(xv.w, wv.w, acc[c]);
    }
  }
}

__device__ __forceinline__ void bar_bulk() { asm volatile("bar.sync 1, %0;" ::"n"(BULK) : "memory"); }

// Software pipeline of one CTA (one stream).  Taps 0..11 of every layer only need data that is at least three
// steps old (previous slice, row h-1); only tap 12 (the left neighbour, previous step) and tap 13 (the position
// itself, this step) are on the critical path -- and the coder's chain order ends with exactly those two.  So while
// the coder warp range-codes step u, the nine bulk warps already accumulate taps 0..11 of step u+1 for all four
// layers (partial chains stay in registers); after the barrier they finish the chains with taps 12/13 out of
// shared-memory rings, layer by layer, and build the frequency tables of step u+1.
__global__ void __launch_bounds__(THREADS, 1) pc_codec_kernel(const CodecArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* s_w1 = smem;                       // [NT][K][KP]   transposed: (tap, co, ci)
  float* s_w2 = s_w1 + NT * K * KP;         // [NT][K][KP]
  float* s_w3 = s_w2 + NT * K * KP;         // [NT][MAXL][KP]
  float* s_w0 = s_w3 + NT * MAXL * KP;      // [NT0][K]
  float* s_b0 = s_w0 + NT0 * K;             // [K] x 3, [MAXL]
  float* s_b1 = s_b0 + K;
  float* s_b2 = s_b1 + K;
  float* s_b3 = s_b2 + K;
  float* s_k0 = s_b3 + MAXL;                // constant activations of the padding-only slices d < 0
  float* s_k1 = s_k0 + K;
  float* s_k2 = s_k1 + K;
  float* s_cent = s_k2 + K;                 // [MAXL]
  float* s_logit = s_cent + MAXL;           // [MAXPOS][MAXL]
  uint32_t* s_freq = reinterpret_cast<uint32_t*>(s_logit + MAXPOS * MAXL);  // [MAXPOS][MAXL] cumulative tables
  float* s_in = reinterpret_cast<float*>(s_freq + MAXPOS * MAXL);           // [MAXPOS][NT][K] staged old taps
  float* s_c0 = s_in + MAXPOS * NT * K;     // [2][MAXH][K] layer-0 values of the last two steps, by row
  float* s_c1 = s_c0 + 2 * MAXH * K;
  float* s_c2 = s_c1 + 2 * MAXH * K;
  float* s_qc = s_c2 + 2 * MAXH * K;        // [MAXH] centre value decoded in the last step, by row
  int* s_sym = reinterpret_cast<int*>(s_qc + MAXH);                         // [MAXPOS] symbols of the step (encode)
  long long* s_pos = reinterpret_cast<long long*>(s_sym + MAXPOS + 1);      // decoder read position (8-byte aligned)
  uint8_t* s_win = reinterpret_cast<uint8_t*>(s_pos + 1);                   // [WIN] stream bytes of the step (decode)

  const int tid = threadIdx.x;
  const int img = blockIdx.x / p.nstreams, stream = blockIdx.x % p.nstreams;
  const int C = p.C, H = p.H, W = p.W, L = p.L;
  const float pad = p.centers[0];

  // ---- stage the weights (transposed to (tap, co, ci), ci padded to KP) and the constants
  for (int i = tid; i < NT * K * KP; i += THREADS) {
    const int ci = i % KP, co = (i / KP) % K, t = i / (KP * K);
    s_w1[i] = ci < K ? p.w1[((size_t)t * K + ci) * K + co] : 0.f;
    s_w2[i] = ci < K ? p.w2[((size_t)t * K + ci) * K + co] : 0.f;
  }
  for (int i = tid; i < NT * MAXL * KP; i += THREADS) {
    const int ci = i % KP, co = (i / KP) % MAXL, t = i / (KP * MAXL);
    s_w3[i] = (ci < K && co < L) ? p.w3[((size_t)t * K + ci) * L + co] : 0.f;
  }
  for (int i = tid; i < NT0 * K; i += THREADS) s_w0[i] = p.w0[i];
  if (tid < K) { s_b0[tid] = p.b0[tid]; s_b1[tid] = p.b1[tid]; s_b2[tid] = p.b2[tid]; }
  if (tid < MAXL) { s_b3[tid] = tid < L ? p.b3[tid] : 0.f; s_cent[tid] = tid < L ? p.centers[tid] : 0.f; }
  __syncthreads();
  // activations of the padding-only slices d < 0 are position independent: three constant vectors
  if (tid < K) {
    float acc = s_b0[tid];
    for (int t = 0; t < NT0; ++t) acc = __fmaf_rn(pad, s_w0[t * K + tid], acc);
    s_k0[tid] = relu(acc);
  }
  __syncthreads();
  if (tid < K) {
    float acc[1] = {s_b1[tid]};
    for (int t = 0; t < NT; ++t) chain_tap<1>(acc, s_w1 + (t * K + tid) * KP, 0, s_k0);
    s_k1[tid] = relu(acc[0]);
  }
  __syncthreads();
  if (tid < K) {
    float acc[1] = {s_b2[tid]};
    for (int t = 0; t < NT; ++t) chain_tap<1>(acc, s_w2 + (t * K + tid) * KP, 0, s_k1);
    s_k2[tid] = __fadd_rn(acc[0], s_k0[tid]);
  }
  __syncthreads();

  // ---- volumes of this image
  const size_t q_plane = (size_t)(H + 8) * (W + 8);
  const size_t a0_plane = (size_t)(H + 6) * (W + 6) * K, a1_plane = (size_t)(H + 4) * (W + 4) * K,
               a2_plane = (size_t)(H + 2) * (W + 2) * K;
  float* q = p.q + (size_t)img * C * q_plane;
  float* a0 = p.a0 + (size_t)img * C * a0_plane;
  float* a1 = p.a1 + (size_t)img * C * a1_plane;
  float* a2 = p.a2 + (size_t)img * C * a2_plane;
  int* progress = p.progress + (size_t)img * C;
  int64_t* sym = p.sym + (size_t)img * C * H * W;
  auto q_at = [&](int d, int h, int w) { return q + (size_t)d * q_plane + (size_t)(h + 4) * (W + 8) + (w + 4); };
  auto a0_at = [&](int d, int h, int w) { return a0 + (size_t)d * a0_plane + ((size_t)(h + 3) * (W + 6) + (w + 3)) * K; };
  auto a1_at = [&](int d, int h, int w) { return a1 + (size_t)d * a1_plane + ((size_t)(h + 2) * (W + 4) + (w + 2)) * K; };
  auto a2_at = [&](int d, int h, int w) { return a2 + (size_t)d * a2_plane + ((size_t)(h + 1) * (W + 2) + (w + 1)) * K; };

  const uint8_t* my_stream = p.bytes + ((size_t)img * p.nstreams + stream) * p.cap;
  const int64_t my_len = p.decode ? p.sizes[(size_t)img * p.nstreams + stream] : 0;
  RcEnc enc;
  uint32_t dec_code = 0, dec_range = 0xFFFFFFFFu;
  int64_t dec_pos = 0;
  if (tid == BULK) {
    if (p.decode) {
      for (int k = 0; k < 4; ++k) {
        dec_code = (dec_code << 8) | (dec_pos < my_len ? my_stream[dec_pos] : 0);
        ++dec_pos;
      }
      *s_pos = dec_pos;
    } else {
      enc.init(p.bytes + ((size_t)img * p.nstreams + stream) * p.cap, p.cap);
    }
  }
  __syncthreads();  // s_pos is read by the bulk warps

  const bool is_coder = tid >= BULK;           // warp 9; only its first lane works
  const int j = tid / 8, c3 = tid % 8;         // bulk thread: position j of the step, channels c3, c3 + 8, c3 + 16
  const bool lane_ok = tid < MAXPOS * 8;       // 264 of the 288 bulk threads own outputs
  const int u_min = 5 * -3 - 3, u_max = 5 * (H + 2) + (W + 2);
  float p0[3] = {0.f, 0.f, 0.f}, p1[3] = {0.f, 0.f, 0.f}, p2[3] = {0.f, 0.f, 0.f}, p3[1] = {0.f};

  // positions of step u: h in [h_lo, h_lo + npos), w = u - 5 h in [-3, W + 2]
  auto positions = [&](int u, int& h_lo, int& npos) {
    int lo = u - (W + 2);
    lo = lo > 0 ? (lo + 4) / 5 : -((-lo) / 5);        // ceil
    if (lo < -3) lo = -3;
    int hi = u + 3;
    hi = hi >= 0 ? hi / 5 : -((-hi + 4) / 5);         // floor
    if (hi > H + 2) hi = H + 2;
    h_lo = lo;
    npos = hi - lo + 1;
  };
  // old taps (0..11) of every position of the step into shared memory: one L2 round trip per layer
  auto gather_old = [&](auto at_fn, const float* kconst, int d, int u, int h_lo, int npos, int lo, int hiH, int hiW) {
    const float4* k4 = reinterpret_cast<const float4*>(kconst);
    float4* dst = reinterpret_cast<float4*>(s_in);
    for (int idx = tid; idx < npos * 12 * (K / 4); idx += BULK) {
      const int jj = idx / (12 * (K / 4)), r = idx % (12 * (K / 4));
      const int t = r / (K / 4), g = r % (K / 4);
      const int hh = h_lo + jj, ww = u - 5 * hh;
      if (hh < lo || hh >= hiH || ww < lo || ww >= hiW) continue;
      const int dd = t < 9 ? d - 1 : d;               // taps 0..8: previous slice; 9..11: row h - 1
      const int dh = t < 9 ? t / 3 - 1 : -1;
      const int dw = t < 9 ? t % 3 - 1 : t - 10;
      dst[(jj * NT + t) * (K / 4) + g] =
          dd < 0 ? k4[g] : __ldcg(reinterpret_cast<const float4*>(at_fn(dd, hh + dh, ww + dw)) + g);
    }
  };

  // ---- bulk warps: partial chains (taps 0..11) of step u for all four layers
  auto bulk_partials = [&](int d, int u) {
    int h_lo, npos;
    positions(u, h_lo, npos);
    if (tid == 0 && d > 0) {  // slice d - 1 must be complete through step u + 6
      const int need = u + 6 < u_max ? u + 6 : u_max;
      while (ld_acquire(progress + d - 1) < need) __nanosleep(32);
    }
    bar_bulk();
    const int h = h_lo + j, w = u - 5 * h;
    const bool act = lane_ok && j < npos;
    if (act) {
#pragma unroll
      for (int c = 0; c < 3; ++c) p0[c] = s_b0[c3 + 8 * c];
#pragma unroll
      for (int t = 0; t < 12; ++t) {
        const int dd = d + c_taps[t][0];
        const float x = dd < 0 ? pad : __ldcg(q_at(dd, h + c_taps[t][1], w + c_taps[t][2]));
#pragma unroll
        for (int c = 0; c < 3; ++c) p0[c] = __fmaf_rn(x, s_w0[t * K + c3 + 8 * c], p0[c]);
      }
    }
    gather_old(a0_at, s_k0, d, u, h_lo, npos, -2, H + 2, W + 2);
    bar_bulk();
    if (act && h >= -2 && h < H + 2 && w >= -2 && w < W + 2) {
#pragma unroll
      for (int c = 0; c < 3; ++c) p1[c] = s_b1[c3 + 8 * c];
#pragma unroll
      for (int t = 0; t < 12; ++t) chain_tap<3>(p1, s_w1 + (t * K + c3) * KP, 8 * KP, s_in + (j * NT + t) * K);
    }
    bar_bulk();
    gather_old(a1_at, s_k1, d, u, h_lo, npos, -1, H + 1, W + 1);
    bar_bulk();
    if (act && h >= -1 && h < H + 1 && w >= -1 && w < W + 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) p2[c] = s_b2[c3 + 8 * c];
#pragma unroll
      for (int t = 0; t < 12; ++t) chain_tap<3>(p2, s_w2 + (t * K + c3) * KP, 8 * KP, s_in + (j * NT + t) * K);
    }
    bar_bulk();
    gather_old(a2_at, s_k2, d, u, h_lo, npos, 0, H, W);
    bar_bulk();
    if (act && c3 < L && h >= 0 && h < H && w >= 0 && w < W) {
      p3[0] = s_b3[c3];
#pragma unroll
      for (int t = 0; t < 12; ++t) chain_tap<1>(p3, s_w3 + (t * MAXL + c3) * KP, 0, s_in + (j * NT + t) * K);
    }
  };

  // ---- bulk warps: finish the chains of step u (taps 12, 13), store the activations, build the frequency tables
  auto finish = [&](int d, int u) {
    int h_lo, npos;
    positions(u, h_lo, npos);
    const int h = h_lo + j, w = u - 5 * h;
    const bool act = lane_ok && j < npos;
    const int par = u & 1, prev = par ^ 1;
    float* c0 = s_c0 + (par * MAXH + h + 3) * K;
    float* c1 = s_c1 + (par * MAXH + h + 3) * K;
    float* c2 = s_c2 + (par * MAXH + h + 3) * K;
    if (act) {  // layer 0: tap 12 = the centre decoded one step ago in this row (or padding)
      const float x = (h >= 0 && h < H && w - 1 >= 0 && w - 1 < W) ? s_qc[h] : pad;
      float* o = a0_at(d, h, w);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = relu(__fmaf_rn(x, s_w0[12 * K + c3 + 8 * c], p0[c]));
        o[c3 + 8 * c] = v;
        c0[c3 + 8 * c] = v;
      }
    }
    bar_bulk();
    if (act && h >= -2 && h < H + 2 && w >= -2 && w < W + 2) {
      chain_tap<3>(p1, s_w1 + (12 * K + c3) * KP, 8 * KP, s_c0 + (prev * MAXH + h + 3) * K);
      chain_tap<3>(p1, s_w1 + (13 * K + c3) * KP, 8 * KP, c0);
      float* o = a1_at(d, h, w);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = relu(p1[c]);
        o[c3 + 8 * c] = v;
        c1[c3 + 8 * c] = v;
      }
    }
    bar_bulk();
    if (act && h >= -1 && h < H + 1 && w >= -1 && w < W + 1) {
      chain_tap<3>(p2, s_w2 + (12 * K + c3) * KP, 8 * KP, s_c1 + (prev * MAXH + h + 3) * K);
      chain_tap<3>(p2, s_w2 + (13 * K + c3) * KP, 8 * KP, c1);
      float* o = a2_at(d, h, w);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = __fadd_rn(p2[c], c0[c3 + 8 * c]);  // + skip (layer-0 value of this position)
        o[c3 + 8 * c] = v;
        c2[c3 + 8 * c] = v;
      }
    }
    bar_bulk();
    if (act && c3 < L && h >= 0 && h < H && w >= 0 && w < W) {
      chain_tap<1>(p3, s_w3 + (12 * MAXL + c3) * KP, 0, s_c2 + (prev * MAXH + h + 3) * K);
      chain_tap<1>(p3, s_w3 + (13 * MAXL + c3) * KP, 0, c2);
      s_logit[j * MAXL + c3] = relu(p3[0]);
    }
    bar_bulk();
    // frequency tables (cumulative), one thread per position; other threads stage the coder's global reads
    if (tid < npos) {
      const int hf = h_lo + tid, wf = u - 5 * hf;
      if (hf >= 0 && hf < H && wf >= 0 && wf < W) {
        uint32_t f[MAXL];
        logits_to_freqs(s_logit + tid * MAXL, L, f);
        uint32_t c = 0;  // entry i = sum of f[0..i-1]; the last boundary (65536) is implied
        for (int i = 0; i < L; ++i) {
          s_freq[tid * MAXL + i] = c;
          c += f[i];
        }
        if (!p.decode) s_sym[tid] = (int)sym[((size_t)d * H + hf) * W + wf];
      }
    } else if (p.decode && tid >= 64 && tid < 64 + WIN) {
      const int64_t at = (int64_t)*s_pos + (tid - 64);
      s_win[tid - 64] = at < my_len ? my_stream[at] : 0;
    }
  };

  // ---- coder thread: the symbols of step u in increasing h.  Decoding is division-free: the symbol is the number of
  // cumulative thresholds r * cum[i] the code value has reached (floor(code / r) >= cum[i]  <=>  code >= r * cum[i]).
  auto coder = [&](int d, int u) {
    int h_lo, npos;
    positions(u, h_lo, npos);
    int wofs = 0;                      // read offset into the staged window (s_win = stream[dec_pos ...])
    for (int jj = 0; jj < npos; ++jj) {
      const int hc = h_lo + jj, wc = u - 5 * hc;
      if (hc < 0 || hc >= H || wc < 0 || wc >= W) continue;
      const uint32_t* cumt = s_freq + jj * MAXL;
      int sy;
      if (p.decode) {
        const uint32_t r = dec_range >> TOTAL_BITS;
        uint32_t lo = 0;
        sy = 0;
#pragma unroll
        for (int i = 1; i < MAXL; ++i)  // thresholds ascend: the last one reached is the symbol
          if (i < L) {
            const uint32_t ci = cumt[i];
            if (dec_code >= r * ci) { sy = i; lo = ci; }  // r < 2^16 and ci < 2^16: the product fits 32 bits
          }
        const uint32_t hi = sy + 1 < L ? cumt[sy + 1] : TOTAL;
        dec_code -= lo * r;
        dec_range = r * (hi - lo);
        while (dec_range < (1u << 24)) {
          const int64_t at = dec_pos + wofs;
          const uint32_t byte = wofs < WIN ? s_win[wofs] : (at < my_len ? my_stream[at] : 0);
          ++wofs;
          dec_code = (dec_code << 8) | byte;
          dec_range <<= 8;
        }
        sym[((size_t)d * H + hc) * W + wc] = sy;
        *q_at(d, hc, wc) = s_cent[sy];
      } else {
        sy = s_sym[jj];
        const uint32_t lo = cumt[sy], hi = sy + 1 < L ? cumt[sy + 1] : TOTAL;
        enc.encode(lo, hi - lo);
      }
      s_qc[hc] = s_cent[sy];
    }
    if (p.decode) {
      dec_pos += wofs;
      *s_pos = dec_pos;
    }
  };

  for (int d = stream; d < C; d += p.nstreams) {
    if (!is_coder) {  // fill the pipeline: the first step of the slice has no coder work before it
      bulk_partials(d, u_min);
      bar_bulk();
      finish(d, u_min);
    }
    for (int u = u_min; u <= u_max; ++u) {
      __syncthreads();  // tables of step u are ready; everything of earlier steps is written
      if (is_coder) {
        if (tid == BULK) coder(d, u);
      } else if (u < u_max) {
        bulk_partials(d, u + 1);
      }
      __syncthreads();  // symbols of step u are coded; partial chains of step u + 1 sit in registers
      if (tid == 32) st_release(progress + d, u);  // step u is complete: publish it off the critical path
      if (!is_coder && u < u_max) finish(d, u + 1);
    }
  }
  if (tid == BULK && !p.decode) {
    enc.flush();
    p.sizes[(size_t)img * p.nstreams + stream] = enc.pos;
    if (enc.overflow) atomicExch(p.status, 1);
  }
}

// q volume with its halo: centres of the symbols (encode) or the pad value everywhere (decode); progress reset
__global__ void pc_codec_prepare_kernel(CodecArgs p) {
  const size_t plane = (size_t)(p.H + 8) * (p.W + 8);
  const size_t total = (size_t)p.n * p.C * plane;
  const float pad = p.centers[0];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % (p.W + 8)) - 4, h = (int)((i / (p.W + 8)) % (p.H + 8)) - 4;
    const size_t nd = i / plane;
    float v = pad;
    if (!p.decode && h >= 0 && h < p.H && w >= 0 && w < p.W) {
      const int64_t s = p.sym[(nd * p.H + h) * p.W + w];
      v = p.centers[s >= 0 && s < p.L ? s : 0];
    }
    p.q[i] = v;
  }
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid < (size_t)p.n * p.C) p.progress[gid] = PROGRESS_INIT;
  if (gid == 0 && p.reset_status) *p.status = 0;
}

// ---------------------------------------------------------------- fast encoder (all symbols known)
__global__ void pc_qhard_kernel(const int64_t* __restrict__ sym, const float* __restrict__ centers, int L,
                                float* __restrict__ q, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = sym[i];
    q[i] = centers[s >= 0 && s < L ? s : 0];
  }
}

// stage 2 of the fast encoder: one CTA per stream walks its depth slices in coding order (u = 5 h + w, then h);
// all threads fetch the (cum, freq) words of UB wavefront steps into shared memory, one thread range-codes them.
constexpr int UB = 32;
__global__ void __launch_bounds__(256) pc_symbol_coder_kernel(const uint32_t* __restrict__ packed, int C, int H, int W,
                                                              int nstreams, uint8_t* bytes, int64_t cap, int64_t* sizes,
                                                              int* status) {
  __shared__ uint32_t s_tab[UB * 32];
  const int img = blockIdx.x / nstreams, stream = blockIdx.x % nstreams;
  const uint32_t* tab = packed + (size_t)img * C * H * W;
  RcEnc enc;
  if (threadIdx.x == 0) enc.init(bytes + ((size_t)img * nstreams + stream) * cap, cap);
  const int u_last = 5 * (H - 1) + (W - 1);
  for (int d = stream; d < C; d += nstreams) {
    for (int u0 = 0; u0 <= u_last; u0 += UB) {
      for (int i = threadIdx.x; i < UB * 32; i += blockDim.x) {
        const int u = u0 + i / 32;
        int h_lo = u - (W - 1);
        h_lo = h_lo > 0 ? (h_lo + 4) / 5 : 0;
        const int h = h_lo + i % 32, w = u - 5 * h;
        s_tab[i] = (u <= u_last && h < H && w >= 0 && w < W) ? tab[((size_t)d * H + h) * W + w] : 0u;
      }
      __syncthreads();
      if (threadIdx.x == 0)
        for (int i = 0; i < UB * 32; ++i) {
          const uint32_t e = s_tab[i];
          if (e != 0u) enc.encode(e >> 16, e & 0xFFFFu);  // every frequency is >= 1, so 0 marks "no symbol here"
        }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    enc.flush();
    sizes[(size_t)img * nstreams + stream] = enc.pos;
    if (enc.overflow) atomicExch(status, 1);
  }
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct Layout {
  size_t q, a0, a1, a2, progress, total;
};
Layout layout(int n, int c, int h, int w) {
  Layout l;
  size_t off = 0;
  l.q = off; off += align256((size_t)n * c * (h + 8) * (w + 8) * sizeof(float));
  l.a0 = off; off += align256((size_t)n * c * (h + 6) * (w + 6) * K * sizeof(float));
  l.a1 = off; off += align256((size_t)n * c * (h + 4) * (w + 4) * K * sizeof(float));
  l.a2 = off; off += align256((size_t)n * c * (h + 2) * (w + 2) * K * sizeof(float));
  l.progress = off; off += align256((size_t)n * c * sizeof(int));
  l.total = off;
  return l;
}

constexpr size_t kSmemBytes =
    (size_t)(2 * NT * K * KP + NT * MAXL * KP + NT0 * K + 3 * K + MAXL + 3 * K + MAXL + MAXPOS * MAXL) * sizeof(float) +
    (size_t)MAXPOS * MAXL * sizeof(uint32_t) + (size_t)MAXPOS * NT * K * sizeof(float) +
    (size_t)(3 * 2 * MAXH * K + MAXH) * sizeof(float) + (size_t)(MAXPOS + 1) * sizeof(int) + sizeof(long long) + WIN;

int run_codec(dsin_handle_t h, int decode, bool wavefront_encode, int64_t* symbols, int n, int c, int hh, int ww, const float* centers, int L,
              const float* const* wb, int k, int nstreams, uint8_t* bytes, int64_t cap, int64_t* sizes, int* status_out,
              void* workspace, cudaStream_t st) {
  DSIN_REQUIRE(h, symbols && centers && wb && bytes && sizes && workspace && status_out, "null pointer");
  DSIN_REQUIRE(h, n > 0 && c > 0 && hh > 0 && ww > 0 && cap > 0, "bad shape");
  DSIN_REQUIRE(h, k == K, "the codec is built for 24 hidden channels");
  DSIN_REQUIRE(h, L >= 2 && L <= MAXL, "2..8 centres");
  DSIN_REQUIRE(h, nstreams >= 1 && nstreams <= 64, "1..64 streams per image");
  DSIN_REQUIRE(h, (ww + 5) / 5 + 1 <= MAXPOS, "volume wider than 159 symbols (one CTA walks a slice 33 positions at a time)");
  DSIN_REQUIRE(h, hh + 6 <= MAXH, "volume taller than 58 symbols (rows of the shared-memory rings)");
  if (!decode && !wavefront_encode && L == 6 && (ww + 4) / 5 + 1 <= 32) {
    // every symbol is known: frequency tables for the whole volume in parallel (probclass.cu), then one CTA per
    // stream range-codes its symbols.  Same bytes as the wavefront kernel in encode mode.
    const size_t nsym = (size_t)n * c * hh * ww;
    uint8_t* ws = (uint8_t*)workspace;
    float* qh = (float*)ws;
    uint32_t* packed = (uint32_t*)(ws + align256(nsym * sizeof(float)));
    void* tws = ws + 2 * align256(nsym * sizeof(float));
    if (cudaMemsetAsync(status_out, 0, sizeof(int), st) != cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: memset failed", __func__);
    pc_qhard_kernel<<<h->sm_count * 4, 256, 0, st>>>(symbols, centers, L, qh, (int64_t)nsym);
    DSIN_LAUNCHED(h);
    int rc = pc1_symbol_tables(h, qh, symbols, n, c, hh, ww, centers, L, wb, packed, tws, st);
    if (rc != DSIN_OK) return rc;
    pc_symbol_coder_kernel<<<n * nstreams, 256, 0, st>>>(packed, c, hh, ww, nstreams, bytes, cap, sizes, status_out);
    DSIN_LAUNCHED(h);
    return DSIN_OK;
  }
  static bool configured[DSIN_MAX_DEVICES] = {};  // cudaFuncSetAttribute is per device
  if (!configured[h->device]) {
    if (cudaFuncSetAttribute(pc_codec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: cannot raise dynamic shared memory", __func__);
    configured[h->device] = true;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pc_codec_kernel, THREADS, kSmemBytes) != cudaSuccess || per_sm < 1)
    return dsin_fail(h, DSIN_ERR_CUDA, "%s: kernel does not fit an SM", __func__);
  const int resident = per_sm * h->sm_count;
  DSIN_REQUIRE(h, nstreams <= resident, "more streams than co-resident CTAs");
  const int imgs_per_launch = resident / nstreams;
  for (int i0 = 0; i0 < n; i0 += imgs_per_launch) {
    const int ni = n - i0 < imgs_per_launch ? n - i0 : imgs_per_launch;
    const Layout l = layout(ni, c, hh, ww);
    CodecArgs a;
    memset(&a, 0, sizeof(a));
    a.n = ni; a.C = c; a.H = hh; a.W = ww; a.L = L; a.nstreams = nstreams; a.decode = decode;
    a.centers = centers;
    a.w0 = wb[0]; a.b0 = wb[1]; a.w1 = wb[2]; a.b1 = wb[3]; a.w2 = wb[4]; a.b2 = wb[5]; a.w3 = wb[6]; a.b3 = wb[7];
    uint8_t* ws = (uint8_t*)workspace;
    a.q = (float*)(ws + l.q); a.a0 = (float*)(ws + l.a0); a.a1 = (float*)(ws + l.a1); a.a2 = (float*)(ws + l.a2);
    a.progress = (int*)(ws + l.progress); a.status = status_out; a.reset_status = i0 == 0;
    a.sym = symbols + (size_t)i0 * c * hh * ww;
    a.bytes = bytes + (size_t)i0 * nstreams * cap; a.cap = cap; a.sizes = sizes + (size_t)i0 * nstreams;
    pc_codec_prepare_kernel<<<h->sm_count * 4, 256, 0, st>>>(a);
    DSIN_LAUNCHED(h);
    void* params[] = {&a};
    if (cudaLaunchCooperativeKernel((const void*)pc_codec_kernel, dim3(ni * nstreams), dim3(THREADS), params, kSmemBytes,
                                    st) != cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: cooperative launch failed (are all CTAs co-resident?)", __func__);
    DSIN_LAUNCHED(h);
  }
  return DSIN_OK;
}

}  // namespace

extern "C" {

int64_t dsin_pc_codec_workspace_bytes(int n, int c, int hh, int ww) {
  if (n <= 0 || c <= 0 || hh <= 0 || ww <= 0) return -1;
  const int64_t wavefront = (int64_t)layout(n, c, hh, ww).total;
  const int64_t fast = 2 * (int64_t)align256((size_t)n * c * hh * ww * sizeof(float)) +
                       pc1_symbol_tables_workspace(n, c, hh, ww) + 256;
  return wavefront > fast ? wavefront : fast;
}

int dsin_pc_encode(dsin_handle_t h, const int64_t* symbols, int n, int c, int hh, int ww, const float* centers, int L,
                   const float* const* weights, int k, int nstreams, uint8_t* bytes, int64_t cap, int64_t* sizes,
                   int* status, void* workspace, void* stream) {
  return run_codec(h, 0, false, const_cast<int64_t*>(symbols), n, c, hh, ww, centers, L, weights, k, nstreams, bytes, cap, sizes,
                   status, workspace, (cudaStream_t)stream);
}

int dsin_pc_encode_wavefront(dsin_handle_t h, const int64_t* symbols, int n, int c, int hh, int ww, const float* centers,
                             int L, const float* const* weights, int k, int nstreams, uint8_t* bytes, int64_t cap,
                             int64_t* sizes, int* status, void* workspace, void* stream) {
  return run_codec(h, 0, true, const_cast<int64_t*>(symbols), n, c, hh, ww, centers, L, weights, k, nstreams, bytes, cap,
                   sizes, status, workspace, (cudaStream_t)stream);
}

int dsin_pc_decode(dsin_handle_t h, const uint8_t* bytes, int64_t cap, const int64_t* sizes, int n, int c, int hh, int ww,
                   const float* centers, int L, const float* const* weights, int k, int nstreams, int64_t* symbols,
                   int* status, void* workspace, void* stream) {
  return run_codec(h, 1, false, symbols, n, c, hh, ww, centers, L, weights, k, nstreams, const_cast<uint8_t*>(bytes), cap,
                   const_cast<int64_t*>(sizes), status, workspace, (cudaStream_t)stream);
}

}  // extern "C"
