/* CPU oracle of the DSIN "PC1" entropy coder.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu legs).
 *
 * What it restates.  The reference ships no entropy coder, only the pieces one would be built from
 * (/root/reference/src/probclass_imgcomp.py:361-482: symbols are coded one at a time, each with the
 * probabilities the context model predicts from the already-coded symbols -- PredictionNetwork.get_pr /
 * get_freqs, `freqs = int64(pr * resolution)`, `max(freqs, 1)` -- in the causal order of the 3-D masked
 * convolutions, probclass_imgcomp.py:150-176,214-261,268-292).  This file fixes everything the reference
 * leaves open, bit for bit, so that an encoder and a decoder on different hardware agree:
 *
 *   network   the four masked 3-D convolutions of _ResShallow (probclass_imgcomp.py:214-221) on the volume of
 *             quantiser centres, padded with centres[0] (auto_pad_value, :59-61; pad_for_probclass3d, :268-292),
 *             evaluated in fp32 with ONE fixed operation order: acc = bias; for live taps in (kd, kh, kw)
 *             raster order; for ci ascending: acc = fmaf(x, w, acc).  ReLU = (x > 0 ? x : 0).
 *   freqs     m = max logit; e_i = exp_det(l_i - m) (a fixed fmaf polynomial, below); Z = left-to-right sum;
 *             f_i = 1 + trunc(e_i * (65530 / Z)); the remainder to 65536 goes to the first arg-max symbol.
 *   coder     carry-propagating 32-bit range coder (64-bit low, byte renormalisation), 16-bit frequencies.  The
 *             coder's first output byte (its initial cache, always 0) is not stored; the final flush stores the
 *             value of [low, low+range) with the most trailing zero bytes and omits those (a decoder reads
 *             zeros past the end), so a stream costs at most two bytes more than its code length.
 *   order     depth slice d (= bottleneck channel) goes to stream d % nstreams; within a stream the symbols are
 *             coded slice by slice, inside a slice by wavefront time u = 5*h + w, then by h.  (25*d + 5*h + w is
 *             a valid parallel schedule of the causal context, which is what the CUDA decoder exploits.)
 *
 * PARITY UNPINNED: there is nothing in the reference to pin a bitstream against.  The CUDA codec must produce
 * these exact bytes; this file's decoder must invert its encoder; the code length must match the cross entropy
 * of oracle/dsin_oracle.py's probclass_bitcost to within the frequency quantisation.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libpc_codec.so oracle/pc_codec.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PC1_TOTAL_BITS 16
#define PC1_TOTAL (1u << PC1_TOTAL_BITS)
#define PC1_MAXL 16

typedef struct {
  int C, H, W;      /* symbol volume (depth = bottleneck channels) */
  int L, K;         /* centres, hidden channels (24) */
  const float* centers;
  float pad;        /* centres[0] */
  /* live-tap-major weights: w0[13][1][K], w1[14][K][K], w2[14][K][K], w3[14][K][L]; biases b0..b3 */
  const float *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3;
} pc1_model;

/* live taps as position offsets (dd, dh, dw): previous slice 3x3 in raster order, then the causal part of the
 * current slice; the first layer excludes the centre (probclass_imgcomp.py:150-176) */
static const int TAPS[14][3] = {{-1, -1, -1}, {-1, -1, 0}, {-1, -1, 1}, {-1, 0, -1}, {-1, 0, 0}, {-1, 0, 1}, {-1, 1, -1},
                                {-1, 1, 0},   {-1, 1, 1},  {0, -1, -1}, {0, -1, 0}, {0, -1, 1}, {0, 0, -1}, {0, 0, 0}};

static float relu(float x) { return x > 0.0f ? x : 0.0f; }

static float exp_det(float x) { /* x <= 0; only +, *, fma, floor and an exact power-of-two scaling */
  if (x < -80.0f) x = -80.0f;
  float t = x * 1.4426950408889634f;
  float n = floorf(t);
  float f = t - n;
  float p = 1.5353362e-4f;
  p = fmaf(p, f, 1.3398874e-3f);
  p = fmaf(p, f, 9.6184370e-3f);
  p = fmaf(p, f, 5.5503324e-2f);
  p = fmaf(p, f, 2.4022648e-1f);
  p = fmaf(p, f, 6.9314720e-1f);
  p = fmaf(p, f, 1.0f);
  union { uint32_t u; float v; } s;
  s.u = (uint32_t)((int)n + 127) << 23;
  return p * s.v;
}

static void logits_to_freqs(const float* l, int L, uint32_t* f) {
  float m = l[0];
  int am = 0;
  for (int i = 1; i < L; ++i)
    if (l[i] > m) { m = l[i]; am = i; }
  float e[PC1_MAXL], Z = 0.0f;
  for (int i = 0; i < L; ++i) { e[i] = exp_det(l[i] - m); Z = Z + e[i]; }
  float scale = (float)(PC1_TOTAL - (unsigned)L) / Z;
  uint32_t sum = 0;
  for (int i = 0; i < L; ++i) { f[i] = 1u + (uint32_t)(e[i] * scale); sum += f[i]; }
  f[am] += PC1_TOTAL - sum;
}

/* ---------------------------------------------------------------- range coder */
typedef struct {
  uint64_t low;
  uint32_t range;
  uint8_t cache;
  uint64_t cache_size;
  uint8_t* out;
  int64_t pos, cap;
  int overflow, skip_first;
} rc_enc;

static void rc_put(rc_enc* e, uint8_t b) {
  if (e->skip_first) { e->skip_first = 0; return; }
  if (e->pos < e->cap) e->out[e->pos] = b; else e->overflow = 1;
  e->pos++;
}
static void rc_enc_init(rc_enc* e, uint8_t* out, int64_t cap) {
  e->low = 0; e->range = 0xFFFFFFFFu; e->cache = 0; e->cache_size = 1; e->out = out; e->pos = 0; e->cap = cap;
  e->overflow = 0; e->skip_first = 1;
}
static void rc_shift_low(rc_enc* e) {
  if ((uint32_t)e->low < 0xFF000000u || (e->low >> 32) != 0) {
    uint8_t carry = (uint8_t)(e->low >> 32);
    uint8_t c = e->cache;
    do {
      rc_put(e, (uint8_t)(c + carry));
      c = 0xFF;
    } while (--e->cache_size != 0);
    e->cache = (uint8_t)((e->low >> 24) & 0xFF);
  }
  e->cache_size++;
  e->low = (e->low & 0x00FFFFFFu) << 8;
}
static void rc_encode(rc_enc* e, uint32_t cum, uint32_t freq) {
  uint32_t r = e->range >> PC1_TOTAL_BITS;
  e->low += (uint64_t)r * cum;
  e->range = r * freq;
  while (e->range < (1u << 24)) { e->range <<= 8; rc_shift_low(e); }
}
static void rc_flush(rc_enc* e) {
  /* range >= 2^24 here, so [low, low+range) holds a multiple of 2^24: at least three trailing zero bytes */
  const uint64_t hi = e->low + e->range - 1;
  int k = 4;
  uint64_t v = 0;
  for (; k >= 0; --k) {
    v = hi & ~((1ull << (8 * k)) - 1);
    if (v >= e->low) break;
  }
  e->low = v;
  for (int i = 0; i < 5 - k; ++i) rc_shift_low(e);
}

typedef struct {
  uint32_t code, range;
  const uint8_t* in;
  int64_t pos, len;
} rc_dec;
static uint8_t rc_get(rc_dec* d) { return d->pos < d->len ? d->in[d->pos++] : (d->pos++, 0); }
static void rc_dec_init(rc_dec* d, const uint8_t* in, int64_t len) {
  d->in = in; d->pos = 0; d->len = len; d->range = 0xFFFFFFFFu; d->code = 0;
  for (int i = 0; i < 4; ++i) d->code = (d->code << 8) | rc_get(d);
}
static int rc_decode(rc_dec* d, const uint32_t* f, int L) {
  uint32_t r = d->range >> PC1_TOTAL_BITS;
  uint32_t v = d->code / r;
  if (v > PC1_TOTAL - 1) v = PC1_TOTAL - 1;
  uint32_t cum = 0;
  int s = 0;
  while (s < L - 1 && cum + f[s] <= v) { cum += f[s]; ++s; }
  d->code -= cum * r;
  d->range = r * f[s];
  while (d->range < (1u << 24)) { d->code = (d->code << 8) | rc_get(d); d->range <<= 8; }
  return s;
}

/* ---------------------------------------------------------------- the context model on extended domains */
typedef struct {
  const pc1_model* m;
  float *q, *a0, *a1, *a2; /* q: [C+4][H+8][W+8]; a0: [C+3][H+6][W+6][K]; a1: [C+2][H+4][W+4][K]; a2: [C+1][H+2][W+2][K] */
} pc1_state;

static inline float* at(float* base, int halo, int d, int h, int w, int H, int W, int K) {
  /* volume with `halo` extra planes in front depth and `halo` rows/cols on each side */
  return base + ((((size_t)(d + halo) * (size_t)(H + 2 * halo)) + (size_t)(h + halo)) * (size_t)(W + 2 * halo) +
                 (size_t)(w + halo)) * (size_t)K;
}

static void layer0(pc1_state* s, int d, int h, int w) {
  const pc1_model* m = s->m;
  float* o = at(s->a0, 3, d, h, w, m->H, m->W, m->K);
  for (int co = 0; co < m->K; ++co) {
    float acc = m->b0[co];
    for (int t = 0; t < 13; ++t) {
      float x = *at(s->q, 4, d + TAPS[t][0], h + TAPS[t][1], w + TAPS[t][2], m->H, m->W, 1);
      acc = fmaf(x, m->w0[t * m->K + co], acc);
    }
    o[co] = relu(acc);
  }
}
static void layerK(const pc1_model* m, const float* wt, const float* b, float* in, int in_halo, int d, int h, int w,
                   int cout, float* acc_out) {
  for (int co = 0; co < cout; ++co) {
    float acc = b[co];
    for (int t = 0; t < 14; ++t) {
      const float* x = at(in, in_halo, d + TAPS[t][0], h + TAPS[t][1], w + TAPS[t][2], m->H, m->W, m->K);
      const float* ww = wt + (size_t)t * m->K * cout;
      for (int ci = 0; ci < m->K; ++ci) acc = fmaf(x[ci], ww[(size_t)ci * cout + co], acc);
    }
    acc_out[co] = acc;
  }
}
static void layer1(pc1_state* s, int d, int h, int w) {
  const pc1_model* m = s->m;
  float acc[64];
  layerK(m, m->w1, m->b1, s->a0, 3, d, h, w, m->K, acc);
  float* o = at(s->a1, 2, d, h, w, m->H, m->W, m->K);
  for (int co = 0; co < m->K; ++co) o[co] = relu(acc[co]);
}
static void layer2(pc1_state* s, int d, int h, int w) {
  const pc1_model* m = s->m;
  float acc[64];
  layerK(m, m->w2, m->b2, s->a1, 2, d, h, w, m->K, acc);
  float* o = at(s->a2, 1, d, h, w, m->H, m->W, m->K);
  const float* skip = at(s->a0, 3, d, h, w, m->H, m->W, m->K); /* net + res_in[:, 2:, 2:-2, 2:-2] */
  for (int co = 0; co < m->K; ++co) o[co] = acc[co] + skip[co];
}
static void layer3(pc1_state* s, int d, int h, int w, float* logits) {
  const pc1_model* m = s->m;
  float acc[PC1_MAXL];
  layerK(m, m->w3, m->b3, s->a2, 1, d, h, w, m->L, acc);
  for (int i = 0; i < m->L; ++i) logits[i] = relu(acc[i]); /* the logits layer keeps slim's default ReLU */
}

static int in_dom(const pc1_model* m, int halo, int h, int w) {
  return h >= -halo && h < m->H + halo && w >= -halo && w < m->W + halo;
}

/* mode 0 = encode (sym in, streams out), 1 = decode (streams in, sym out) */
static int run(const pc1_model* m, int mode, int32_t* sym, int nstreams, uint8_t* bytes, int64_t cap, int64_t* sizes,
               double* ideal_bits) {
  if (m->L > PC1_MAXL || m->K > 64 || nstreams < 1 || nstreams > 64) return -1;
  const int C = m->C, H = m->H, W = m->W, K = m->K;
  pc1_state s;
  s.m = m;
  size_t nq = (size_t)(C + 4) * (H + 8) * (W + 8);
  s.q = (float*)malloc(nq * sizeof(float));
  s.a0 = (float*)calloc((size_t)(C + 3) * (H + 6) * (W + 6) * K, sizeof(float));
  s.a1 = (float*)calloc((size_t)(C + 2) * (H + 4) * (W + 4) * K, sizeof(float));
  s.a2 = (float*)calloc((size_t)(C + 1) * (H + 2) * (W + 2) * K, sizeof(float));
  rc_enc* enc = (rc_enc*)calloc(nstreams, sizeof(rc_enc));
  rc_dec* dec = (rc_dec*)calloc(nstreams, sizeof(rc_dec));
  if (!s.q || !s.a0 || !s.a1 || !s.a2 || !enc || !dec) return -2;
  for (size_t i = 0; i < nq; ++i) s.q[i] = m->pad;
  /* q rows/planes outside the volume above are addressed through at(q, 4, ...) with a depth halo of 4 in FRONT
   * only: the allocation has C+4 planes, plane index d+4 */
  for (int k = 0; k < nstreams; ++k) {
    if (mode == 0) rc_enc_init(&enc[k], bytes + (size_t)k * cap, cap);
    else rc_dec_init(&dec[k], bytes + (size_t)k * cap, sizes[k]);
  }
  double bits = 0.0;
  /* padding-only slices in front of the volume */
  for (int d = -3; d < 0; ++d) {
    for (int h = -3; h < H + 3; ++h)
      for (int w = -3; w < W + 3; ++w) layer0(&s, d, h, w);
    if (d >= -2)
      for (int h = -2; h < H + 2; ++h)
        for (int w = -2; w < W + 2; ++w) layer1(&s, d, h, w);
    if (d >= -1)
      for (int h = -1; h < H + 1; ++h)
        for (int w = -1; w < W + 1; ++w) layer2(&s, d, h, w);
  }
  for (int d = 0; d < C; ++d) {
    const int k = d % nstreams;
    for (int u = 5 * -3 - 3; u <= 5 * (H + 2) + (W + 2); ++u)
      for (int h = -3; h < H + 3; ++h) {
        const int w = u - 5 * h;
        if (w < -3 || w >= W + 3) continue;
        layer0(&s, d, h, w);
        if (in_dom(m, 2, h, w)) layer1(&s, d, h, w);
        if (in_dom(m, 1, h, w)) layer2(&s, d, h, w);
        if (in_dom(m, 0, h, w)) {
          float logits[PC1_MAXL];
          uint32_t f[PC1_MAXL];
          layer3(&s, d, h, w, logits);
          logits_to_freqs(logits, m->L, f);
          int32_t* sp = sym + ((size_t)d * H + h) * W + w;
          int sy;
          if (mode == 0) {
            sy = *sp;
            if (sy < 0 || sy >= m->L) return -3;
            uint32_t cum = 0;
            for (int i = 0; i < sy; ++i) cum += f[i];
            rc_encode(&enc[k], cum, f[sy]);
          } else {
            sy = rc_decode(&dec[k], f, m->L);
            *sp = sy;
          }
          bits -= log2((double)f[sy] / (double)PC1_TOTAL);
          *at(s.q, 4, d, h, w, H, W, 1) = m->centers[sy];
        }
      }
  }
  int rc = 0;
  if (mode == 0)
    for (int k = 0; k < nstreams; ++k) {
      rc_flush(&enc[k]);
      sizes[k] = enc[k].pos;
      if (enc[k].overflow) rc = -4;
    }
  if (ideal_bits) *ideal_bits = bits;
  free(s.q); free(s.a0); free(s.a1); free(s.a2); free(enc); free(dec);
  return rc;
}

int pc1_encode(const pc1_model* m, const int32_t* sym, int nstreams, uint8_t* bytes, int64_t cap, int64_t* sizes,
               double* ideal_bits) {
  return run(m, 0, (int32_t*)sym, nstreams, bytes, cap, sizes, ideal_bits);
}
int pc1_decode(const pc1_model* m, int32_t* sym, int nstreams, const uint8_t* bytes, int64_t cap, const int64_t* sizes,
               double* ideal_bits) {
  return run(m, 1, sym, nstreams, (uint8_t*)bytes, cap, (int64_t*)sizes, ideal_bits);
}
/* frequencies of one logits vector (unit tests) */
void pc1_freqs(const float* logits, int L, uint32_t* f) { logits_to_freqs(logits, L, f); }
float pc1_exp(float x) { return exp_det(x); }
/* range coder alone: encode n symbols with given per-symbol frequency tables (n x L, each summing to 65536),
 * decode them again; returns the number of mismatches (or < 0), *size = stream bytes */
int pc1_rc_selftest(const uint32_t* freqs, const int32_t* sym, int n, int L, uint8_t* buf, int64_t cap, int64_t* size) {
  rc_enc e;
  rc_enc_init(&e, buf, cap);
  for (int i = 0; i < n; ++i) {
    const uint32_t* f = freqs + (size_t)i * L;
    uint32_t cum = 0;
    for (int j = 0; j < sym[i]; ++j) cum += f[j];
    rc_encode(&e, cum, f[sym[i]]);
  }
  rc_flush(&e);
  if (e.overflow) return -4;
  *size = e.pos;
  rc_dec d;
  rc_dec_init(&d, buf, e.pos);
  int bad = 0;
  for (int i = 0; i < n; ++i) bad += rc_decode(&d, freqs + (size_t)i * L, L) != sym[i];
  return bad;
}
