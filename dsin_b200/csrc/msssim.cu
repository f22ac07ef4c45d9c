// K9: MS-SSIM (metric of record) on the device, float64 like the reference's numpy/scipy code.
//   replaces ms_ssim_np_imgcomp.MultiScaleSSIM / _SSIMForMultiScale / _FSpecialGauss
//   (src/ms_ssim_np_imgcomp.py:51-110,113-124,127-200) as called by utils.msssim_x_vs_rec
//   (src/utils.py:94-99).
// Images are (G groups, batch, height, width, depth) fp32; every group is reduced separately
// (one group = one image: standard form batch=1,(H,W,3); the reference's literal call form is
// batch=H, height=W, width=3, depth=1).  Per level: 'valid' correlation with the size x size Gaussian
// (size = min(11, height, width), sigma = size*1.5/11), SSIM and CS maps, means; then a 2x2 box
// average with 'reflect' boundary and stride 2.  Output: (G, 5, 2) doubles = mean ssim, mean cs.
#include "common.cuh"

#define MS_LEVELS 5

__global__ void ms_f32_to_f64_kernel(const float* __restrict__ a, double* __restrict__ o, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = (double)a[i];
}

__global__ void ms_down_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t nb, int H, int W,
                               int C) {
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nb * H2 * W2 * C) return;
  int c = (int)(idx % C);
  int64_t t = idx / C;
  int j = (int)(t % W2);
  t /= W2;
  int i = (int)(t % H2);
  int64_t b = t / H2;
  int i0 = 2 * i, i1 = min(2 * i + 1, H - 1), j0 = 2 * j, j1 = min(2 * j + 1, W - 1);  // 'reflect'
  const double* p = in + b * H * W * C;
  out[idx] = (p[((int64_t)i0 * W + j0) * C + c] + p[((int64_t)i0 * W + j1) * C + c] +
              p[((int64_t)i1 * W + j0) * C + c] + p[((int64_t)i1 * W + j1) * C + c]) *
             0.25;
}

__global__ void ms_ssim_level_kernel(const double* __restrict__ a, const double* __restrict__ b, int G, int batch,
                                     int H, int W, int C, int size, double sigma, double* __restrict__ sums) {
  __shared__ double s_w[121];
  __shared__ double s_red[2][256];
  if (threadIdx.x == 0) {
    // fspecial('gaussian'): grid offset 0.5 for even sizes
    const int radius = size / 2;
    const double off = (size % 2 == 0) ? 0.5 : 0.0;
    double tot = 0.0;
    for (int y = 0; y < size; ++y)
      for (int x = 0; x < size; ++x) {
        double yy = off - radius + y, xx = off - radius + x;
        double g = exp(-((xx * xx + yy * yy) / (2.0 * sigma * sigma)));
        s_w[y * size + x] = g;
        tot += g;
      }
    for (int k = 0; k < size * size; ++k) s_w[k] /= tot;
  }
  __syncthreads();
  const int Hv = H - size + 1, Wv = W - size + 1;
  const int g = blockIdx.y;
  const int64_t per_group = (int64_t)batch * Hv * Wv * C;
  double my_ssim = 0.0, my_cs = 0.0;
  const double c1 = (0.01 * 255.0) * (0.01 * 255.0), c2 = (0.03 * 255.0) * (0.03 * 255.0);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < per_group;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    int64_t t = idx / C;
    int j = (int)(t % Wv);
    t /= Wv;
    int i = (int)(t % Hv);
    int64_t bb = (int64_t)g * batch + t / Hv;
    const double* pa = a + bb * H * W * C;
    const double* pb = b + bb * H * W * C;
    double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int y = 0; y < size; ++y)
      for (int x = 0; x < size; ++x) {
        double w = s_w[y * size + x];
        double va = pa[((int64_t)(i + y) * W + j + x) * C + c], vb = pb[((int64_t)(i + y) * W + j + x) * C + c];
        m1 += w * va;
        m2 += w * vb;
        s11 += w * va * va;
        s22 += w * vb * vb;
        s12 += w * va * vb;
      }
    double m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    s11 -= m11;
    s22 -= m22;
    s12 -= m12;
    double v1 = 2.0 * s12 + c2, v2 = s11 + s22 + c2;
    my_ssim += ((2.0 * m12 + c1) * v1) / ((m11 + m22 + c1) * v2);
    my_cs += v1 / v2;
  }
  s_red[0][threadIdx.x] = my_ssim;
  s_red[1][threadIdx.x] = my_cs;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      s_red[0][threadIdx.x] += s_red[0][threadIdx.x + s];
      s_red[1][threadIdx.x] += s_red[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(sums + g * 2 + 0, s_red[0][0] / (double)per_group);
    atomicAdd(sums + g * 2 + 1, s_red[1][0] / (double)per_group);
  }
}

namespace {
struct MsPlan {
  int H[MS_LEVELS], W[MS_LEVELS];
  int64_t off[MS_LEVELS];  // offset (in doubles) of image-1 data of each level; image 2 follows it
  int64_t lvl_off, total;
};
MsPlan ms_plan(int groups, int batch, int height, int width, int depth) {
  MsPlan p;
  int64_t nb = (int64_t)groups * batch, o = 0;
  int H = height, W = width;
  for (int l = 0; l < MS_LEVELS; ++l) {
    p.H[l] = H;
    p.W[l] = W;
    p.off[l] = o;
    o += 2 * nb * H * W * depth;
    H = (H + 1) / 2;
    W = (W + 1) / 2;
  }
  p.lvl_off = o;
  p.total = o + (int64_t)groups * MS_LEVELS * 2;
  return p;
}
}  // namespace

extern "C" int64_t dsin_msssim_workspace_bytes(int groups, int batch, int height, int width, int depth) {
  return ms_plan(groups, batch, height, width, depth).total * (int64_t)sizeof(double) + 256;
}

extern "C" int dsin_msssim(dsin_handle_t h, const float* img1, const float* img2, int groups, int batch, int height,
                           int width, int depth, double* out_g52, void* workspace, void* stream) {
  DSIN_REQUIRE(h, img1 && img2 && out_g52 && workspace, "null pointer");
  DSIN_REQUIRE(h, groups > 0 && batch > 0 && height > 0 && width > 0 && depth > 0, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const MsPlan pl = ms_plan(groups, batch, height, width, depth);
  const int64_t nb = (int64_t)groups * batch;
  double* ws = (double*)workspace;
  double* lvl = ws + pl.lvl_off;  // (level, group, 2)
  const int64_t n0 = nb * height * width * depth;
  ms_f32_to_f64_kernel<<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(img1, ws + pl.off[0], n0);
  DSIN_LAUNCHED(h);
  ms_f32_to_f64_kernel<<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(img2, ws + pl.off[0] + n0, n0);
  DSIN_LAUNCHED(h);
  if (cudaMemsetAsync(lvl, 0, sizeof(double) * groups * MS_LEVELS * 2, st) != cudaSuccess)
    return dsin_fail(h, DSIN_ERR_CUDA, "%s: memset failed", __func__);
  for (int level = 0; level < MS_LEVELS; ++level) {
    const int H = pl.H[level], W = pl.W[level];
    const int64_t nl = nb * H * W * depth;
    double* a = ws + pl.off[level];
    double* b = a + nl;
    int size = 11 < H ? 11 : H;
    size = size < W ? size : W;
    const double sigma = size * 1.5 / 11.0;
    const int64_t per_group = (int64_t)batch * (H - size + 1) * (W - size + 1) * depth;
    int blocks = (int)((per_group + 255) / 256);
    blocks = blocks > 1024 ? 1024 : (blocks < 1 ? 1 : blocks);
    ms_ssim_level_kernel<<<dim3(blocks, groups), 256, 0, st>>>(a, b, groups, batch, H, W, depth, size, sigma,
                                                             lvl + (int64_t)level * groups * 2);
    DSIN_LAUNCHED(h);
    if (level + 1 < MS_LEVELS) {
      const int64_t n2 = nb * pl.H[level + 1] * pl.W[level + 1] * depth;
      double* na = ws + pl.off[level + 1];
      ms_down_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(a, na, nb, H, W, depth);
      DSIN_LAUNCHED(h);
      ms_down_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(b, na + n2, nb, H, W, depth);
      DSIN_LAUNCHED(h);
    }
  }
  // (level, group, 2) -> (group, level, 2)
  for (int level = 0; level < MS_LEVELS; ++level)
    if (cudaMemcpy2DAsync(out_g52 + level * 2, sizeof(double) * MS_LEVELS * 2, lvl + (int64_t)level * groups * 2,
                          sizeof(double) * 2, sizeof(double) * 2, groups, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return dsin_fail(h, DSIN_ERR_CUDA, "%s: copy failed", __func__);
  return DSIN_OK;
}
