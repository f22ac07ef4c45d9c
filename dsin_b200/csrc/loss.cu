// Reduction terms of the validation loss (AE.siNet_validate -> loss_test, src/AE.py:76-99,120-131): one pass over the
// images and one over the bottleneck, fp64 per-image sums.  HBM-bound: 3 images of 4.7 MB + 2 volumes of 0.8 MB per pair.
//   terms[img][0] = sum |x_dec - x|  (or sum (x_dec - x)^2)      Distortions.get_mae_per_img / get_mse_per_img
//   terms[img][1] = sum |x - x_with_si|                          tf.losses.absolute_difference (src/AE.py:94)
//   terms[img][2] = sum bc                                       get_loss: H_real (src/Distortions_imgcomp.py:120)
//   terms[img][3] = sum bc * heatmap                             get_loss: H_mask (:119,121)
// The scalar arithmetic that follows (means, max(H_soft - H_target, 0), weights) is a dozen flops on the host.
#include "common.cuh"

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// blockDim.x == 256; out[k] += block sum of v[k]
template <int K>
__device__ __forceinline__ void block_accumulate(double (&v)[K], double* out) {
  __shared__ double s_part[K][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double w = warp_sum(v[k]);
    if (lane == 0) s_part[k][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double w = lane < 8 ? s_part[k][lane] : 0.0;
      w = warp_sum(w);
      if (lane == 0) atomicAdd(out + k, w);
    }
  }
}

__device__ __forceinline__ float dist(float a, float b, int squared) {
  const float d = __fsub_rn(a, b);
  return squared ? __fmul_rn(d, d) : fabsf(d);
}

// grid (blocks, n): image part.  vec = 1 when img_elems % 4 == 0 (every image then starts 16-byte aligned)
__global__ void __launch_bounds__(256) loss_image_terms_kernel(const float* __restrict__ x, const float* __restrict__ x_dec,
                                                                const float* __restrict__ x_si, int64_t img_elems,
                                                                int squared, int vec, double* __restrict__ terms) {
  const int img = blockIdx.y;
  const int64_t base = (int64_t)img * img_elems;
  double acc[2] = {0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    const float4* d4 = reinterpret_cast<const float4*>(x_dec + base);
    const float4* s4 = x_si ? reinterpret_cast<const float4*>(x_si + base) : nullptr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < img_elems / 4; i += stride) {
      const float4 a = x4[i], b = d4[i];
      // fp32 partial of four terms (each < 2^16): exact enough to be invisible after the fp64 accumulation
      acc[0] += (double)dist(b.x, a.x, squared) + (double)dist(b.y, a.y, squared) + (double)dist(b.z, a.z, squared) +
                (double)dist(b.w, a.w, squared);
      if (s4) {
        const float4 c = s4[i];
        acc[1] += (double)fabsf(__fsub_rn(a.x, c.x)) + (double)fabsf(__fsub_rn(a.y, c.y)) +
                  (double)fabsf(__fsub_rn(a.z, c.z)) + (double)fabsf(__fsub_rn(a.w, c.w));
      }
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < img_elems; i += stride) {
      const float a = x[base + i];
      acc[0] += (double)dist(x_dec[base + i], a, squared);
      if (x_si) acc[1] += (double)fabsf(__fsub_rn(a, x_si[base + i]));
    }
  }
  block_accumulate<2>(acc, terms + (int64_t)img * 4);
}

// grid (blocks, n): bottleneck part
__global__ void __launch_bounds__(256) loss_rate_terms_kernel(const float* __restrict__ bc, const float* __restrict__ hm,
                                                               int64_t sym_elems, double* __restrict__ terms) {
  const int img = blockIdx.y;
  const int64_t base = (int64_t)img * sym_elems;
  double acc[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sym_elems; i += (int64_t)gridDim.x * blockDim.x) {
    const float b = bc[base + i];
    acc[0] += (double)b;
    if (hm) acc[1] += (double)__fmul_rn(b, hm[base + i]);
  }
  block_accumulate<2>(acc, terms + (int64_t)img * 4 + 2);
}

}  // namespace

extern "C" int dsin_validation_terms(dsin_handle_t h, const float* x, const float* x_dec, const float* x_with_si,
                                     const float* bitcost, const float* heatmap, int n, int64_t img_elems,
                                     int64_t sym_elems, int squared, double* terms_n4, void* stream) {
  DSIN_REQUIRE(h, x && x_dec && bitcost && terms_n4 && n > 0 && img_elems > 0 && sym_elems > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(terms_n4, 0, sizeof(double) * 4 * (size_t)n, st) != cudaSuccess)
    return dsin_fail(h, DSIN_ERR_CUDA, "%s", "cudaMemsetAsync of the loss terms failed");
  const int vec = (img_elems % 4 == 0) && (((uintptr_t)x | (uintptr_t)x_dec | (uintptr_t)x_with_si) % 16 == 0);
  const int64_t work = vec ? img_elems / 4 : img_elems;
  // two waves of 148 SMs x 8 resident blocks at most; one image of 320x1224 is 1 148 blocks of 256 float4s
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (2 * 148 * 8 + n - 1) / n;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  loss_image_terms_kernel<<<dim3((unsigned)blocks, (unsigned)n), 256, 0, st>>>(x, x_dec, x_with_si, img_elems, squared,
                                                                               vec, terms_n4);
  DSIN_LAUNCHED(h);
  int64_t rblocks = (sym_elems + 255) / 256;
  if (rblocks > cap) rblocks = cap;
  loss_rate_terms_kernel<<<dim3((unsigned)rblocks, (unsigned)n), 256, 0, st>>>(bitcost, heatmap, sym_elems, terms_n4);
  DSIN_LAUNCHED(h);
  return DSIN_OK;
}
