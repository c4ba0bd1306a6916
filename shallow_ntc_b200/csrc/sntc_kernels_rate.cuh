// Rate term of the decode path: per-image bits of the latents under the entropy models the
// reference builds in frame_loss_given_latent_rvs (mshyper/models.py:246-259, 278-279, 300-310), training=False:
//   bits_y[b] = -sum log2 P(q | sigma = SCALE_FN(i_c)),  P = Phi((q+.5)/sigma) - Phi((q-.5)/sigma)   (tfc.NoisyNormal,
//               loc removed; i_c = clamp(exp(raw_sigma), 0, S-1) is the CONTINUOUS index, as in the eval path)
//   bits_z[b] = -sum log2 ( c(z+.5) - c(z-.5) ),  c = sigmoid(logits_cumulative)                      (tfc.NoisyDeepFactorized)
// Both are evaluated in the log domain (survival function on the upper tail), the way tfc's UniformNoiseAdapter
// does, so far tails stay finite in float32.  Per-image sums are deterministic: fixed partial slots, fixed order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sntc {

struct RateConst { float max_index, log_scale_min, scale_factor; };   // mshyper/models.py:27-32

// bits of integer symbol q under NoisyNormal(0, SCALE_FN(clamp(exp(raw_sigma)))):  P = Phi(b) - Phi(a), a = (|q| - .5) / sigma,
// b = (|q| + .5) / sigma.  One branch-free form for the bulk AND the far tails, through the scaled complementary error
// function erfcx(x) = exp(x^2) erfc(x) (smooth, ~1/(x sqrt(pi)) for large x), with a' = a / sqrt(2), b' = b / sqrt(2):
//     P = 1/2 exp(-a'^2) [ erfcx(a') - exp(-(b'^2 - a'^2)) erfcx(b') ]
//     bits = a'^2 log2(e) - log2( 1/2 [ ... ] )
// exp(-a'^2) never underflows because it stays in the log domain (|q| = 127 at sigma = 0.11: a' = 813, bits = 9.5e5), which
// is what tfc's UniformNoiseAdapter achieves with log survival functions; and there is no erfc / log-sf pair per edge, no
// expm1 / log1p and no divergent tail branch: ~150 instructions per element instead of ~500 (rate kernel 0.21 -> 0.08 ms per
// 24-image step).  q = 0 (a < 0) is P = erf(b').  Against the float64 log_ndtr form on 2 M random (q, i_c):
// max relative error 4.5e-5 per element, 1.7e-7 on the sum.
__device__ __forceinline__ float noisy_normal_bits(float q, float raw_sigma, const RateConst& rc) {
  const float i_c = fminf(fmaxf(expf(raw_sigma), 0.f), rc.max_index);
  const float sigma = expf(rc.log_scale_min + rc.scale_factor * i_c);   // SCALE_FN(i)   :32
  const float aq = fabsf(q), k = 0.70710678118f / sigma;
  const float bp = (aq + 0.5f) * k;
  if (aq == 0.f) return -log2f(erff(bp));
  const float ap = (aq - 0.5f) * k;
  const float delta = (bp - ap) * (bp + ap);
  const float t = erfcxf(ap) - expf(-delta) * erfcxf(bp);
  return ap * ap * 1.44269504089f - log2f(0.5f * t);
}

__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

// DeepFactorized(num_filters=(3,3,3)) parameters of one channel, pre-transformed on the host:
// [0,3) softplus(matrix_0) | [3,6) bias_0 | [6,9) tanh(factor_0) | [9,18) softplus(matrix_1) | [18,21) bias_1 | [21,24) tanh(factor_1)
// | [24,33) softplus(matrix_2) | [33,36) bias_2 | [36,39) tanh(factor_2) | [39,42) softplus(matrix_3) | [42] bias_3   (stride 44)
constexpr int DF_STRIDE = 44;
__device__ __forceinline__ float df_logits(const float* p, float x) {
  float h[3], g[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { float v = fmaf(p[i], x, p[3 + i]); h[i] = v + p[6 + i] * tanhf(v); }
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const float* m = p + 9 + 15 * l;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float v = m[9 + i];
#pragma unroll
      for (int j = 0; j < 3; ++j) v = fmaf(m[i * 3 + j], h[j], v);
      g[i] = v + m[12 + i] * tanhf(v);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) h[i] = g[i];
  }
  float v = p[42];
#pragma unroll
  for (int j = 0; j < 3; ++j) v = fmaf(p[39 + j], h[j], v);
  return v;
}

__device__ __forceinline__ float deep_factorized_bits(const float* p, float z) {
  const float lower = df_logits(p, z - 0.5f), upper = df_logits(p, z + 0.5f);
  // sigmoid(u) - sigmoid(l) on the side of the median where both are small (tfc deep_factorized.py _prob sign trick)
  const float sgn = (lower + upper) > 0.f ? -1.f : 1.f;
  const float u = sgn * upper, l = sgn * lower;
  const float big = fmaxf(u, l), small = fminf(u, l);
  const float Lb = log_sigmoid(big), Ls = log_sigmoid(small);
  const float lp = Lb + logf(-expm1f(Ls - Lb));
  return -lp * 1.44269504089f;
}

__device__ __forceinline__ double block_sum_256(double v, double* sh) {   // deterministic tree, blockDim.x == 256
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  return sh[0];
}

// fp32 path: hs [npix, 2C] is materialised; grid (blocks_per_image, B), block 256; slots[b * gridDim.x + blockIdx.x]
__global__ void __launch_bounds__(256) rate_y_kernel(const float* __restrict__ hs, const void* __restrict__ q, int q_kind, size_t pix_per_image, int C,
                                                     RateConst rc, double* __restrict__ slots) {
  __shared__ double sh[256];
  const int b = blockIdx.y;
  const size_t n = pix_per_image * (size_t)C;
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t pix = (size_t)b * pix_per_image + i / C;
    const int c = (int)(i % C);
    const float raw = __ldg(hs + pix * 2 * C + C + c);
    const size_t e = pix * C + c;
    float qv;
    if (q_kind == 0) qv = __ldg(reinterpret_cast<const float*>(q) + e);
    else if (q_kind == 1) qv = (float)reinterpret_cast<const int16_t*>(q)[e];
    else qv = (float)reinterpret_cast<const int8_t*>(q)[e];
    acc += (double)noisy_normal_bits(qv, raw, rc);
  }
  const double tot = block_sum_256(acc, sh);
  if (threadIdx.x == 0) slots[(size_t)b * gridDim.x + blockIdx.x] = tot;
}

// Tensor-core path: the hyper-head epilogue leaves raw sigma [B, n_per_image] (fp32) next to idx, and the bits are summed
// here by a full-occupancy elementwise kernel (2048 threads per SM) instead of by the 8 epilogue warps of the GEMM, where
// the two erfc + log per element made the epilogue longer than the MMAs (hyper-synthesis layer 2: 0.36 -> 0.63 ms).
// 4 elements per thread and iteration, 16-byte loads; n_per_image % 4 == 0; grid (blocks_per_image, B), block 256.
__global__ void __launch_bounds__(256) rate_y_flat_kernel(const float* __restrict__ sigma, const void* __restrict__ q, int q_kind, size_t n_per_image,
                                                          RateConst rc, double* __restrict__ slots) {
  __shared__ double sh[256];
  const int b = blockIdx.y;
  const size_t n4 = n_per_image / 4, base4 = (size_t)b * n4;
  float acc = 0.f;
  double tot_acc = 0.0;
  int cnt = 0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    const float4 sg = __ldg(reinterpret_cast<const float4*>(sigma) + base4 + i);
    float4 qv;
    if (q_kind == 0) qv = __ldg(reinterpret_cast<const float4*>(q) + base4 + i);
    else if (q_kind == 1) {
      const uint2 r = __ldg(reinterpret_cast<const uint2*>(q) + base4 + i);
      qv = make_float4((float)(int16_t)(r.x & 0xFFFFu), (float)(int16_t)(r.x >> 16), (float)(int16_t)(r.y & 0xFFFFu), (float)(int16_t)(r.y >> 16));
    } else {
      const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(q) + base4 + i);
      qv = make_float4((float)(int8_t)(r & 0xFFu), (float)(int8_t)((r >> 8) & 0xFFu), (float)(int8_t)((r >> 16) & 0xFFu), (float)(int8_t)(r >> 24));
    }
    acc += noisy_normal_bits(qv.x, sg.x, rc) + noisy_normal_bits(qv.y, sg.y, rc) + noisy_normal_bits(qv.z, sg.z, rc) + noisy_normal_bits(qv.w, sg.w, rc);
    if (++cnt == 16) { tot_acc += (double)acc; acc = 0.f; cnt = 0; }   // short fp32 runs, double across them
  }
  tot_acc += (double)acc;
  const double tot = block_sum_256(tot_acc, sh);
  if (threadIdx.x == 0) slots[(size_t)b * gridDim.x + blockIdx.x] = tot;
}

// z_hat [B, n_per_image = hz*wz*Cz]; grid (blocks_per_image, B)
__global__ void __launch_bounds__(256) rate_z_kernel(const float* __restrict__ z, size_t n_per_image, int Cz, const float* __restrict__ prior,
                                                     double* __restrict__ slots) {
  __shared__ double sh[256];
  const int b = blockIdx.y;
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_per_image; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % Cz);
    acc += (double)deep_factorized_bits(prior + (size_t)c * DF_STRIDE, __ldg(z + (size_t)b * n_per_image + i));
  }
  const double tot = block_sum_256(acc, sh);
  if (threadIdx.x == 0) slots[(size_t)b * gridDim.x + blockIdx.x] = tot;
}

// out[b * out_stride] = sum of the slots of image b, in slot order.  slot_img == nullptr: slots are laid out
// [B][slots_per_image]; otherwise slot i belongs to image slot_img[i] (-1 = unused).  grid B, block 256.
__global__ void __launch_bounds__(256) rate_reduce_kernel(const double* __restrict__ slots, const int* __restrict__ slot_img, int nslots, int slots_per_image,
                                                          double* __restrict__ out, int out_stride) {
  __shared__ double sh[256];
  const int b = blockIdx.x;
  double acc = 0.0;
  if (slot_img) {
    for (int i = threadIdx.x; i < nslots; i += 256) if (slot_img[i] == b) acc += slots[i];
  } else {
    for (int i = threadIdx.x; i < slots_per_image; i += 256) acc += slots[(size_t)b * slots_per_image + i];
  }
  const double tot = block_sum_256(acc, sh);
  if (threadIdx.x == 0) out[(size_t)b * out_stride] = tot;
}

}  // namespace sntc
