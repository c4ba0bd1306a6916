// LPIPS on the device (the `lpips` column of the evaluate records).
// Reference: lpips_tf2/lpips_tensorflow.py:14-72 as called by mshyper/models.py:334-340 / factorized/models.py:158-164:
//   preprocess (x / 127.5 - 1, (x - shift) / scale)  ->  Keras VGG16 features of block1_conv2, block2_conv2, block3_conv3,
//   block4_conv3, block5_conv3  ->  per pixel x * rsqrt(sum_c x^2) (no epsilon)  ->  (a - b)^2  ->  1x1 conv to one channel (no
//   bias)  ->  spatial mean  ->  sum over the five layers.
// The 13 3x3 'same' convolutions are transposed convolutions with the flipped kernel (ConvT(3, 1, p = 1): out[o] = sum_a' in[o - a'
// + 1] W[a'], W[a'] = K[2 - a']), so they run on the band-GEMM machinery of the decode path: conv_0 (3 input channels, not
// TMA-addressable) on the FFMA band GEMM, conv_1 .. conv_12 on band_gemm_tc_kernel (tcgen05, split-fp16 3-pass, relu in the
// epilogue).  This file holds the stages around them: preprocessing, 2x2 max-pool fused with the fp16 hi/lo split of the next
// conv's A operand, and the normalise / difference / lin / mean head (deterministic two-level sums in double).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace sntc {

struct LpipsPre { float scale[3]; float shift[3]; };

// images [n,H,W,3] (uint8 or float32 in [0, 255]) -> [n,H,W,4] float32 preprocessed, channel 3 = 0 (conv_0's padded input)
__global__ void lpips_preprocess_kernel(const void* __restrict__ img, int is_u8, size_t npix, LpipsPre P, float4* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  float v[3];
  if (is_u8) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(img) + i * 3;
    v[0] = (float)p[0]; v[1] = (float)p[1]; v[2] = (float)p[2];
  } else {
    const float* p = reinterpret_cast<const float*>(img) + i * 3;
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn(__fsub_rn(__fsub_rn(__fdiv_rn(v[c], 127.5f), 1.0f), P.shift[c]), P.scale[c]);
  out[i] = make_float4(v[0], v[1], v[2], 0.f);
}

// MaxPooling2D(2, 2, 'valid') of an fp32 NHWC tensor, written as the fp16 hi/lo planes the next tensor-core conv reads
// (and optionally as fp32).  One thread = 8 channels of one output pixel.  C % 8 == 0.
__global__ void lpips_maxpool_planes_kernel(const float* __restrict__ x, int n, int h, int w, int C, __half* __restrict__ hi, __half* __restrict__ lo,
                                            float* __restrict__ out_f32) {
  const int ho = h / 2, wo = w / 2, c8 = C / 8;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n * ho * wo * c8;
  if (i >= total) return;
  const int cg = (int)(i % c8);
  size_t p = i / c8;
  const int ox = (int)(p % wo); p /= wo;
  const int oy = (int)(p % ho);
  const int b = (int)(p / ho);
  const float* base = x + (((size_t)b * h + 2 * oy) * w + 2 * ox) * C + cg * 8;
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(base + ((size_t)dy * w + dx) * C));
      const float4 c = __ldg(reinterpret_cast<const float4*>(base + ((size_t)dy * w + dx) * C) + 1);
      m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], a.z); m[3] = fmaxf(m[3], a.w);
      m[4] = fmaxf(m[4], c.x); m[5] = fmaxf(m[5], c.y); m[6] = fmaxf(m[6], c.z); m[7] = fmaxf(m[7], c.w);
    }
  const size_t o = ((((size_t)b * ho + oy) * wo + ox) * C) + cg * 8;
  if (out_f32) {
    *reinterpret_cast<float4*>(out_f32 + o) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(out_f32 + o + 4) = make_float4(m[4], m[5], m[6], m[7]);
  }
  if (hi) {
    __align__(16) __half hh[8];
    __align__(16) __half ll[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float c = fminf(fmaxf(m[k], -65504.f), 65504.f);
      hh[k] = __float2half_rn(c);
      ll[k] = __float2half_rn(c - __half2float(hh[k]));
    }
    *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(hh);
    *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(ll);
  }
}

// One layer of the head.  feat [2B, h, w, C] fp32: images a = feat[0..B), b = feat[B..2B).  One WARP per pixel: lanes stride the
// channels, two passes (norms, then the weighted squared difference of the unit vectors).  Per-block partial sums go to
// slots[b][blockIdx.x] (double): fixed order, deterministic.  x * rsqrt(0) = NaN for an all-zero feature vector, as in the
// reference (no epsilon in lpips_tensorflow.py:53-56).
__global__ void __launch_bounds__(256) lpips_head_kernel(const float* __restrict__ feat, int B, int npix, int C, const float* __restrict__ lin,
                                                         double* __restrict__ slots, int slots_per_image) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * 8;
  double acc = 0.0;
  for (int p = blockIdx.x * 8 + warp; p < npix; p += warps_total) {
    const float* fa = feat + ((size_t)b * npix + p) * C;
    const float* fb = feat + ((size_t)(B + b) * npix + p) * C;
    float sa = 0.f, sb = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(fa + c)), q = __ldg(reinterpret_cast<const float4*>(fb + c));
      sa += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
      sb += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    const float ra = 1.0f / sqrtf(sa), rb = 1.0f / sqrtf(sb);
    float d = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(fa + c)), q = __ldg(reinterpret_cast<const float4*>(fb + c));
      const float4 wv = __ldg(reinterpret_cast<const float4*>(lin + c));
      // normalise each side (rounded), THEN subtract, as the reference does: a fused a*ra - q*rb would leave the rounding error of
      // one product behind and make lpips(x, x) non-zero
      float t;
      t = __fsub_rn(__fmul_rn(a.x, ra), __fmul_rn(q.x, rb)); d = fmaf(t * t, wv.x, d);
      t = __fsub_rn(__fmul_rn(a.y, ra), __fmul_rn(q.y, rb)); d = fmaf(t * t, wv.y, d);
      t = __fsub_rn(__fmul_rn(a.z, ra), __fmul_rn(q.z, rb)); d = fmaf(t * t, wv.z, d);
      t = __fsub_rn(__fmul_rn(a.w, ra), __fmul_rn(q.w, rb)); d = fmaf(t * t, wv.w, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    acc += (double)d;
  }
  __shared__ double sh[8];
  if (lane == 0) sh[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    slots[(size_t)b * slots_per_image + blockIdx.x] = t;
  }
}

// out[b] += (sum of the image's slots) / npix      (spatial mean of this layer, added to the running sum over layers)
__global__ void lpips_layer_finalize_kernel(const double* __restrict__ slots, int slots_per_image, double inv_npix, double* __restrict__ out) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < slots_per_image; ++i) t += slots[(size_t)b * slots_per_image + i];
    out[b] += t * inv_npix;
  }
}

}  // namespace sntc
