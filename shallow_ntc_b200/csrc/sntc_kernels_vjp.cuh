// Decoder backward: the pointwise stages of the vector-Jacobian products of the synthesis / hyper-synthesis transforms
// (what tf.GradientTape computes through self._synthesis / self._hyper_synthesis in itinf_train_step,
// mshyper/models.py:401-408, :273, :297).  The input-gradient of every transposed conv runs as a FORWARD stride-1 band GEMM
// (sntc_plan.hpp make_backward_conv; tcgen05 under the tensor-core precision) on the shifted space-to-depth of the
// gradient that s2d_grad_kernel writes; the adjoints of GDN1 / relu / leaky_relu are the kernels below.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>

#include "sntc_kernels_f32.cuh"

namespace sntc {

// G'[b, my, mx, (ry*s + rx)*Cf + co] = g[b, s*my + ry - p, s*mx + rx - p, co] * act'(a[...]),  zero outside g and for the
// padding channels c' >= s*s*Cf.  `a` (nullable) is the saved POST-activation output of the forward layer (relu / leaky_relu
// fused into it: the derivative is 1 where a > 0, else 0 / 0.2).  One thread per 4 consecutive output channels.
struct S2dGradParams {
  const float* g; const float* a; int act;
  int B, H, W, Cf;                 // g, a: [B,H,W,Cf]
  int s, p, hm, wm, Cpad;          // G': [B,hm,wm,Cpad]
  float* out; __half* hi; __half* lo;
};

__global__ void __launch_bounds__(256) s2d_grad_kernel(const S2dGradParams P) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cq = P.Cpad / 4;
  const size_t total = (size_t)P.B * P.hm * P.wm * cq;
  if (i >= total) return;
  const int c0 = (int)(i % cq) * 4;
  size_t t = i / cq;
  const int mx = (int)(t % P.wm); t /= P.wm;
  const int my = (int)(t % P.hm); const int b = (int)(t / P.hm);
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + k, r = c / P.Cf, co = c - r * P.Cf;
    v[k] = 0.f;
    if (r < P.s * P.s) {
      const int y = P.s * my + r / P.s - P.p, x = P.s * mx + r % P.s - P.p;
      if (y >= 0 && y < P.H && x >= 0 && x < P.W) {
        const size_t e = (((size_t)b * P.H + y) * P.W + x) * P.Cf + co;
        float gv = __ldg(P.g + e);
        if (P.a) {
          const float av = __ldg(P.a + e);
          if (!(av > 0.f)) gv = P.act == SNTC_ACT_LEAKY_RELU ? gv * 0.2f : 0.f;
        }
        v[k] = gv;
      }
    }
  }
  const size_t o = i * 4;
  if (P.out) *reinterpret_cast<float4*>(P.out + o) = make_float4(v[0], v[1], v[2], v[3]);
  if (P.hi) {
    __align__(8) __half h[4];
    __align__(8) __half l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float c = fminf(fmaxf(v[k], -65504.f), 65504.f);
      h[k] = __float2half_rn(c);
      l[k] = __float2half_rn(c - __half2float(h[k]));
    }
    *reinterpret_cast<uint2*>(P.hi + o) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(P.lo + o) = *reinterpret_cast<const uint2*>(l);
  }
}

// Adjoint of a pointwise activation over C channels at the saved input x:  out = J_act(x)^T g.
//   relu / leaky_relu / none : g * act'(x)
//   GDN1  (common/transforms.py:27-63; n_j = beta_j + sum_i |x_i| gamma_ij):
//     inverse: out_k = g_k n_k + sign(x_k) sum_j g_j x_j gamma_kj          forward: out_k = g_k / n_k - sign(x_k) sum_j g_j x_j / n_j^2 gamma_kj
//   classic GDN (n_j = sqrt(beta_j + sum_i x_i^2 gamma_ij)):
//     inverse: out_k = g_k n_k + x_k sum_j g_j x_j / n_j gamma_kj          forward: out_k = g_k / n_k - x_k sum_j g_j x_j / n_j^3 gamma_kj
// `res_copy`: out has 2C channels and its second half receives g unchanged (the residual branch of TwoLayerResSynthesis,
// whose forward input is [base || res] with stride 2C).
struct ActBwdParams {
  const float* x; int x_stride; const float* g; int g_stride; float* out; int out_stride;
  size_t npix; int C; int act; int inverse; int classic; int res_copy;
  const float* beta; const float* gamma; int gamma_stride;   // gamma [in][out]
  const float* gamma_t;                                      // [out][in] (wide kernel only), row stride gamma_stride
};

__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// C <= 64: one thread per pixel; gamma, beta and the x / w tiles in shared memory
__global__ void __launch_bounds__(128) act_bwd_kernel(const ActBwdParams P) {
  extern __shared__ float sm[];                 // gamma [C][C] | beta [C] | x tile [128][C+1] | w tile [128][C+1]
  float* sg = sm; float* sb = sg + P.C * P.C; float* sx = sb + P.C; float* sw = sx + 128 * (P.C + 1);
  const bool gdn = P.act == SNTC_ACT_IGDN1 || P.act == SNTC_ACT_GDN1 || P.act == SNTC_ACT_IGDN_CLASSIC;
  if (gdn) {
    for (int i = threadIdx.x; i < P.C * P.C; i += blockDim.x) sg[i] = P.gamma[(size_t)(i / P.C) * P.gamma_stride + (i % P.C)];
    for (int i = threadIdx.x; i < P.C; i += blockDim.x) sb[i] = P.beta[i];
  }
  const size_t pix0 = (size_t)blockIdx.x * blockDim.x;
  const int npx = (int)min((size_t)blockDim.x, P.npix - pix0);
  for (int i = threadIdx.x; i < npx * P.C; i += blockDim.x) {
    const int r = i / P.C, c = i - r * P.C;
    sx[r * (P.C + 1) + c] = P.x[(pix0 + r) * P.x_stride + c];
    sw[r * (P.C + 1) + c] = P.g[(pix0 + r) * P.g_stride + c];
  }
  __syncthreads();
  if ((int)threadIdx.x >= npx) return;
  float* xr = sx + threadIdx.x * (P.C + 1);
  float* wr = sw + threadIdx.x * (P.C + 1);       // g on entry; w_j in place afterwards
  const size_t pix = pix0 + threadIdx.x;
  float* o = P.out + pix * P.out_stride;
  if (P.res_copy) for (int j = 0; j < P.C; ++j) o[P.C + j] = wr[j];
  if (!gdn) {
    for (int j = 0; j < P.C; ++j) {
      const float x = xr[j];
      float d = 1.f;
      if (P.act == SNTC_ACT_RELU) d = x > 0.f ? 1.f : 0.f;
      else if (P.act == SNTC_ACT_LEAKY_RELU) d = x > 0.f ? 1.f : 0.2f;
      o[j] = wr[j] * d;
    }
    return;
  }
  // first term and w_j; the first term goes straight to `out`
  for (int j = 0; j < P.C; ++j) {
    float n = sb[j];
    if (P.classic) { for (int i = 0; i < P.C; ++i) n = fmaf(xr[i] * xr[i], sg[i * P.C + j], n); n = sqrtf(n); }
    else for (int i = 0; i < P.C; ++i) n = fmaf(fabsf(xr[i]), sg[i * P.C + j], n);
    const float gj = wr[j], xj = xr[j];
    float w;
    if (P.inverse) { o[j] = gj * n; w = P.classic ? gj * xj / n : gj * xj; }
    else { o[j] = gj / n; w = P.classic ? -gj * xj / (n * n * n) : -gj * xj / (n * n); }
    wr[j] = w;
  }
  for (int k = 0; k < P.C; ++k) {
    float s = 0.f;
    for (int j = 0; j < P.C; ++j) s = fmaf(wr[j], sg[k * P.C + j], s);
    o[k] += (P.classic ? xr[k] : sgnf(xr[k])) * s;
  }
}

// any C (deep decoders: 192 / 256): a block handles GB_PT pixels, thread t owns columns t, t + 256, ...; gamma streams from L2
constexpr int GB_PT = 32;
__global__ void __launch_bounds__(256) gdn_bwd_wide_kernel(const ActBwdParams P) {
  extern __shared__ float sm[];                 // x tile [GB_PT][C] | w tile [GB_PT][C]
  float* sx = sm; float* sw = sx + GB_PT * P.C;
  const size_t pix0 = (size_t)blockIdx.x * GB_PT;
  const int npx = (int)min((size_t)GB_PT, P.npix - pix0);
  for (int i = threadIdx.x; i < GB_PT * P.C; i += blockDim.x) {
    const int r = i / P.C, c = i - r * P.C;
    sx[i] = r < npx ? P.x[(pix0 + r) * P.x_stride + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < P.C; j += blockDim.x) {
    float n[GB_PT];
    const float b = P.beta[j];
#pragma unroll
    for (int p = 0; p < GB_PT; ++p) n[p] = b;
    for (int i = 0; i < P.C; ++i) {
      const float gm = __ldg(P.gamma + (size_t)i * P.gamma_stride + j);
#pragma unroll
      for (int p = 0; p < GB_PT; ++p) {
        const float xv = sx[p * P.C + i];
        n[p] = fmaf(P.classic ? xv * xv : fabsf(xv), gm, n[p]);
      }
    }
#pragma unroll
    for (int p = 0; p < GB_PT; ++p) {
      float w = 0.f;
      if (p < npx) {
        const float nn = P.classic ? sqrtf(n[p]) : n[p];
        const float gj = P.g[(pix0 + p) * P.g_stride + j], xj = sx[p * P.C + j];
        float first;
        if (P.inverse) { first = gj * nn; w = P.classic ? gj * xj / nn : gj * xj; }
        else { first = gj / nn; w = P.classic ? -gj * xj / (nn * nn * nn) : -gj * xj / (nn * nn); }
        P.out[(pix0 + p) * P.out_stride + j] = first;
      }
      sw[p * P.C + j] = w;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < P.C; k += blockDim.x) {
    float s[GB_PT];
#pragma unroll
    for (int p = 0; p < GB_PT; ++p) s[p] = 0.f;
    for (int j = 0; j < P.C; ++j) {
      const float gm = __ldg(P.gamma_t + (size_t)j * P.gamma_stride + k);   // gamma[k][j]
#pragma unroll
      for (int p = 0; p < GB_PT; ++p) s[p] = fmaf(sw[p * P.C + j], gm, s[p]);
    }
#pragma unroll
    for (int p = 0; p < GB_PT; ++p)
      if (p < npx) {
        const float xk = sx[p * P.C + k];
        P.out[(pix0 + p) * P.out_stride + k] += (P.classic ? xk : sgnf(xk)) * s[p];   // same thread wrote the first term
      }
  }
}

}  // namespace sntc
