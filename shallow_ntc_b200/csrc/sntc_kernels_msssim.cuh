// MS-SSIM of two uint8 image batches on the device: the validation metric of the reference's evaluate loop
// (mshyper/models.py:321-332, factorized/models.py:145-156 -> tf.image.ssim / tf.image.ssim_multiscale, TF 2.10
// image_ops_impl.py).  HBM-class work: every scale reads its two images once.
//
//   x = u8 / 255 (convert_image_dtype), max_val = 1  ->  c1 = 1e-4, c2 = 9e-4
//   per scale and channel, VALID 11x11 Gaussian window (sigma 1.5; the softmax-normalised 2-D window of tf.image is the
//   outer product of the normalised 1-D window, so the filter runs separably: rows, then columns, out of shared memory):
//       m0 = G*x, m1 = G*y, lum = (2 m0 m1 + c1) / (m0^2 + m1^2 + c1)
//       cs  = (2 G*(xy) - 2 m0 m1 + c2) / (G*(x^2 + y^2) - m0^2 - m1^2 + c2)
//       ssim = mean(lum * cs), cs = mean(cs)
//   5 scales, 2x2 mean pooling between them (odd sizes: last row / column repeated = SYMMETRIC pad of one);
//   msssim = mean_c  prod_k relu(f_k)^w_k,  f_k = cs_k (k < 4), ssim_4;  images with both sides < 160 px: single-scale ssim.
// Sums are deterministic: one (lum*cs, cs) partial per tile in double, reduced in a fixed order.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>

namespace sntc {

constexpr int MS_TX = 32, MS_TY = 16, MS_K = 11;
constexpr int MS_RX = MS_TX + MS_K - 1, MS_RY = MS_TY + MS_K - 1;   // input region of a tile: 42 x 26
constexpr int MS_SCALES = 5;

struct MsScaleParams {
  const void* a; const void* b; int is_u8;   // [B,H,W,C]: uint8 (scale 0) or float32 (pooled scales)
  int B, H, W, C, tiles_x, tiles_y;
  double* partial;                            // [B][C][tiles][2]
  float g[MS_K];
};

__device__ __forceinline__ float ms_load(const void* p, int is_u8, size_t i) {
  return is_u8 ? (float)reinterpret_cast<const uint8_t*>(p)[i] * (1.0f / 255.0f) : reinterpret_cast<const float*>(p)[i];
}

// grid (tiles, C, B), 256 threads: one 32 x 16 tile of the SSIM map of one channel of one image
__global__ void __launch_bounds__(256) msssim_scale_kernel(const MsScaleParams P) {
  __shared__ float sx[MS_RY][MS_RX + 1], sy[MS_RY][MS_RX + 1];
  __shared__ float sh[4][MS_RY][MS_TX];          // row-filtered x, y, xy, x^2 + y^2
  __shared__ double red[2][8];
  const int tile = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
  const int tx0 = (tile % P.tiles_x) * MS_TX, ty0 = (tile / P.tiles_x) * MS_TY;
  const int tid = threadIdx.x;
  for (int i = tid; i < MS_RY * MS_RX; i += 256) {
    const int r = i / MS_RX, q = i - r * MS_RX;
    const int gy = ty0 + r, gx = tx0 + q;
    float vx = 0.f, vy = 0.f;
    if (gy < P.H && gx < P.W) {
      const size_t e = (((size_t)b * P.H + gy) * P.W + gx) * P.C + c;
      vx = ms_load(P.a, P.is_u8, e); vy = ms_load(P.b, P.is_u8, e);
    }
    sx[r][q] = vx; sy[r][q] = vy;
  }
  __syncthreads();
  for (int i = tid; i < MS_RY * MS_TX; i += 256) {
    const int r = i / MS_TX, q = i - r * MS_TX;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int k = 0; k < MS_K; ++k) {
      const float x = sx[r][q + k], y = sy[r][q + k], w = P.g[k];
      s0 = fmaf(w, x, s0); s1 = fmaf(w, y, s1); s2 = fmaf(w, x * y, s2); s3 = fmaf(w, fmaf(x, x, y * y), s3);
    }
    sh[0][r][q] = s0; sh[1][r][q] = s1; sh[2][r][q] = s2; sh[3][r][q] = s3;
  }
  __syncthreads();
  const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
  double acc_ssim = 0.0, acc_cs = 0.0;
  const int ox = tid & 31;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int oy = (tid >> 5) + 8 * j;
    float m0 = 0.f, m1 = 0.f, exy = 0.f, esq = 0.f;
#pragma unroll
    for (int k = 0; k < MS_K; ++k) {
      const float w = P.g[k];
      m0 = fmaf(w, sh[0][oy + k][ox], m0); m1 = fmaf(w, sh[1][oy + k][ox], m1);
      exy = fmaf(w, sh[2][oy + k][ox], exy); esq = fmaf(w, sh[3][oy + k][ox], esq);
    }
    if (ty0 + oy < P.H - (MS_K - 1) && tx0 + ox < P.W - (MS_K - 1)) {
      const float num0 = m0 * m1 * 2.0f, den0 = m0 * m0 + m1 * m1;
      const float lum = (num0 + c1) / (den0 + c1);
      const float cs = (exy * 2.0f - num0 + c2) / (esq - den0 + c2);
      acc_ssim += (double)(lum * cs); acc_cs += (double)cs;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    acc_ssim += __shfl_down_sync(0xffffffffu, acc_ssim, off);
    acc_cs += __shfl_down_sync(0xffffffffu, acc_cs, off);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = acc_ssim; red[1][tid >> 5] = acc_cs; }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0, t = 0.0;
    for (int w = 0; w < 8; ++w) { s += red[0][w]; t += red[1][w]; }
    double* o = P.partial + (((size_t)b * P.C + c) * gridDim.x + tile) * 2;
    o[0] = s; o[1] = t;
  }
}

// grid (C, B), 256 threads: fixed-order sum of the tile partials -> stats[((b * MS_SCALES + scale) * C + c) * 2 + {ssim, cs}] (means)
__global__ void __launch_bounds__(256) msssim_finalize_kernel(const double* partial, int tiles, int C, int scale, double inv_count, double* stats) {
  __shared__ double red[2][256];
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const double* p = partial + ((size_t)b * C + c) * tiles * 2;
  double s = 0.0, t = 0.0;
  for (int i = tid; i < tiles; i += 256) { s += p[2 * i]; t += p[2 * i + 1]; }
  red[0][tid] = s; red[1][tid] = t;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) { red[0][tid] += red[0][tid + off]; red[1][tid] += red[1][tid + off]; }
    __syncthreads();
  }
  if (tid == 0) {
    double* o = stats + (((size_t)b * MS_SCALES + scale) * C + c) * 2;
    o[0] = red[0][0] * inv_count; o[1] = red[1][0] * inv_count;
  }
}

// 2x2 mean pooling [B,H,W,C] -> [B,(H+1)/2,(W+1)/2,C]; the last row / column of an odd size is repeated
__global__ void __launch_bounds__(256) msssim_pool_kernel(const void* in, int is_u8, float* out, int B, int H, int W, int C) {
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const size_t n = (size_t)B * H2 * W2 * C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int x = (int)(r % W2); r /= W2;
    const int y = (int)(r % H2); const int b = (int)(r / H2);
    const int y0 = 2 * y, y1 = min(2 * y + 1, H - 1), x0 = 2 * x, x1 = min(2 * x + 1, W - 1);
    const size_t base = (size_t)b * H * W;
    const float v00 = ms_load(in, is_u8, ((base + (size_t)y0 * W + x0) * C) + c), v01 = ms_load(in, is_u8, ((base + (size_t)y0 * W + x1) * C) + c);
    const float v10 = ms_load(in, is_u8, ((base + (size_t)y1 * W + x0) * C) + c), v11 = ms_load(in, is_u8, ((base + (size_t)y1 * W + x1) * C) + c);
    out[i] = ((v00 + v01) + (v10 + v11)) * 0.25f;
  }
}

// Host driver.  d_a / d_b: device uint8 [B,H,W,C].  ws: device scratch of at least msssim_ws_bytes().  stats_host: [B][5][C][2].
inline size_t msssim_ws_bytes(int B, int H, int W, int C) {
  size_t pooled = 0;
  int h = H, w = W;
  for (int k = 1; k < MS_SCALES; ++k) { h = (h + 1) / 2; w = (w + 1) / 2; pooled += (size_t)B * h * w * C * 4 * 2; }
  const size_t tiles = (size_t)((W + MS_TX - 1) / MS_TX) * ((H + MS_TY - 1) / MS_TY);
  return pooled + (size_t)B * C * tiles * 2 * 8 + (size_t)B * MS_SCALES * C * 2 * 8 + 1024;
}

inline bool msssim_single_scale(int H, int W) { return H < 160 && W < 160; }   // mshyper/models.py:325-327

inline int msssim_run(const uint8_t* d_a, const uint8_t* d_b, int B, int H, int W, int C, uint8_t* ws, double* d_stats_out_host, cudaStream_t s,
                      uint64_t* launches, std::string* err) {
  const int scales = msssim_single_scale(H, W) ? 1 : MS_SCALES;
  {
    int h = H, w = W;
    for (int k = 1; k < scales; ++k) { h = (h + 1) / 2; w = (w + 1) / 2; }
    if (h < MS_K || w < MS_K) { *err = "msssim: image too small for " + std::to_string(scales) + " scale(s) of an 11x11 window"; return 1; }
  }
  MsScaleParams P{};
  {
    double g[MS_K], sum = 0.0;
    for (int i = 0; i < MS_K; ++i) { const double c = i - (MS_K - 1) / 2.0; g[i] = std::exp(-0.5 * c * c / (1.5 * 1.5)); sum += g[i]; }
    for (int i = 0; i < MS_K; ++i) P.g[i] = (float)(g[i] / sum);
  }
  // scratch layout: pooled image pairs of scales 1..4 | tile partials | stats
  size_t off = 0;
  float* lvl_a[MS_SCALES] = {nullptr}; float* lvl_b[MS_SCALES] = {nullptr};
  {
    int h = H, w = W;
    for (int k = 1; k < MS_SCALES; ++k) {
      h = (h + 1) / 2; w = (w + 1) / 2;
      lvl_a[k] = reinterpret_cast<float*>(ws + off); off += (size_t)B * h * w * C * 4;
      lvl_b[k] = reinterpret_cast<float*>(ws + off); off += (size_t)B * h * w * C * 4;
    }
  }
  off = (off + 255) / 256 * 256;
  double* partial = reinterpret_cast<double*>(ws + off);
  off += (size_t)B * C * ((size_t)((W + MS_TX - 1) / MS_TX) * ((H + MS_TY - 1) / MS_TY)) * 2 * 8;
  double* stats = reinterpret_cast<double*>(ws + off);
  const void* ca = d_a; const void* cb = d_b;
  int is_u8 = 1, h = H, w = W;
  for (int k = 0; k < scales; ++k) {
    if (k > 0) {
      const int h2 = (h + 1) / 2, w2 = (w + 1) / 2;
      const size_t n = (size_t)B * h2 * w2 * C;
      const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
      msssim_pool_kernel<<<grid, 256, 0, s>>>(ca, is_u8, lvl_a[k], B, h, w, C);
      msssim_pool_kernel<<<grid, 256, 0, s>>>(cb, is_u8, lvl_b[k], B, h, w, C);
      if (launches) *launches += 2;
      ca = lvl_a[k]; cb = lvl_b[k]; is_u8 = 0; h = h2; w = w2;
    }
    P.a = ca; P.b = cb; P.is_u8 = is_u8; P.B = B; P.H = h; P.W = w; P.C = C;
    P.tiles_x = (w - (MS_K - 1) + MS_TX - 1) / MS_TX; P.tiles_y = (h - (MS_K - 1) + MS_TY - 1) / MS_TY;
    P.partial = partial;
    const int tiles = P.tiles_x * P.tiles_y;
    msssim_scale_kernel<<<dim3((unsigned)tiles, (unsigned)C, (unsigned)B), 256, 0, s>>>(P);
    const double count = (double)(h - (MS_K - 1)) * (double)(w - (MS_K - 1));
    msssim_finalize_kernel<<<dim3((unsigned)C, (unsigned)B), 256, 0, s>>>(partial, tiles, C, k, 1.0 / count, stats);
    if (launches) *launches += 2;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_stats_out_host, stats, (size_t)B * MS_SCALES * C * 2 * 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { *err = std::string("msssim: ") + cudaGetErrorString(e); return 2; }
  return 0;
}

// stats [B][5][C][2] (means of lum*cs and cs per scale and channel) -> MS-SSIM per image (tf.image.ssim_multiscale's
// relu / weighted geometric mean over scales / mean over channels; single-scale: mean over channels of ssim)
inline void msssim_combine(const double* stats, int B, int C, bool single, double* out) {
  static const double wts[MS_SCALES] = {0.0448, 0.2856, 0.3001, 0.2363, 0.1333};
  for (int b = 0; b < B; ++b) {
    double acc = 0.0;
    for (int c = 0; c < C; ++c) {
      auto st = [&](int k, int which) { return stats[(((size_t)b * MS_SCALES + k) * C + c) * 2 + which]; };
      if (single) { acc += st(0, 0); continue; }
      double prod = 1.0;
      for (int k = 0; k < MS_SCALES; ++k) {
        const double f = std::max(k + 1 < MS_SCALES ? st(k, 1) : st(k, 0), 0.0);
        prod *= std::pow(f, wts[k]);
      }
      acc += prod;
    }
    out[b] = acc / C;
  }
}

}  // namespace sntc
