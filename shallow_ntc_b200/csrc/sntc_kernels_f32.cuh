// fp32 CUDA-core kernels of the decode path (SNTC_PRECISION_FP32) and the pointwise / epilogue
// kernels shared with the tensor-core path.  All tensors NHWC.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace sntc {

enum { A_NONE = 0, A_ABS = 1, A_SQUARE = 2 };
// gdn epilogue modes of the band GEMM: out = combine(gx, acc + beta)
enum { G_NONE = 0, G_MUL = 1, G_MUL_SQRT = 2, G_DIV = 3, G_DIV_SQRT = 4 };

struct BandGemmParams {
  const float* x; int B, hin, win, cin;       // input [B,hin,win,cin] (cin multiple of 4)
  int a_transform;                              // A_* applied to the A operand on load
  const float* w; int K, N, Npad;               // band matrix [K][Npad]
  const float* bias; int cout;                  // bias[cout]; columns are (fy, fx, co)
  int s, p, phy0, nphx, phx0, Ty, Tx;
  int mloy, cnty, mlox, cntx;                   // cell ranges of this band
  int act;                                      // SNTC_ACT_NONE / RELU / LEAKY_RELU
  float* out; int hout, wout, cstride;          // f32 destination [B,hout,wout,cstride]
  uint8_t* out_u8; float* out_crop; int H, W;   // cropped [B,H,W,cout] uint8 / f32 destinations (final layer)
  const float* gx; int gdn_mode;                // G_*: gx has the layout of `out`
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == SNTC_ACT_RELU) return fmaxf(v, 0.f);
  if (act == SNTC_ACT_LEAKY_RELU) return v >= 0.f ? v : v * 0.2f;
  return v;
}

// data_lib.floats_to_pixels(training=False): saturate_cast_u8(round_half_even((x + 0.5) * 255))
__device__ __forceinline__ uint8_t float_to_pixel(float x) {
  float v = __fmul_rn(__fadd_rn(x, 0.5f), 255.f);   // add first, then multiply; no FMA contraction
  v = rintf(v);
  v = fminf(fmaxf(v, 0.f), 255.f);                  // NaN -> 0 via fmaxf
  return (uint8_t)v;
}

__device__ __forceinline__ float a_xform(float v, int t) {
  return t == A_ABS ? fabsf(v) : (t == A_SQUARE ? v * v : v);
}

// Band GEMM on CUDA cores: C[M cells, N] = A_gathered[M, K] * Wb[K, N], register-tiled, double-buffered.
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
band_gemm_f32_kernel(const BandGemmParams P) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int A_V = BM * BK / 4, B_V = BK * BN / 4;      // float4 loads per tile
  constexpr int A_LD = (A_V + NT - 1) / NT;
  constexpr int B_LD = (B_V + NT - 1) / NT;
  static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 register tiles");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = P.B * P.cnty * P.cntx;

  // ---- A load assignment ----
  int a_row[A_LD], a_kq[A_LD], a_my[A_LD], a_mx[A_LD];
  const float* a_img[A_LD];
  bool a_ok[A_LD];
#pragma unroll
  for (int i = 0; i < A_LD; ++i) {
    int idx = tid + i * NT;
    a_row[i] = idx / (BK / 4);
    a_kq[i] = (idx % (BK / 4)) * 4;
    int m = m0 + a_row[i];
    a_ok[i] = m < M && idx < A_V;
    int mm = a_ok[i] ? m : 0;
    int ix = mm % P.cntx; int t = mm / P.cntx; int iy = t % P.cnty; int b = t / P.cnty;
    a_my[i] = P.mloy + iy; a_mx[i] = P.mlox + ix;
    a_img[i] = P.x + (size_t)b * P.hin * P.win * P.cin;
  }
  int b_row[B_LD], b_nq[B_LD];
  bool b_ok[B_LD];
#pragma unroll
  for (int i = 0; i < B_LD; ++i) {
    int idx = tid + i * NT;
    b_ok[i] = idx < B_V;
    b_row[i] = b_ok[i] ? idx / (BN / 4) : 0;
    b_nq[i] = b_ok[i] ? (idx % (BN / 4)) * 4 : 0;
  }

  const int cblocks = (P.cin + BK - 1) / BK;
  const int steps = P.Ty * P.Tx * cblocks;
  float4 ra[A_LD], rb[B_LD];

  auto load_step = [&](int step) {
    int tap = step / cblocks, c0 = (step - tap * cblocks) * BK;
    int jy = tap / P.Tx, jx = tap - jy * P.Tx;
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int ny = a_my[i] - jy, nx = a_mx[i] - jx, c = c0 + a_kq[i];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[i] && ny >= 0 && ny < P.hin && nx >= 0 && nx < P.win && c < P.cin)
        v = __ldg(reinterpret_cast<const float4*>(a_img[i] + ((size_t)ny * P.win + nx) * P.cin + c));
      if (P.a_transform != A_NONE) {
        v.x = a_xform(v.x, P.a_transform); v.y = a_xform(v.y, P.a_transform);
        v.z = a_xform(v.z, P.a_transform); v.w = a_xform(v.w, P.a_transform);
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int c = c0 + b_row[i], n = n0 + b_nq[i];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b_ok[i] && c < P.cin && n < P.Npad)
        v = __ldg(reinterpret_cast<const float4*>(P.w + ((size_t)tap * P.cin + c) * P.Npad + n));
      rb[i] = v;
    }
  };
  auto store_step = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      if (tid + i * NT >= A_V) continue;
      As[buf][a_kq[i] + 0][a_row[i]] = ra[i].x;
      As[buf][a_kq[i] + 1][a_row[i]] = ra[i].y;
      As[buf][a_kq[i] + 2][a_row[i]] = ra[i].z;
      As[buf][a_kq[i] + 3][a_row[i]] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i)
      if (b_ok[i]) *reinterpret_cast<float4*>(&Bs[buf][b_row[i]][b_nq[i]]) = rb[i];
  };

  const int ty = tid / (BN / TN), tx = tid % (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (steps > 0) {
    load_step(0);
    store_step(0);
  }
  __syncthreads();
  for (int step = 0; step < steps; ++step) {
    int cur = step & 1;
    if (step + 1 < steps) load_step(step + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (step + 1 < steps) store_step(cur ^ 1);
    __syncthreads();
  }

  // ---- epilogue: column n -> (fy, fx, co); row m -> cell; one store per output element ----
  int c_co[TN], c_phy[TN], c_phx[TN];
  bool c_ok[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int n = n0 + tx * TN + j;
    c_ok[j] = n < P.N;
    int nn = c_ok[j] ? n : 0;
    c_co[j] = nn % P.cout; int ph = nn / P.cout;
    c_phx[j] = P.phx0 + ph % P.nphx; c_phy[j] = P.phy0 + ph / P.nphx;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
    int ix = m % P.cntx; int t = m / P.cntx; int iy = t % P.cnty; int b = t / P.cnty;
    int my = P.mloy + iy, mx = P.mlox + ix;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      if (!c_ok[j]) continue;
      int oy = P.s * my + c_phy[j] - P.p, ox = P.s * mx + c_phx[j] - P.p;
      if (oy < 0 || oy >= P.hout || ox < 0 || ox >= P.wout) continue;
      float v = acc[i][j] + (P.bias ? __ldg(P.bias + c_co[j]) : 0.f);
      v = apply_act(v, P.act);
      size_t pix = ((size_t)b * P.hout + oy) * P.wout + ox;
      if (P.gdn_mode != G_NONE) {
        float g = __ldg(P.gx + pix * P.cstride + c_co[j]);
        if (P.gdn_mode == G_MUL) v = g * v;
        else if (P.gdn_mode == G_MUL_SQRT) v = g * sqrtf(v);
        else if (P.gdn_mode == G_DIV) v = g / v;
        else v = g / sqrtf(v);
      }
      if (P.out) P.out[pix * P.cstride + c_co[j]] = v;
      if ((P.out_u8 || P.out_crop) && oy < P.H && ox < P.W) {
        size_t q = (((size_t)b * P.H + oy) * P.W + ox) * P.cout + c_co[j];
        if (P.out_u8) P.out_u8[q] = float_to_pixel(v);
        if (P.out_crop) P.out_crop[q] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Final transposed conv to <= 4 channels (two-layer tail k5s2, mbt2018 / bls2017 / cnn last layer):
// one thread per cell computes all S*S phases; weights are warp-uniform smem broadcasts.
struct RgbCellParams {
  const float* x; int B, hin, win, cin;   // cin multiple of 4 (cin_pad)
  const float* w;                          // [k*k][cin][4]
  const float* bias; int cout, k, s, p;
  int cnty, cntx;                          // cells m in [0,cnt)
  int cc;                                  // cin chunk held in smem
  float* out; int hout, wout;              // optional f32 [B,hout,wout,cout]
  uint8_t* out_u8; float* out_crop; int H, W;
};

template <int S>
__global__ void __launch_bounds__(128) convt_rgb_cell_kernel(const RgbCellParams P) {
  extern __shared__ __align__(16) float sw[];   // [k*k][cc][4]
  const int cells = P.cnty * P.cntx;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const bool live = cell < cells;
  const int my = live ? cell / P.cntx : 0, mx = live ? cell % P.cntx : 0;
  const int T = (P.k + S - 1) / S;
  float acc[S][S][3];
#pragma unroll
  for (int a = 0; a < S; ++a)
#pragma unroll
    for (int c = 0; c < S; ++c) { acc[a][c][0] = 0.f; acc[a][c][1] = 0.f; acc[a][c][2] = 0.f; }
  const float* img = P.x + (size_t)b * P.hin * P.win * P.cin;
  const int kk2 = P.k * P.k;

  for (int c0 = 0; c0 < P.cin; c0 += P.cc) {
    const int cn = min(P.cc, P.cin - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < kk2 * cn; i += blockDim.x) {
      int tap = i / cn, ci = i - tap * cn;
      *reinterpret_cast<float4*>(&sw[((size_t)tap * P.cc + ci) * 4]) =
        __ldg(reinterpret_cast<const float4*>(P.w + ((size_t)tap * P.cin + c0 + ci) * 4));
    }
    __syncthreads();
    if (!live) continue;
    for (int jy = 0; jy < T; ++jy) {
      int ny = my - jy;
      if (ny < 0 || ny >= P.hin) continue;
      for (int jx = 0; jx < T; ++jx) {
        int nx = mx - jx;
        if (nx < 0 || nx >= P.win) continue;
        const float* px = img + ((size_t)ny * P.win + nx) * P.cin + c0;
        for (int ci = 0; ci < cn; ci += 4) {
          float4 xv = __ldg(reinterpret_cast<const float4*>(px + ci));
          float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int fy = 0; fy < S; ++fy) {
            int ay = fy + S * jy;
            if (ay >= P.k) continue;
#pragma unroll
            for (int fx = 0; fx < S; ++fx) {
              int ax = fx + S * jx;
              if (ax >= P.k) continue;
              const float* wp = &sw[((size_t)(ay * P.k + ax) * P.cc + ci) * 4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float4 wv = *reinterpret_cast<const float4*>(wp + u * 4);
                acc[fy][fx][0] = fmaf(xs[u], wv.x, acc[fy][fx][0]);
                acc[fy][fx][1] = fmaf(xs[u], wv.y, acc[fy][fx][1]);
                acc[fy][fx][2] = fmaf(xs[u], wv.z, acc[fy][fx][2]);
              }
            }
          }
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int fy = 0; fy < S; ++fy) {
    int oy = S * my + fy - P.p;
    if (oy < 0 || oy >= P.hout) continue;
#pragma unroll
    for (int fx = 0; fx < S; ++fx) {
      int ox = S * mx + fx - P.p;
      if (ox < 0 || ox >= P.wout) continue;
      for (int co = 0; co < P.cout; ++co) {
        float v = acc[fy][fx][co] + (P.bias ? __ldg(P.bias + co) : 0.f);
        if (P.out) P.out[(((size_t)b * P.hout + oy) * P.wout + ox) * P.cout + co] = v;
        if (oy < P.H && ox < P.W) {
          size_t q = (((size_t)b * P.H + oy) * P.W + ox) * P.cout + co;
          if (P.out_u8) P.out_u8[q] = float_to_pixel(v);
          if (P.out_crop) P.out_crop[q] = v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tail of the two-layer synthesis: stride-2 transposed conv C1 -> 3 with the fused crop + uint8 epilogue
// (common/transforms.py:311-313, 350-353 + image_utils.py:22-23, 69-71).  HBM-class: reads t once, writes the
// image once.  One thread = 4 (y) x 2 (x) output pixels; the input tile and the weights sit in shared
// memory; every tap index is a compile-time constant (a = (P&1) + d + 2(T-1) - 2i).
struct TailParams {
  const float* x; int B, hin, win;          // t [B,hin,win,C1]
  const float* w;                            // [K*K][C1][4]
  const float* bias;
  float* out; int hout, wout;                // optional f32 [B,hout,wout,3]
  uint8_t* out_u8; float* out_crop; int H, W;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int K, int P, int C1, int TR>
__global__ void __launch_bounds__(32 * TR) tail_s2_kernel(const TailParams Q) {
  constexpr int T = (K + 1) / 2;
  constexpr int RY = 8, TCX = 32;                                  // one thread = 8 (y) x 2 (x) output pixels; 32 x TR threads
  constexpr int NT = TCX * TR;
  constexpr int NY = ((P + RY - 1) >> 1) - (P >> 1) + T, NX = ((P + 1) >> 1) - (P >> 1) + T;
  constexpr int TILE_Y = (RY / 2) * (TR - 1) + NY, TILE_X = (TCX - 1) + NX;
  extern __shared__ __align__(16) float sm_tail[];
  float* sx = sm_tail;                                             // [TILE_Y][TILE_X][C1]
  float* sw = sm_tail + TILE_Y * TILE_X * C1;                      // [K*K][C1][4]
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * (RY * TR), ox0 = blockIdx.x * (2 * TCX);
  const int ny0 = oy0 / 2 + (P >> 1) - (T - 1), nx0 = ox0 / 2 + (P >> 1) - (T - 1);
  const float* img = Q.x + (size_t)b * Q.hin * Q.win * C1;
  // global -> shared with cp.async (16 B per request, zero-filled outside the image): all requests of a
  // thread are in flight together instead of one dependent load/store pair per loop trip
  for (int i = threadIdx.x; i < TILE_Y * TILE_X * (C1 / 4); i += NT) {
    int c4 = i % (C1 / 4), px = i / (C1 / 4);
    int tx = px % TILE_X, ty = px / TILE_X;
    int ny = ny0 + ty, nx = nx0 + tx;
    const bool ok = ny >= 0 && ny < Q.hin && nx >= 0 && nx < Q.win;
    const float* src = ok ? img + ((size_t)ny * Q.win + nx) * C1 + c4 * 4 : img;
    cp_async16(sx + (size_t)px * C1 + c4 * 4, src, ok ? 16 : 0);
  }
  for (int i = threadIdx.x; i < K * K * C1; i += NT) cp_async16(sw + (size_t)i * 4, Q.w + (size_t)i * 4, 16);
  cp_async_wait_all();
  __syncthreads();
  const int ti = threadIdx.x % TCX, tr = threadIdx.x / TCX;
  float acc[RY][2][3];
#pragma unroll
  for (int dy = 0; dy < RY; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) { acc[dy][dx][0] = 0.f; acc[dy][dx][1] = 0.f; acc[dy][dx][2] = 0.f; }
  const float* tbase = sx + ((size_t)((RY / 2) * tr) * TILE_X + ti) * C1;
  // taps outermost: each weight vector feeds the RY/2 output rows of its phase; all indices are compile-time
#pragma unroll
  for (int ay = 0; ay < K; ++ay) {
#pragma unroll
    for (int ax = 0; ax < K; ++ax) {
      const int dx = ((ax - P) % 2 + 2) % 2;                       // the output column of this thread with phase ax
      const int ix = (dx + P - ax) / 2 - (P >> 1) + (T - 1);
      const int dy0 = ((ay - P) % 2 + 2) % 2;
#pragma unroll 1
      for (int c = 0; c < C1; c += 4) {
        const float* wp = sw + ((size_t)(ay * K + ax) * C1 + c) * 4;
        const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
        const float4 w2 = *reinterpret_cast<const float4*>(wp + 8), w3 = *reinterpret_cast<const float4*>(wp + 12);
#pragma unroll
        for (int e = 0; e < RY / 2; ++e) {
          const int dy = dy0 + 2 * e;
          const int iy = (dy + P - ay) / 2 - (P >> 1) + (T - 1);
          const float4 xv = *reinterpret_cast<const float4*>(tbase + ((size_t)iy * TILE_X + ix) * C1 + c);
          float* a = acc[dy][dx];
          a[0] = fmaf(xv.x, w0.x, a[0]); a[1] = fmaf(xv.x, w0.y, a[1]); a[2] = fmaf(xv.x, w0.z, a[2]);
          a[0] = fmaf(xv.y, w1.x, a[0]); a[1] = fmaf(xv.y, w1.y, a[1]); a[2] = fmaf(xv.y, w1.z, a[2]);
          a[0] = fmaf(xv.z, w2.x, a[0]); a[1] = fmaf(xv.z, w2.y, a[1]); a[2] = fmaf(xv.z, w2.z, a[2]);
          a[0] = fmaf(xv.w, w3.x, a[0]); a[1] = fmaf(xv.w, w3.y, a[1]); a[2] = fmaf(xv.w, w3.z, a[2]);
        }
      }
    }
  }
  const float b0 = __ldg(Q.bias), b1 = __ldg(Q.bias + 1), b2 = __ldg(Q.bias + 2);
#pragma unroll
  for (int dy = 0; dy < RY; ++dy) {
    const int oy = oy0 + RY * tr + dy;
    if (oy >= Q.hout) continue;
    const int ox = ox0 + 2 * ti;
    if (ox >= Q.wout) continue;
    float v[2][3];
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) { v[dx][0] = acc[dy][dx][0] + b0; v[dx][1] = acc[dy][dx][1] + b1; v[dx][2] = acc[dy][dx][2] + b2; }
    if (Q.out) {
      float* o = Q.out + (((size_t)b * Q.hout + oy) * Q.wout + ox) * 3;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) if (ox + dx < Q.wout) { o[dx * 3] = v[dx][0]; o[dx * 3 + 1] = v[dx][1]; o[dx * 3 + 2] = v[dx][2]; }
    }
    if (oy < Q.H) {
      if (Q.out_u8) {
        uint8_t* o = Q.out_u8 + (((size_t)b * Q.H + oy) * Q.W + ox) * 3;
        if (ox + 1 < Q.W && ((Q.W * 3) % 2 == 0)) {     // 6 contiguous bytes, 2-byte aligned: three 16-bit stores
          const uint8_t p0 = float_to_pixel(v[0][0]), p1 = float_to_pixel(v[0][1]), p2 = float_to_pixel(v[0][2]);
          const uint8_t p3 = float_to_pixel(v[1][0]), p4 = float_to_pixel(v[1][1]), p5 = float_to_pixel(v[1][2]);
          uint16_t* o16 = reinterpret_cast<uint16_t*>(o);
          o16[0] = (uint16_t)(p0 | (p1 << 8)); o16[1] = (uint16_t)(p2 | (p3 << 8)); o16[2] = (uint16_t)(p4 | (p5 << 8));
        } else {
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) if (ox + dx < Q.W) { o[dx * 3] = float_to_pixel(v[dx][0]); o[dx * 3 + 1] = float_to_pixel(v[dx][1]); o[dx * 3 + 2] = float_to_pixel(v[dx][2]); }
        }
      }
      if (Q.out_crop) {
        float* o = Q.out_crop + (((size_t)b * Q.H + oy) * Q.W + ox) * 3;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) if (ox + dx < Q.W) { o[dx * 3] = v[dx][0]; o[dx * 3 + 1] = v[dx][1]; o[dx * 3 + 2] = v[dx][2]; }
      }
    }
  }
}

// Same computation with the weights passed BY VALUE as a kernel parameter: after full unrolling every weight
// is a constant-bank operand of its FFMA (no shared-memory weight traffic, the kernel becomes FMA-issue bound).
// Fits the 4 KB parameter space for K=5, C1=12 (3600 B): the headline two_layer_syn configuration.
template <int K, int C1>
struct TailWeights { float w[K * K * C1 * 3]; float bias[3]; };   // [ay*K+ax][ci][co]

template <int K, int P, int C1, int TR>
__global__ void __launch_bounds__(32 * TR, 4) tail_s2_const_kernel(const TailParams Q, const __grid_constant__ TailWeights<K, C1> Wt) {
  constexpr int T = (K + 1) / 2;
  constexpr int RY = 8, TCX = 32;
  constexpr int NT = TCX * TR;
  constexpr int NY = ((P + RY - 1) >> 1) - (P >> 1) + T, NX = ((P + 1) >> 1) - (P >> 1) + T;
  constexpr int TILE_Y = (RY / 2) * (TR - 1) + NY, TILE_X = (TCX - 1) + NX;
  __shared__ __align__(16) float sx[TILE_Y * TILE_X * C1];
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent of the layer-1 kernel
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * (RY * TR), ox0 = blockIdx.x * (2 * TCX);
  const int ny0 = oy0 / 2 + (P >> 1) - (T - 1), nx0 = ox0 / 2 + (P >> 1) - (T - 1);
  const float* img = Q.x + (size_t)b * Q.hin * Q.win * C1;
  for (int i = threadIdx.x; i < TILE_Y * TILE_X * (C1 / 4); i += NT) {
    int c4 = i % (C1 / 4), px = i / (C1 / 4);
    int tx = px % TILE_X, ty = px / TILE_X;
    int ny = ny0 + ty, nx = nx0 + tx;
    const bool ok = ny >= 0 && ny < Q.hin && nx >= 0 && nx < Q.win;
    const float* src = ok ? img + ((size_t)ny * Q.win + nx) * C1 + c4 * 4 : img;
    cp_async16(sx + (size_t)px * C1 + c4 * 4, src, ok ? 16 : 0);
  }
  cp_async_wait_all();
  __syncthreads();
  const int ti = threadIdx.x % TCX, tr = threadIdx.x / TCX;
  float acc[RY][2][3];
#pragma unroll
  for (int dy = 0; dy < RY; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) { acc[dy][dx][0] = 0.f; acc[dy][dx][1] = 0.f; acc[dy][dx][2] = 0.f; }
  const float* tbase = sx + ((size_t)((RY / 2) * tr) * TILE_X + ti) * C1;
#pragma unroll
  for (int ay = 0; ay < K; ++ay) {
#pragma unroll
    for (int ax = 0; ax < K; ++ax) {
      const int dx = ((ax - P) % 2 + 2) % 2;
      const int ix = (dx + P - ax) / 2 - (P >> 1) + (T - 1);
      const int dy0 = ((ay - P) % 2 + 2) % 2;
#pragma unroll
      for (int c = 0; c < C1; c += 4) {
#pragma unroll
        for (int e = 0; e < RY / 2; ++e) {
          const int dy = dy0 + 2 * e;
          const int iy = (dy + P - ay) / 2 - (P >> 1) + (T - 1);
          const float4 xv = *reinterpret_cast<const float4*>(tbase + ((size_t)iy * TILE_X + ix) * C1 + c);
          const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
          float* a = acc[dy][dx];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float* w = Wt.w + ((ay * K + ax) * C1 + c + u) * 3;
            a[0] = fmaf(xs[u], w[0], a[0]); a[1] = fmaf(xs[u], w[1], a[1]); a[2] = fmaf(xs[u], w[2], a[2]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int dy = 0; dy < RY; ++dy) {
    const int oy = oy0 + RY * tr + dy;
    if (oy >= Q.hout) continue;
    const int ox = ox0 + 2 * ti;
    if (ox >= Q.wout) continue;
    float v[2][3];
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) { v[dx][0] = acc[dy][dx][0] + Wt.bias[0]; v[dx][1] = acc[dy][dx][1] + Wt.bias[1]; v[dx][2] = acc[dy][dx][2] + Wt.bias[2]; }
    if (Q.out) {
      float* o = Q.out + (((size_t)b * Q.hout + oy) * Q.wout + ox) * 3;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) if (ox + dx < Q.wout) { o[dx * 3] = v[dx][0]; o[dx * 3 + 1] = v[dx][1]; o[dx * 3 + 2] = v[dx][2]; }
    }
    if (oy < Q.H) {
      if (Q.out_u8) {
        uint8_t* o = Q.out_u8 + (((size_t)b * Q.H + oy) * Q.W + ox) * 3;
        if (ox + 1 < Q.W && ((Q.W * 3) % 2 == 0)) {
          const uint8_t p0 = float_to_pixel(v[0][0]), p1 = float_to_pixel(v[0][1]), p2 = float_to_pixel(v[0][2]);
          const uint8_t p3 = float_to_pixel(v[1][0]), p4 = float_to_pixel(v[1][1]), p5 = float_to_pixel(v[1][2]);
          uint16_t* o16 = reinterpret_cast<uint16_t*>(o);
          o16[0] = (uint16_t)(p0 | (p1 << 8)); o16[1] = (uint16_t)(p2 | (p3 << 8)); o16[2] = (uint16_t)(p4 | (p5 << 8));
        } else {
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) if (ox + dx < Q.W) { o[dx * 3] = float_to_pixel(v[dx][0]); o[dx * 3 + 1] = float_to_pixel(v[dx][1]); o[dx * 3 + 2] = float_to_pixel(v[dx][2]); }
        }
      }
      if (Q.out_crop) {
        float* o = Q.out_crop + (((size_t)b * Q.H + oy) * Q.W + ox) * 3;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) if (ox + dx < Q.W) { o[dx * 3] = v[dx][0]; o[dx * 3 + 1] = v[dx][1]; o[dx * 3 + 2] = v[dx][2]; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// mshyper/models.py:274-279: split, exp, clamp, round -> idx ; y_hat = q + mu.
// hs: [P, 2*C] (mu || raw_sigma).  q may be f32 / i16 / i8.  One thread per 4 channels.
struct DequantParams {
  const float* hs; const void* q; int q_kind;   // 0 f32, 1 i16, 2 i8
  size_t npix; int C; float max_index; int trunc;
  float* y_hat; uint8_t* idx;
};

__device__ __forceinline__ float4 load_q4(const void* q, int kind, size_t e) {
  if (kind == 3) return make_float4(0.f, 0.f, 0.f, 0.f);   // no symbols yet (phase 1 of the two-phase decode): y_hat = mu
  if (kind == 0) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q) + e));
  if (kind == 1) {
    short4 v = *reinterpret_cast<const short4*>(reinterpret_cast<const int16_t*>(q) + e);
    return make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
  }
  char4 v = *reinterpret_cast<const char4*>(reinterpret_cast<const int8_t*>(q) + e);
  return make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
}

__device__ __forceinline__ uint8_t scale_index(float raw, float max_index, int trunc) {
  float i_f = expf(raw);                                  // sigma = tf.exp(sigma)   :275
  float i_c = fminf(fmaxf(i_f, 0.f), max_index);          // _normalize_indexes clamp to [0, S-1]
  return (uint8_t)(trunc ? floorf(i_c) : rintf(i_c));     // table row (round-half-even)
}

__global__ void dequant_index_kernel(const DequantParams P) {
  const int c4 = P.C / 4;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.npix * c4) return;
  size_t pix = i / c4; int c = (int)(i - pix * c4) * 4;
  const float* row = P.hs + pix * 2 * P.C;
  float4 mu = __ldg(reinterpret_cast<const float4*>(row + c));
  size_t e = pix * P.C + c;
  if (P.y_hat) {
    float4 q = load_q4(P.q, P.q_kind, e);
    float4 y = make_float4(__fadd_rn(q.x, mu.x), __fadd_rn(q.y, mu.y), __fadd_rn(q.z, mu.z), __fadd_rn(q.w, mu.w));
    *reinterpret_cast<float4*>(P.y_hat + e) = y;
  }
  if (P.idx) {
    float4 sg = __ldg(reinterpret_cast<const float4*>(row + P.C + c));
    uchar4 o;
    o.x = scale_index(sg.x, P.max_index, P.trunc); o.y = scale_index(sg.y, P.max_index, P.trunc);
    o.z = scale_index(sg.z, P.max_index, P.trunc); o.w = scale_index(sg.w, P.max_index, P.trunc);
    *reinterpret_cast<uchar4*>(P.idx + e) = o;
  }
}

// Phase 2 of the two-phase decode: y_hat = q + mu (mshyper/models.py:278) from the mu kept by phase 1; writes fp32
// and/or the fp16 hi/lo planes the tensor-core synthesis reads.  One thread per 8 elements.
__global__ void dequant_planes_kernel(const float* __restrict__ mu, const void* __restrict__ q, int kind, size_t n,
                                      float* __restrict__ y_hat, __half* __restrict__ hi, __half* __restrict__ lo) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t e = i * 8;
  if (e >= n) return;
  const bool full = e + 8 <= n;   // n % 4 == 0: the last thread may hold 4 elements only (fp32 output only, see the caller)
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 m0 = __ldg(reinterpret_cast<const float4*>(mu + e)), m1 = full ? __ldg(reinterpret_cast<const float4*>(mu + e + 4)) : z4;
  const float4 q0 = load_q4(q, kind, e), q1 = full ? load_q4(q, kind, e + 4) : z4;
  const float y[8] = {__fadd_rn(q0.x, m0.x), __fadd_rn(q0.y, m0.y), __fadd_rn(q0.z, m0.z), __fadd_rn(q0.w, m0.w),
                      __fadd_rn(q1.x, m1.x), __fadd_rn(q1.y, m1.y), __fadd_rn(q1.z, m1.z), __fadd_rn(q1.w, m1.w)};
  if (y_hat) {
    *reinterpret_cast<float4*>(y_hat + e) = make_float4(y[0], y[1], y[2], y[3]);
    if (full) *reinterpret_cast<float4*>(y_hat + e + 4) = make_float4(y[4], y[5], y[6], y[7]);
  }
  if (hi && full) {
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float c = fminf(fmaxf(y[k], -65504.f), 65504.f);
      h[k] = __float2half_rn(c);
      l[k] = __float2half_rn(c - __half2float(h[k]));
    }
    *reinterpret_cast<uint4*>(hi + e) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + e) = *reinterpret_cast<const uint4*>(l);
  }
}

// q (f32 / i16 / i8) -> f32, for the factorized model (y_hat = q)
__global__ void convert_q_kernel(const void* q, int kind, float* out, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  *reinterpret_cast<float4*>(out + i * 4) = load_q4(q, kind, i * 4);
}

// ---------------------------------------------------------------------------------------------
// Pointwise activation over C channels with small C (two-layer hidden width): optional GDN1 and
// optional residual: in [P, Cin_stride] with base = in[:, :C], res = in[:, C:2C] -> out [P, C].
struct ActResParams {
  const float* in; int in_stride; float* out; size_t npix; int C;
  int act; int has_res;
  const float* beta; const float* gamma; int gamma_stride; int inverse;  // gamma [in][out]
  // has_res == 2 (TwoLayerResSynthesis res_type="d2s", transforms.py:339-348): the residual is the last depth_to_space(2) of
  // res_ext [B, H1/2, W1/2, 4C]; the pixels of `in` are [B, H1, W1]
  const float* res_ext; int H1, W1;
};

__global__ void __launch_bounds__(128) act_res_kernel(const ActResParams P) {
  extern __shared__ float sm[];                 // gamma [C][C] | beta [C] | x tile [128][C+1]
  float* sg = sm; float* sb = sg + P.C * P.C; float* sx = sb + P.C;
  const bool gdn = P.act == SNTC_ACT_IGDN1 || P.act == SNTC_ACT_GDN1;
  if (gdn) {
    for (int i = threadIdx.x; i < P.C * P.C; i += blockDim.x) sg[i] = P.gamma[(size_t)(i / P.C) * P.gamma_stride + (i % P.C)];
    for (int i = threadIdx.x; i < P.C; i += blockDim.x) sb[i] = P.beta[i];
  }
  size_t pix0 = (size_t)blockIdx.x * blockDim.x;
  int npx = (int)min((size_t)blockDim.x, P.npix - pix0);
  // coalesced load of the base part of the tile
  for (int i = threadIdx.x; i < npx * P.C; i += blockDim.x) {
    int r = i / P.C, c = i - r * P.C;
    sx[r * (P.C + 1) + c] = P.in[(pix0 + r) * P.in_stride + c];
  }
  __syncthreads();
  if ((int)threadIdx.x >= npx) return;
  const float* xr = sx + threadIdx.x * (P.C + 1);
  size_t pix = pix0 + threadIdx.x;
  for (int j = 0; j < P.C; ++j) {
    float x = xr[j], v;
    if (gdn) {
      float norm = sb[j];
      for (int i = 0; i < P.C; ++i) norm = fmaf(fabsf(xr[i]), sg[i * P.C + j], norm);
      v = P.inverse ? x * norm : x / norm;
    } else {
      v = apply_act(x, P.act);
    }
    if (P.has_res == 1) v += P.in[pix * P.in_stride + P.C + j];
    else if (P.has_res == 2) {
      const int x = (int)(pix % P.W1), y = (int)((pix / P.W1) % P.H1); const size_t b = pix / ((size_t)P.W1 * P.H1);
      v += P.res_ext[((b * (P.H1 / 2) + y / 2) * (P.W1 / 2) + x / 2) * (4 * P.C) + (2 * (y & 1) + (x & 1)) * P.C + j];
    }
    P.out[pix * P.C + j] = v;
  }
}

// JPEGLikeSynthesis(use_offset=True): x -> [x, 1, 0...] padded to cpad channels
__global__ void append_ones_kernel(const float* in, int cin, float* out, int cpad, size_t npix) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * cpad) return;
  size_t pix = i / cpad; int c = (int)(i - pix * cpad);
  out[i] = c < cin ? in[pix * cin + c] : (c == cin ? 1.f : 0.f);
}

// ---------------------------------------------------------------------------------------------
// image_utils.mse_psnr numerator: exact integer sum of squared uint8 differences per image.
__global__ void __launch_bounds__(256) ssd_kernel(const uint8_t* a, const uint8_t* b, size_t per_image,
                                                  unsigned long long* ssd) {
  const int img = blockIdx.y;
  const uint8_t* pa = a + (size_t)img * per_image;
  const uint8_t* pb = b + (size_t)img * per_image;
  unsigned long long s = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (size_t)gridDim.x * blockDim.x) {
    int d = (int)pa[i] - (int)pb[i];
    s += (unsigned)(d * d);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  __shared__ unsigned long long ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < 8; ++i) t += ws[i];
    atomicAdd(ssd + img, t);
  }
}

}  // namespace sntc
