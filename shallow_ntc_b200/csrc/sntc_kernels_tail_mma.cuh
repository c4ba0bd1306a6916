// Warp-MMA tail of the two-layer synthesis: ConvT(5, 2, C1 -> 3, p = 1) + crop + uint8 epilogue
// (common/transforms.py:311-313, 350-353 + image_utils.py:22-23, 69-71 + data_lib.py:48-52), C1 <= 16.
//
// Why not tcgen05 here: the op has 900 useful MACs per t-pixel and N = 12 output columns; a 128-row UMMA is then
// bound by the delivery of its A operand from shared memory (sntc_kernels_tail_tc.cuh: ties with the FFMA kernel).
// Register-operand warp MMAs (mma.sync m16n8k16) do not have that floor and ldmatrix gives the im2col for free:
//     D[16 t-pixels of one row, (phx, co)] += A_tap[16 pixels shifted by (dy, jx), 16 channels] * W_tap[16, 8]
// per output-row parity phy (o = 2n + a - 1  ->  a = ph + 1 - 2d:  even rows use d in {-1,0}, odd rows d in {-1,0,+1}).
//   * a CTA (4 warps) stages an (8+2) x (64+2) halo tile of t: fp32 from global -> split x = hi + lo (two fp16) ->
//     shared memory, octet-planar [plane][octet][row][x][8 ch] (channels zero-padded to 16), so the 8 rows of an
//     ldmatrix 8x8 block are 8 consecutive pixels = 128 contiguous bytes (conflict-free for every tap shift);
//   * a warp owns a 16-pixel-wide strip and walks down its 8 rows keeping the A fragments of the three live input
//     rows in registers (3 rows x 3 x-shifts x hi/lo): each step loads ONE new row (6 ldmatrix.x4) and issues 45 MMAs
//     (15 (parity, tap) pairs x {hi*hi, lo*hi, hi*lo}; lo*lo < 2^-22 relative is dropped as everywhere);
//   * the 15 weight fragments (hi/lo) live in registers for the whole kernel;
//   * epilogue: scale, bias, crop, sat_u8(rint((x + .5) * 255)); a thread holds two adjacent bytes of the image row.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cmath>
#include <string>
#include <vector>
#include "sntc_plan.hpp"
#include "sntc_kernels_f32.cuh"
#include "sntc_kernels_tc.cuh"

namespace sntc {

#ifndef SNTC_TAIL_TX
#define SNTC_TAIL_TX 64
#endif
constexpr int TM_TX = SNTC_TAIL_TX, TM_RW = 8;    // t-pixels per CTA: 8 rows x TM_TX columns; one warp = one 16-column strip
constexpr int TM_TXH = TM_TX + 2, TM_TYH = TM_RW + 2;
constexpr int TM_NFRAG = 15;                      // (parity, tap): phy=0 -> dy in {-1,0}; phy=1 -> dy in {-1,0,+1}; jx in {-1,0,+1}
constexpr int TM_THREADS = 32 * (TM_TX / 16);

struct TailMmaParams {
  const float* x; int B, hin, win;               // t [B,hin,win,C1] fp32
  const uint2* wfrag;                            // [TM_NFRAG][2 planes][32 lanes]: B fragments of mma.m16n8k16 (fp16 pairs)
  float inv_scale; float bias[3];
  float* out; int hout, wout;                    // optional f32 [B,hout,wout,3]
  uint8_t* out_u8; float* out_crop; int H, W;    // cropped pixels / floats [B,H,W,3]
};

__device__ __forceinline__ void tm_mma16816(float* d, const uint32_t* a, const uint32_t* b) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void tm_ldmatrix_x4(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void tm_mma16816_first(float* d, const uint32_t* a, const uint32_t* b) {   // D = A * B (C = 0)
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.f));
}
__device__ __forceinline__ uint32_t tm_h2_bits(const __half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }
// x = hi + lo for 4 values: hi = fp16(x) (saturated to the finite range), lo = fp16(x - hi); packed as 4 x fp16 each
__device__ __forceinline__ void tm_split4(const float4 v, uint2& hi, uint2& lo) {
  const float c0 = fminf(fmaxf(v.x, -65504.f), 65504.f), c1 = fminf(fmaxf(v.y, -65504.f), 65504.f);
  const float c2 = fminf(fmaxf(v.z, -65504.f), 65504.f), c3 = fminf(fmaxf(v.w, -65504.f), 65504.f);
  const __half2 h01 = __floats2half2_rn(c0, c1), h23 = __floats2half2_rn(c2, c3);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(c0 - f01.x, c1 - f01.y), l23 = __floats2half2_rn(c2 - f23.x, c3 - f23.y);
  hi = make_uint2(tm_h2_bits(h01), tm_h2_bits(h23));
  lo = make_uint2(tm_h2_bits(l01), tm_h2_bits(l23));
}
// data_lib.floats_to_pixels(training=False): saturate_cast_u8(round_half_even((x + 0.5) * 255)); the float -> u8
// conversion rounds to nearest even and saturates (NaN -> 0), i.e. the same function as float_to_pixel()
__device__ __forceinline__ uint32_t tm_pixel(float x) {
  const float v = __fmul_rn(__fadd_rn(x, 0.5f), 255.f);   // add first, then multiply; no FMA contraction
  uint32_t r;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// FAST: the only destination is the uint8 image and W is even (every thread's two bytes are one aligned 16-bit store)
//
// Persistent CTAs (3 per SM) loop over the tiles.  The fp32 halo tile of t arrives by ONE TMA tiled load per tile
// (4-D map (c, x, y, b), box (C1, 66, 10, 1); out-of-image coordinates are zero-filled = the conv padding) into a raw
// staging buffer; the load of tile n+1 is issued as soon as tile n has been converted, so it overlaps the MMA phase.
// WLO = false drops the t_hi * w_lo cross term (SNTC_PRECISION_TC_F16X3_SYN2: 30 instead of 45 MMAs per row step).
template <int C1, bool FAST, bool WLO = true>
__global__ void __launch_bounds__(TM_THREADS, (TM_TX >= 64 ? 3 : (TM_TX >= 32 ? 5 : 10))) tail_s2_mma_kernel(const __grid_constant__ CUtensorMap mapT, const TailMmaParams Q) {
  static_assert(C1 % 4 == 0 && C1 <= 16, "one K=16 step per tap");
  constexpr int NPX = TM_TYH * TM_TXH;
  constexpr int PLANE = 2 * NPX;                                  // uint4 units per plane: [octet][row][x]
  constexpr uint32_t RAW_BYTES = (uint32_t)NPX * C1 * 4;
  extern __shared__ uint8_t tm_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tm_smem_raw) + 127) & ~(uintptr_t)127);
  const float4* raw = reinterpret_cast<const float4*>(smem);                            // [row][x][C1] fp32, written by TMA
  uint4* st = reinterpret_cast<uint4*>(smem + ((RAW_BYTES + 127u) & ~127u));            // [plane hi/lo][octet][row][x] x 8 fp16
  uint64_t* bar = reinterpret_cast<uint64_t*>(st + 2 * PLANE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = (Q.win + TM_TX - 1) / TM_TX, tiles_y = (Q.hin + TM_RW - 1) / TM_RW;
  const int ntiles = tiles_x * tiles_y * Q.B;
  auto tile_origin = [&](int t, int& b, int& ty0, int& tx0) {
    const int txi = t % tiles_x, r = t / tiles_x;
    tx0 = txi * TM_TX; ty0 = (r % tiles_y) * TM_RW; b = r / tiles_y;
  };

  // weight fragments -> registers (does not depend on the previous kernel)
  uint32_t wf[TM_NFRAG][2][2];
#pragma unroll
  for (int f = 0; f < TM_NFRAG; ++f)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const uint2 v = __ldg(Q.wfrag + (f * 2 + p) * 32 + lane);
      wf[f][p][0] = v.x; wf[f][p][1] = v.y;
    }
  if (threadIdx.x == 0) {
    tcx::prefetch_tmap(&mapT);
    tcx::mbar_init(bar, 1);
    tcx::fence_barrier_init();
  }
  if (C1 < 16) {   // channels C1..15 of every pixel are constant zeros: written once
    for (int px = threadIdx.x; px < NPX; px += TM_THREADS)
#pragma unroll
      for (int c4 = C1 / 4; c4 < 4; ++c4) {
        const int pos = ((c4 >> 1) * NPX + px) * 2 + (c4 & 1);
        reinterpret_cast<uint2*>(st)[pos] = make_uint2(0u, 0u);
        reinterpret_cast<uint2*>(st + PLANE)[pos] = make_uint2(0u, 0u);
      }
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");   // no-op unless launched as a programmatic dependent of layer 1
  int tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < ntiles) {
    int b, ty0, tx0;
    tile_origin(tile, b, ty0, tx0);
    tcx::mbar_expect_tx(bar, RAW_BYTES);
    tcx::tma_load_4d(smem, &mapT, bar, 0, tx0 - 1, ty0 - 1, b);
  }

  const int xs = warp * 16;                                       // this warp's strip inside the tile
  // ldmatrix.x4: lanes 0-7 -> rows 0-7 / k 0-7, 8-15 -> rows 8-15 / k 0-7, 16-23 -> rows 0-7 / k 8-15, 24-31 -> rows 8-15 / k 8-15
  const int lrow = lane & 15, loct = lane >> 4;
  const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(st + (loct * TM_TYH) * TM_TXH + xs + lrow);
  constexpr uint32_t PLANE_B = PLANE * 16, ROW_B = TM_TXH * 16;
  const int g = lane >> 2, q = lane & 3;
  // this thread's accumulator columns n = 2q, 2q+1 = bytes 2q, 2q+1 of the 6 bytes of output pixels (2x, 2x+1); q = 3 is padding
  const float bias0 = q == 1 ? Q.bias[2] : (q == 2 ? Q.bias[1] : Q.bias[0]);
  const float bias1 = q == 1 ? Q.bias[0] : (q == 2 ? Q.bias[2] : Q.bias[1]);
  const int rowb = Q.W * 3;
  uint32_t phase = 0;

  for (; tile < ntiles; tile += gridDim.x) {
    int b, ty0, tx0;
    tile_origin(tile, b, ty0, tx0);
    // ---- raw fp32 tile -> fp16 hi/lo planes ----
    tcx::mbar_wait(bar, phase);
    phase ^= 1u;
    constexpr int NQ = NPX * (C1 / 4);
#pragma unroll 4
    for (int i = threadIdx.x; i < NQ; i += TM_THREADS) {
      const int px = i / (C1 / 4), c4 = i - px * (C1 / 4);
      uint2 hi, lo;
      tm_split4(raw[i], hi, lo);
      const int pos = ((c4 >> 1) * NPX + px) * 2 + (c4 & 1);
      reinterpret_cast<uint2*>(st)[pos] = hi;
      reinterpret_cast<uint2*>(st + PLANE)[pos] = lo;
    }
    __syncthreads();                                              // planes complete; raw buffer free
    if (threadIdx.x == 0 && tile + (int)gridDim.x < ntiles) {
      int nb, nty0, ntx0;
      tile_origin(tile + gridDim.x, nb, nty0, ntx0);
      tcx::fence_proxy_async();                                   // generic reads of `raw` are ordered before the async-proxy overwrite
      tcx::mbar_expect_tx(bar, RAW_BYTES);
      tcx::tma_load_4d(smem, &mapT, bar, 0, ntx0 - 1, nty0 - 1, nb);
    }

    // ---- MMA phase ----
    uint32_t A[3][3][2][4];                                       // [row slot][jx + 1][hi/lo][regs]
    auto load_row = [&](uint32_t (*dst)[2][4], int yy) {
#pragma unroll
      for (int jx = 0; jx < 3; ++jx)
#pragma unroll
        for (int p = 0; p < 2; ++p) tm_ldmatrix_x4(dst[jx][p], a_base + p * PLANE_B + yy * ROW_B + jx * 16);
    };
    load_row(A[0], 0);
    load_row(A[1], 1);
    const int x_lo = tx0 + xs + g;                                // t-pixel of c0/c1; c2/c3: x_lo + 8
    uint8_t* u8p = Q.out_u8 ? Q.out_u8 + ((size_t)b * Q.H + 2 * ty0) * rowb + 6 * x_lo + 2 * q : nullptr;
    bool okx[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) okx[hh] = q < 3 && x_lo + 8 * hh < Q.win && 2 * (x_lo + 8 * hh) < Q.W;
#pragma unroll
    for (int i = 0; i < TM_RW; ++i) {
      load_row(A[(i + 2) % 3], i + 2);
      float acc[2][2][4];                                         // [parity][hi*hi | cross terms][regs]
      {
        float acx[2][4];                                          // hi * lo: its own dependency chain (6 chains per warp in flight)
#pragma unroll
        for (int f = 0; f < TM_NFRAG; ++f) {
          const int nt = f < 6 ? 0 : 1;
          const int d = nt == 0 ? f / 3 : (f - 6) / 3;           // dy + 1
          const int jx = f % 3;
          const uint32_t (*a)[4] = A[(i + d) % 3][jx];
          if (f == 0 || f == 6) {
            tm_mma16816_first(acc[nt][0], a[0], wf[f][0]);        // hi * hi
            tm_mma16816_first(acc[nt][1], a[1], wf[f][0]);        // lo * hi
            if (WLO) tm_mma16816_first(acx[nt], a[0], wf[f][1]);  // hi * lo
          } else {
            tm_mma16816(acc[nt][0], a[0], wf[f][0]);
            tm_mma16816(acc[nt][1], a[1], wf[f][0]);
            if (WLO) tm_mma16816(acx[nt], a[0], wf[f][1]);
          }
        }
        if (WLO) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][1][e] += acx[nt][e];
        }
      }
      const int ty = ty0 + i;
      if (ty >= Q.hin) continue;
      if (FAST) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          if (2 * ty + nt >= Q.H) continue;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (!okx[hh]) continue;
            const float v0 = fmaf(acc[nt][0][2 * hh] + acc[nt][1][2 * hh], Q.inv_scale, bias0);
            const float v1 = fmaf(acc[nt][0][2 * hh + 1] + acc[nt][1][2 * hh + 1], Q.inv_scale, bias1);
            *reinterpret_cast<uint16_t*>(u8p + (2 * i + nt) * rowb + 48 * hh) = (uint16_t)(tm_pixel(v0) | (tm_pixel(v1) << 8));
          }
        }
      } else {
        if (q == 3) continue;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int oy = 2 * ty + nt;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int x = x_lo + 8 * hh;
            if (x >= Q.win) continue;
            const float v0 = fmaf(acc[nt][0][2 * hh] + acc[nt][1][2 * hh], Q.inv_scale, bias0);
            const float v1 = fmaf(acc[nt][0][2 * hh + 1] + acc[nt][1][2 * hh + 1], Q.inv_scale, bias1);
            const int n0 = 2 * q;
            const int ox0 = 2 * x + n0 / 3, ox1 = 2 * x + (n0 + 1) / 3;
            if (Q.out) {
              float* o = Q.out + (((size_t)b * Q.hout + oy) * Q.wout + 2 * x) * 3 + n0;
              o[0] = v0; o[1] = v1;
            }
            if (oy < Q.H) {
              if (Q.out_u8) {
                uint8_t* o = Q.out_u8 + (((size_t)b * Q.H + oy) * Q.W + 2 * x) * 3 + n0;
                if (ox0 < Q.W) o[0] = (uint8_t)tm_pixel(v0);
                if (ox1 < Q.W) o[1] = (uint8_t)tm_pixel(v1);
              }
              if (Q.out_crop) {
                float* o = Q.out_crop + (((size_t)b * Q.H + oy) * Q.W + 2 * x) * 3 + n0;
                if (ox0 < Q.W) o[0] = v0;
                if (ox1 < Q.W) o[1] = v1;
              }
            }
          }
        }
      }
    }
    __syncthreads();                                              // every warp is done with the planes before the next conversion
  }
}

// ------------------------------------------------------------------------------------------------
// host side
struct TailMma {
  bool ok = false;
  float scale = 1.f;
  uint2* d_wfrag = nullptr;
  float bias[3] = {0.f, 0.f, 0.f};
  bool attr_set = false;
  int ctas_per_sm = 0;            // resident CTAs per SM of the variant in use (occupancy query), sizes the persistent grid
};

inline bool tail_mma_supported(const ConvLayer& c) {
  return c.s == 2 && c.k == 5 && c.p == 1 && c.cout == 3 && !c.append_ones && c.cin % 4 == 0 && (c.cin == 12 || c.cin == 16);
}

inline bool tail_mma_pack(const ConvLayer& c, const HostWeights& hw, TailMma& t, std::vector<void*>& owned, std::string* err) {
  float wmax = 0.f;
  for (auto& s : c.sources) for (float v : hw.at(s.kernel).second) wmax = std::max(wmax, std::fabs(v));
  int e = 0;
  if (wmax > 0.f) std::frexp(wmax, &e);
  t.scale = std::ldexp(1.f, 12 - e);                              // max |w| * S in [2^11, 2^12): the lo part stays normal
  std::vector<uint32_t> frag((size_t)TM_NFRAG * 2 * 32 * 2, 0u);
  auto h16 = [](float v) { return (uint32_t)__half_as_ushort(__float2half_rn(v)); };
  for (int f = 0; f < TM_NFRAG; ++f) {
    const int phy = f < 6 ? 0 : 1;
    const int dy = (phy == 0 ? f / 3 : (f - 6) / 3) - 1, jx = f % 3 - 1;
    const int ay = phy + c.p - 2 * dy;                            // o = 2n + a - p with n = ty + dy, o = 2 ty + phy
    for (int lane = 0; lane < 32; ++lane) {
      const int n = lane >> 2, qd = lane & 3;                     // B fragment: column n, rows k = 2qd, 2qd+1, 2qd+8, 2qd+9
      float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
      if (n < 6 && ay >= 0 && ay < c.k) {
        const int phx = n / 3, co = n % 3;
        const int ax = phx + c.p - 2 * jx;
        if (ax >= 0 && ax < c.k)
          for (int r = 0; r < 4; ++r) {
            const int ci = 2 * qd + (r & 1) + 8 * (r >> 1);
            if (ci >= c.cin) continue;
            const float w = conv_w(c, hw, ay, ax, co, ci) * t.scale;
            hi[r] = __half2float(__float2half_rn(w));
            lo[r] = w - hi[r];
          }
      }
      uint32_t* ph = &frag[((size_t)(f * 2 + 0) * 32 + lane) * 2];
      uint32_t* pl = &frag[((size_t)(f * 2 + 1) * 32 + lane) * 2];
      ph[0] = h16(hi[0]) | (h16(hi[1]) << 16); ph[1] = h16(hi[2]) | (h16(hi[3]) << 16);
      pl[0] = h16(lo[0]) | (h16(lo[1]) << 16); pl[1] = h16(lo[2]) | (h16(lo[3]) << 16);
    }
  }
  if (cudaMalloc((void**)&t.d_wfrag, frag.size() * 4) != cudaSuccess) { *err = "cudaMalloc (tail fragments) failed"; return false; }
  owned.push_back(t.d_wfrag);
  if (cudaMemcpy(t.d_wfrag, frag.data(), frag.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { *err = "cudaMemcpy (tail fragments) failed"; return false; }
  { std::vector<float> bv = pack_bias(c, hw); for (int i = 0; i < 3; ++i) t.bias[i] = bv[i]; }
  t.ok = true;
  return true;
}

struct TailMmaOut { float* f32 = nullptr; uint8_t* u8 = nullptr; float* crop = nullptr; int H = 0, W = 0; };

inline size_t tail_mma_smem(int C1) {
  const size_t raw = ((size_t)TM_TYH * TM_TXH * C1 * 4 + 127) / 128 * 128;
  return 128 + raw + (size_t)2 * 2 * TM_TYH * TM_TXH * 16 + 16;
}

template <int C1, bool FAST, bool WLO>
inline cudaError_t tail_mma_launch(const cudaLaunchConfig_t& cfg, const CUtensorMap& mapT, const TailMmaParams& Q, size_t smem) {
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tail_s2_mma_kernel<C1, FAST, WLO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return cudaLaunchKernelEx(&cfg, tail_s2_mma_kernel<C1, FAST, WLO>, mapT, Q);
}

inline int tail_mma_run(TcDriver& drv, const ConvLayer& c, TailMma& t, const float* in, int B, int h, int w, const TailMmaOut& o, bool pdl,
                        cudaStream_t s, uint64_t* launches, std::string* err, bool w_lo = true) {
  if (!drv.encode) { *err = "cuTensorMapEncodeTiled unavailable"; return 2; }
  TailMmaParams Q{};
  Q.x = in; Q.B = B; Q.hin = h; Q.win = w; Q.wfrag = t.d_wfrag; Q.inv_scale = 1.f / t.scale;
  for (int i = 0; i < 3; ++i) Q.bias[i] = t.bias[i];
  Q.out = o.f32; Q.hout = 2 * h; Q.wout = 2 * w; Q.out_u8 = o.u8; Q.out_crop = o.crop; Q.H = o.H; Q.W = o.W;
  // t [B,h,w,C1] fp32 as a 4-D tensor (c, x, y, b); box = one halo tile
  CUtensorMap mapT;
  {
    const int C1 = c.cin;
    cuuint64_t dims[4] = {(cuuint64_t)C1, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C1 * 4, (cuuint64_t)w * C1 * 4, (cuuint64_t)h * w * C1 * 4};
    cuuint32_t box[4] = {(cuuint32_t)C1, (cuuint32_t)TM_TXH, (cuuint32_t)TM_TYH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = drv.encode(&mapT, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled(tail) failed: " + std::to_string((int)r); return 2; }
  }
  const size_t smem = tail_mma_smem(c.cin);
  if (!t.attr_set) {
    t.attr_set = true;
    int n = 0;
    cudaError_t e = cudaFuncSetAttribute(tail_s2_mma_kernel<12, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_mma_smem(12));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_s2_mma_kernel<16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_mma_smem(16));
    if (e == cudaSuccess)
      e = c.cin == 12 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tail_s2_mma_kernel<12, true, true>, TM_THREADS, smem)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tail_s2_mma_kernel<16, true, true>, TM_THREADS, smem);
    t.ctas_per_sm = (e == cudaSuccess && n > 0) ? n : 1;
    cudaGetLastError();
  }
  const int ntiles = ((w + TM_TX - 1) / TM_TX) * ((h + TM_RW - 1) / TM_RW) * B;
  if (ntiles <= 0) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(ntiles, t.ctas_per_sm * drv.num_sms));
  cfg.blockDim = dim3(TM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  const bool fast = o.u8 && !o.f32 && !o.crop && (o.W % 2) == 0;
  cudaError_t e;
  const int variant = (c.cin == 12 ? 0 : 4) + (fast ? 2 : 0) + (w_lo ? 1 : 0);
  switch (variant) {
    case 0: e = tail_mma_launch<12, false, false>(cfg, mapT, Q, smem); break;
    case 1: e = tail_mma_launch<12, false, true>(cfg, mapT, Q, smem); break;
    case 2: e = tail_mma_launch<12, true, false>(cfg, mapT, Q, smem); break;
    case 3: e = tail_mma_launch<12, true, true>(cfg, mapT, Q, smem); break;
    case 4: e = tail_mma_launch<16, false, false>(cfg, mapT, Q, smem); break;
    case 5: e = tail_mma_launch<16, false, true>(cfg, mapT, Q, smem); break;
    case 6: e = tail_mma_launch<16, true, false>(cfg, mapT, Q, smem); break;
    default: e = tail_mma_launch<16, true, true>(cfg, mapT, Q, smem); break;
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("tail_s2_mma_kernel launch: ") + cudaGetErrorString(e); return 2; }
  if (launches) (*launches)++;
  return 0;
}

}  // namespace sntc
