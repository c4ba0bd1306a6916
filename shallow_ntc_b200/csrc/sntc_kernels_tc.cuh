// tcgen05 / TMEM / TMA band-GEMM path (SNTC_PRECISION_TC_F16X3), sm_100a only.
//
// One kernel computes one band of a transposed convolution (see sntc_plan.hpp) as an implicit GEMM
//     D[128 cells, BN] = sum over taps (jy,jx) and 64-channel blocks  A_tap[128, 64] * W_tap[64, BN]
// on the 5th-gen tensor cores:
//   * A tiles: TMA 4-D tiled loads (c, x, y, b) of a TH x TW patch of the NHWC fp16 activation planes,
//     shifted by the tap (negative / out-of-range coordinates are zero-filled by TMA = conv padding);
//   * W tiles: TMA 2-D loads of the packed K-major band matrix;
//   * both land in 128B-swizzled shared memory and feed tcgen05.mma (kind::f16, M=128, N=BN, K=16)
//     issued by one thread; the fp32 accumulator lives in TMEM;
//   * fp32-class accuracy from fp16 tensor cores: every operand is split x = hi + lo (two fp16
//     planes, weights pre-scaled by a power of two) and the product is accumulated as
//     hi*hi + lo*hi + hi*lo (3 MMA passes per k-block; the dropped lo*lo term is < 2^-22 relative);
//   * epilogue warps read TMEM (tcgen05.ld), apply scale/bias/activation and write either the next
//     layer's fp16 hi/lo planes, an fp32 tensor, the final pixels, or -- for the last hyper-synthesis
//     layer -- directly y_hat = q + mu and the scale-table rows idx (mu/sigma never reach HBM).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <string>
#include <vector>
#include "sntc_plan.hpp"

namespace sntc {

enum { TC_OK = 0, TC_NOT_HANDLED = 1, TC_ERROR = 2 };
enum { TC_EPI_PLAIN = 0, TC_EPI_HYPER_FINAL = 1 };

constexpr int TC_BM = 128;      // cells per tile (UMMA M)
constexpr int TC_BK = 64;       // fp16 channels per k-block (128 bytes = one swizzle row)
constexpr int TC_THREADS = 256;

struct TcParams {
  int B, hin, win;
  int s, p, phy0, nphx, phx0, Ty, Tx, mloy, mlox;
  int TH, TW, tiles_y, tiles_x;
  int kblocks;        // 64-channel blocks per tap
  int last_kmma;      // MMAs (K=16 each) in the last block of a tap (1..4)
  int N, cout, BN, stages, tmem_cols;
  float inv_scale;
  const float* bias;
  int act;
  int hout, wout;
  int epi;
  // TC_EPI_PLAIN destinations (any subset)
  __half* out_hi; __half* out_lo;          // [B,hout,wout,cout] fp16 planes (next layer's A operand)
  float* out_f32;                          // [B,hout,wout,cout]
  uint8_t* out_u8; float* out_crop; int H, W;   // cropped pixels [B,H,W,cout]
  // TC_EPI_HYPER_FINAL: columns [0,Cy) = mu, [Cy,2Cy) = raw sigma
  const void* q; int q_kind; int Cy; float max_index; int trunc;
  float* y_hat; uint8_t* idx;              // + out_hi/out_lo = planes of y_hat [B,hout,wout,Cy]
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
namespace tcx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
    : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> CUDA error on the host), never hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s
  }
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
    ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
    ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], kind::f16, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
    ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
//   bits [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64)
//   | [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace tcx

// x = hi + lo with hi = fp16(x) (saturated to the finite range), lo = fp16(x - hi)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  float c = fminf(fmaxf(x, -65504.f), 65504.f);
  hi = __float2half_rn(c);
  lo = __float2half_rn(c - __half2float(hi));
}

__device__ __forceinline__ void store8_planes(__half* hi, __half* lo, size_t off, const float* v) {
  __align__(16) __half h[8];
  __align__(16) __half l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_f16(v[i], h[i], l[i]);
  *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<const uint4*>(l);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
band_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                    const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = P.BN;
  const uint32_t a_bytes = TC_BM * 128;            // one A plane tile
  const uint32_t b_bytes = (uint32_t)BN * 128;     // one W plane tile
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + P.stages;
  uint64_t* tmem_full_bar = empty_bar + P.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int tx = tile % P.tiles_x, ty = (tile / P.tiles_x) % P.tiles_y, b = tile / (P.tiles_x * P.tiles_y);
  const int n0 = blockIdx.y * BN;
  const int iy0 = ty * P.TH, ix0 = tx * P.TW;      // first cell (index within the band's h x w cell grid)
  const int nk = P.Ty * P.Tx * P.kblocks;

  if (warp == 0 && lane == 0) {
    tcx::prefetch_tmap(&mapAhi); tcx::prefetch_tmap(&mapAlo); tcx::prefetch_tmap(&mapBhi); tcx::prefetch_tmap(&mapBlo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < P.stages; ++i) { tcx::mbar_init(&full_bar[i], 1); tcx::mbar_init(&empty_bar[i], 1); }
    tcx::mbar_init(tmem_full_bar, 1);
    tcx::fence_barrier_init();
  }
  if (warp == 2) tcx::tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tcx::tc_fence_before();
  __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int it = 0; it < nk; ++it) {
        const int st = it % P.stages;
        const uint32_t ph = (uint32_t)(it / P.stages) & 1u;
        tcx::mbar_wait(&empty_bar[st], ph ^ 1u);
        const int tap = it / P.kblocks, kb = it - tap * P.kblocks;
        const int jy = tap / P.Tx, jx = tap - jy * P.Tx;
        uint8_t* sa = smem + (size_t)st * stage_bytes;
        tcx::mbar_expect_tx(&full_bar[st], stage_bytes);
        const int cx = P.mlox + ix0 - jx, cy = P.mloy + iy0 - jy, cc = kb * TC_BK;
        tcx::tma_load_4d(sa, &mapAhi, &full_bar[st], cc, cx, cy, b);
        tcx::tma_load_4d(sa + a_bytes, &mapAlo, &full_bar[st], cc, cx, cy, b);
        const int kcol = (tap * P.kblocks + kb) * TC_BK;
        tcx::tma_load_2d(sa + 2 * a_bytes, &mapBhi, &full_bar[st], kcol, n0);
        tcx::tma_load_2d(sa + 2 * a_bytes + b_bytes, &mapBlo, &full_bar[st], kcol, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = tcx::make_idesc(TC_BM, BN);
      uint32_t acc = 0;
      for (int it = 0; it < nk; ++it) {
        const int st = it % P.stages;
        const uint32_t ph = (uint32_t)(it / P.stages) & 1u;
        tcx::mbar_wait(&full_bar[st], ph);
        tcx::tc_fence_after();
        const uint32_t sa = tcx::smem_u32(smem + (size_t)st * stage_bytes);
        const uint64_t a_hi = tcx::make_smem_desc(sa), a_lo = tcx::make_smem_desc(sa + a_bytes);
        const uint64_t b_hi = tcx::make_smem_desc(sa + 2 * a_bytes), b_lo = tcx::make_smem_desc(sa + 2 * a_bytes + b_bytes);
        const int kb = it % P.kblocks;
        const int nm = (kb == P.kblocks - 1) ? P.last_kmma : 4;
        // lo*hi and hi*lo first (small terms), hi*hi last
        for (int k = 0; k < nm; ++k) { tcx::umma_f16(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, acc); acc = 1; }
        for (int k = 0; k < nm; ++k) tcx::umma_f16(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
        for (int k = 0; k < nm; ++k) tcx::umma_f16(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, 1);
        tcx::umma_commit(&empty_bar[st]);       // frees this smem stage when the MMAs above retire
      }
      tcx::umma_commit(tmem_full_bar);          // accumulator complete
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int ew = warp - 4;                    // TMEM lanes [32*ew, 32*ew+32)
    const int r = ew * 32 + lane;               // row of the tile = cell
    const int iy = iy0 + r / P.TW, ix = ix0 + r % P.TW;
    const bool cell_ok = iy < P.hin && ix < P.win;
    const int my = P.mloy + iy, mx = P.mlox + ix;
    if (nk > 0) {
      tcx::mbar_wait(tmem_full_bar, 0);
      tcx::tc_fence_after();
    }
    const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16);
    const bool vec = (P.cout % 8) == 0;
    for (int c = 0; c < BN; c += 8) {
      float v[8];
      if (nk > 0) tcx::tmem_ld8(trow + (uint32_t)c, v);      // warp-wide: executed by all lanes
      else { for (int i = 0; i < 8; ++i) v[i] = 0.f; }
      const int n = n0 + c;
      if (!cell_ok || n >= P.N) continue;
      if (vec) {
        const int co = n % P.cout, ph = n / P.cout;
        const int oy = P.s * my + P.phy0 + ph / P.nphx - P.p, ox = P.s * mx + P.phx0 + ph % P.nphx - P.p;
        if (oy < 0 || oy >= P.hout || ox < 0 || ox >= P.wout) continue;
        const size_t pix = ((size_t)b * P.hout + oy) * P.wout + ox;
        float4 b0 = __ldg(reinterpret_cast<const float4*>(P.bias + co));
        float4 b1 = __ldg(reinterpret_cast<const float4*>(P.bias + co + 4));
        float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], P.inv_scale, bb[i]);
        if (P.epi == TC_EPI_HYPER_FINAL) {
          if (co < P.Cy) {   // mu half: y_hat = q + mu                     mshyper/models.py:278
            const size_t e = pix * P.Cy + co;
            float4 q0 = load_q4(P.q, P.q_kind, e), q1 = load_q4(P.q, P.q_kind, e + 4);
            float y[8] = {__fadd_rn(q0.x, v[0]), __fadd_rn(q0.y, v[1]), __fadd_rn(q0.z, v[2]), __fadd_rn(q0.w, v[3]),
                          __fadd_rn(q1.x, v[4]), __fadd_rn(q1.y, v[5]), __fadd_rn(q1.z, v[6]), __fadd_rn(q1.w, v[7])};
            if (P.y_hat) {
              *reinterpret_cast<float4*>(P.y_hat + e) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(P.y_hat + e + 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
            if (P.out_hi) store8_planes(P.out_hi, P.out_lo, e, y);
            if (P.out_f32) {
              float* o = P.out_f32 + pix * P.cout + co;
              *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
          } else {           // sigma half: idx = round(clamp(exp(sigma), 0, S-1))   :274-276
            if (P.idx) {
              __align__(8) uint8_t o8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o8[i] = scale_index(v[i], P.max_index, P.trunc);
              *reinterpret_cast<uint2*>(P.idx + pix * P.Cy + (co - P.Cy)) = *reinterpret_cast<const uint2*>(o8);
            }
            if (P.out_f32) {
              float* o = P.out_f32 + pix * P.cout + co;
              *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = apply_act(v[i], P.act);
          if (P.out_hi) store8_planes(P.out_hi, P.out_lo, pix * P.cout + co, v);
          if (P.out_f32) {
            float* o = P.out_f32 + pix * P.cout + co;
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      } else {
        // generic scalar path (final layers with cout = 3, ...)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int nn = n + i;
          if (nn >= P.N) break;
          const int co = nn % P.cout, ph = nn / P.cout;
          const int oy = P.s * my + P.phy0 + ph / P.nphx - P.p, ox = P.s * mx + P.phx0 + ph % P.nphx - P.p;
          if (oy < 0 || oy >= P.hout || ox < 0 || ox >= P.wout) continue;
          float x = apply_act(fmaf(v[i], P.inv_scale, __ldg(P.bias + co)), P.act);
          const size_t pix = ((size_t)b * P.hout + oy) * P.wout + ox;
          if (P.out_f32) P.out_f32[pix * P.cout + co] = x;
          if ((P.out_u8 || P.out_crop) && oy < P.H && ox < P.W) {
            const size_t qi = (((size_t)b * P.H + oy) * P.W + ox) * P.cout + co;
            if (P.out_u8) P.out_u8[qi] = float_to_pixel(x);
            if (P.out_crop) P.out_crop[qi] = x;
          }
        }
      }
    }
  }
  tcx::tc_fence_before();
  __syncthreads();
  if (warp == 2) tcx::tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// f32 NHWC -> fp16 hi/lo planes (input of the first tensor-core layer)
__global__ void split_planes_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, size_t n8) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  store8_planes(hi, lo, i * 8, v);
}

// ------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcDriver {
  PFN_encodeTiled encode = nullptr;
  std::string err;
  void init() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
      err = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
      cudaGetLastError();
      return;
    }
    encode = (PFN_encodeTiled)fn;
  }
};

struct TcBand {
  int N = 0, BN = 0, ntiles = 0, stages = 0, tmem_cols = 0;
  size_t row0 = 0;          // first row of this band in the packed matrices
  CUtensorMap mapBhi, mapBlo;
};

struct TcConv {
  bool ok = false;
  int cin = 0, kblocks = 0, last_kmma = 4, ktot_pad = 0;   // per-tap K padded to kblocks*64
  float scale = 1.f;
  __half* d_hi = nullptr; __half* d_lo = nullptr;          // [rows][Kmax] K-major
  size_t kmax = 0;                                         // row pitch (elements)
  std::vector<TcBand> bands;
};

struct TcDevBuf {
  void* p = nullptr; size_t cap = 0;
  bool ensure(size_t bytes) {
    if (bytes <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 1024;
    if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return false; }
    cap = want;
    return true;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct TcModelState {
  std::vector<TcConv> hyper, syn;     // parallel to Transform::convs
  TcDevBuf plane[4];                  // hi/lo ping-pong activation planes
  TcDevBuf yh[2];                     // hi/lo planes of y_hat (output of the fused hyper-synthesis head)
  bool smem_attr_set = false;
  void release() { for (auto& b : plane) b.release(); yh[0].release(); yh[1].release(); }
};

inline int tc_choose_bn(int N, int* stages) {
  // widest tile that keeps >= 3 pipeline stages in 227 KB and wastes the least padded columns
  int best = 0; long best_cost = -1;
  for (int bn = 160; bn >= 16; bn -= 16) {
    int tiles = (N + bn - 1) / bn;
    long cost = (long)tiles * bn * 1000 + tiles * 40 * 64;   // padded MMA work + per-tile A re-read/epilogue overhead
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  int stage_bytes = 2 * TC_BM * 128 + 2 * best * 128;
  int st = (227 * 1024 - 2048) / stage_bytes;
  *stages = st > 6 ? 6 : st;
  return best;
}

inline bool tc_make_map_2d(TcDriver& drv, CUtensorMap* map, void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems,
                           uint32_t box_cols, uint32_t box_rows, std::string* err) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = drv.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled(2d) failed: " + std::to_string((int)r); return false; }
  return true;
}

inline bool tc_make_map_4d(TcDriver& drv, CUtensorMap* map, void* base, int C, int w, int h, int B, int TW, int TH, std::string* err) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)w * C * 2, (cuuint64_t)h * w * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = drv.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled(4d) failed: " + std::to_string((int)r); return false; }
  return true;
}

// A conv layer runs on the tensor cores when its input channel count is TMA-addressable.
inline bool tc_conv_supported(const ConvLayer& c) {
  return !c.append_ones && c.cin % 8 == 0 && c.cin >= 64;
}

inline bool tc_pack_conv(TcDriver& drv, const ConvLayer& c, const HostWeights& hw, TcConv& t, std::vector<void*>& owned, std::string* err) {
  t.cin = c.cin;
  t.kblocks = (c.cin + TC_BK - 1) / TC_BK;
  int rem = c.cin - (t.kblocks - 1) * TC_BK;
  t.last_kmma = (rem + 15) / 16;
  t.ktot_pad = t.kblocks * TC_BK;
  // power-of-two scale so that max |w| * S lies in [2^11, 2^12)
  float wmax = 0.f;
  for (auto& s : c.sources) for (float v : hw.at(s.kernel).second) wmax = std::max(wmax, std::fabs(v));
  int e = 0;
  if (wmax > 0.f) { std::frexp(wmax, &e); }
  t.scale = std::ldexp(1.f, 12 - e);
  size_t rows = 0, kmax = 0;
  for (auto& b : c.bands) { rows += (size_t)b.N; kmax = std::max(kmax, (size_t)b.Ty * b.Tx * t.ktot_pad); }
  if (kmax == 0) kmax = TC_BK;
  t.kmax = kmax;
  std::vector<__half> hi(rows * kmax, __float2half(0.f)), lo(rows * kmax, __float2half(0.f));
  t.bands.resize(c.bands.size());
  size_t row0 = 0;
  for (size_t bi = 0; bi < c.bands.size(); ++bi) {
    const Band& b = c.bands[bi];
    TcBand& tb = t.bands[bi];
    tb.N = b.N; tb.row0 = row0;
    tb.BN = tc_choose_bn(b.N, &tb.stages);
    tb.ntiles = (b.N + tb.BN - 1) / tb.BN;
    tb.tmem_cols = tb.BN <= 32 ? 32 : (tb.BN <= 64 ? 64 : (tb.BN <= 128 ? 128 : 256));
    for (int fy = 0; fy < b.nphy; ++fy) for (int fx = 0; fx < b.nphx; ++fx) for (int co = 0; co < c.cout; ++co) {
      size_t row = row0 + (size_t)(fy * b.nphx + fx) * c.cout + co;
      for (int jy = 0; jy < b.Ty; ++jy) for (int jx = 0; jx < b.Tx; ++jx) {
        int ay = b.phy0 + fy + c.s * jy, ax = b.phx0 + fx + c.s * jx;
        if (ay >= c.k || ax >= c.k) continue;
        size_t col0 = (size_t)(jy * b.Tx + jx) * t.ktot_pad;
        for (int ci = 0; ci < c.cin; ++ci) {
          float w = conv_w(c, hw, ay, ax, co, ci) * t.scale;
          __half h = __float2half_rn(w);
          hi[row * kmax + col0 + ci] = h;
          lo[row * kmax + col0 + ci] = __float2half_rn(w - __half2float(h));
        }
      }
    }
    row0 += (size_t)b.N;
  }
  size_t bytes = hi.size() * sizeof(__half);
  if (cudaMalloc((void**)&t.d_hi, bytes) != cudaSuccess || cudaMalloc((void**)&t.d_lo, bytes) != cudaSuccess) { *err = "cudaMalloc (tc weights) failed"; return false; }
  owned.push_back(t.d_hi); owned.push_back(t.d_lo);
  if (cudaMemcpy(t.d_hi, hi.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(t.d_lo, lo.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    *err = "cudaMemcpy (tc weights) failed"; return false;
  }
  for (size_t bi = 0; bi < c.bands.size(); ++bi) {
    const Band& b = c.bands[bi];
    TcBand& tb = t.bands[bi];
    if (b.N == 0) continue;
    uint64_t kcols = (uint64_t)std::max(1, b.Ty * b.Tx) * t.ktot_pad;
    if (!tc_make_map_2d(drv, &tb.mapBhi, t.d_hi + tb.row0 * kmax, kcols, (uint64_t)b.N, kmax, TC_BK, (uint32_t)tb.BN, err)) return false;
    if (!tc_make_map_2d(drv, &tb.mapBlo, t.d_lo + tb.row0 * kmax, kcols, (uint64_t)b.N, kmax, TC_BK, (uint32_t)tb.BN, err)) return false;
  }
  t.ok = true;
  return true;
}

inline bool tc_finalize(TcDriver& drv, TcModelState& st, Transform* hyper, Transform* syn, const HostWeights& hw,
                        std::vector<void*>& owned, std::string* err) {
  if (!drv.encode) { *err = drv.err.empty() ? "cuTensorMapEncodeTiled unavailable" : drv.err; return false; }
  auto pack = [&](Transform* t, std::vector<TcConv>& out) {
    if (!t) return true;
    out.resize(t->convs.size());
    for (size_t i = 0; i < t->convs.size(); ++i)
      if (tc_conv_supported(t->convs[i]) && !tc_pack_conv(drv, t->convs[i], hw, out[i], owned, err)) return false;
    return true;
  };
  if (!pack(hyper, st.hyper) || !pack(syn, st.syn)) return false;
  if (!st.smem_attr_set) {
    cudaError_t e = cudaFuncSetAttribute(band_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return false; }
    st.smem_attr_set = true;
  }
  return true;
}

struct TcConvOut {
  __half* hi = nullptr; __half* lo = nullptr; float* f32 = nullptr;
  uint8_t* u8 = nullptr; float* crop = nullptr; int H = 0, W = 0;
  // hyper-final fusion
  bool hyper_final = false; const void* q = nullptr; int q_kind = 0; int Cy = 0; float max_index = 63.f; bool trunc = false;
  float* y_hat = nullptr; uint8_t* idx = nullptr;
};

inline void tc_choose_patch(int h, int w, int* TH, int* TW) {
  static const int cand[8][2] = {{8, 16}, {16, 8}, {4, 32}, {32, 4}, {2, 64}, {64, 2}, {1, 128}, {128, 1}};
  long best = -1;
  for (auto& c : cand) {
    long cost = (long)((h + c[0] - 1) / c[0]) * ((w + c[1] - 1) / c[1]);
    if (best < 0 || cost < best) { best = cost; *TH = c[0]; *TW = c[1]; }
  }
}

// Launches all bands of one conv layer.  Input: fp16 planes [B,h,w,cin].
inline int tc_run_conv(TcDriver& drv, const ConvLayer& c, const TcConv& t, const __half* in_hi, const __half* in_lo, int B, int h, int w,
                       const TcConvOut& o, cudaStream_t s, uint64_t* launches, std::string* err) {
  int TH, TW;
  tc_choose_patch(h, w, &TH, &TW);
  CUtensorMap mapAhi, mapAlo;
  if (!tc_make_map_4d(drv, &mapAhi, (void*)in_hi, c.cin, w, h, B, TW, TH, err)) return TC_ERROR;
  if (!tc_make_map_4d(drv, &mapAlo, (void*)in_lo, c.cin, w, h, B, TW, TH, err)) return TC_ERROR;
  for (size_t yi = 0, bi = 0; yi < c.by.size(); ++yi)
    for (size_t xi = 0; xi < c.bx.size(); ++xi, ++bi) {
      const Band& b = c.bands[bi];
      const TcBand& tb = t.bands[bi];
      if (b.N == 0) continue;
      TcParams P{};
      P.B = B; P.hin = h; P.win = w; P.s = c.s; P.p = c.p;
      P.phy0 = b.phy0; P.nphx = b.nphx; P.phx0 = b.phx0; P.Ty = b.Ty; P.Tx = b.Tx;
      P.mloy = c.by[yi].mlo; P.mlox = c.bx[xi].mlo;
      P.TH = TH; P.TW = TW; P.tiles_y = (h + TH - 1) / TH; P.tiles_x = (w + TW - 1) / TW;
      P.kblocks = t.kblocks; P.last_kmma = t.last_kmma;
      P.N = b.N; P.cout = c.cout; P.BN = tb.BN; P.stages = tb.stages; P.tmem_cols = tb.tmem_cols;
      P.inv_scale = 1.f / t.scale; P.bias = c.d_bias; P.act = c.act;
      P.hout = h * c.s; P.wout = w * c.s;
      P.epi = o.hyper_final ? TC_EPI_HYPER_FINAL : TC_EPI_PLAIN;
      P.out_hi = o.hi; P.out_lo = o.lo; P.out_f32 = o.f32; P.out_u8 = o.u8; P.out_crop = o.crop; P.H = o.H; P.W = o.W;
      P.q = o.q; P.q_kind = o.q_kind; P.Cy = o.Cy; P.max_index = o.max_index; P.trunc = o.trunc ? 1 : 0; P.y_hat = o.y_hat; P.idx = o.idx;
      size_t smem = (size_t)tb.stages * (2 * TC_BM * 128 + 2 * tb.BN * 128) + 1024 + 256;
      dim3 grid((unsigned)(P.tiles_x * P.tiles_y * B), (unsigned)tb.ntiles);
      band_gemm_tc_kernel<<<grid, TC_THREADS, smem, s>>>(mapAhi, mapAlo, tb.mapBhi, tb.mapBlo, P);
      if (launches) (*launches)++;
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { *err = std::string("band_gemm_tc_kernel launch: ") + cudaGetErrorString(e); return TC_ERROR; }
    }
  return TC_OK;
}

}  // namespace sntc
