// Window-GEMM tail of the two-layer synthesis on tcgen05: ConvT(5, 2, 12 -> 3, p = 1) + crop + uint8 epilogue
// (common/transforms.py:311-313, 350-353 + image_utils.py:22-23, 69-71 + data_lib.py:48-52), sm_100a only.
//
// The op has 900 useful MACs per t-pixel; what it costs on the tensor cores is the delivery of the A operand
// (a 128-row UMMA reads 4 KB of shared memory per K = 16 step whatever N is).  With one GEMM row per PIXEL and one K step
// per tap (sntc_kernels_tail_tc.cuh) a quarter of every K step is channel padding (12 -> 16) and N is 16; the warp-MMA
// kernel (sntc_kernels_tail_mma.cuh) is bound by the mma.sync issue rate (2 clk per m16n8k16 and SM, 45 per 16 pixels).
// Here one GEMM row is a WINDOW: TZ_J = 4 consecutive t-pixels of one row that produce 8 x 2 output pixels,
//     D[window, (phy, j, phx, co)] = sum_{dy} sum_{e < 72}  T[row + dy][12 * (4 xb - 1) + e] * Wz[dy][e][(phy, j, phx, co)]
// i.e. per input row dy in {-1, 0, +1} the K dimension is the CONTIGUOUS run of 6 pixels x 12 channels (72 fp16, padded to
// 80 = 5 K steps by 8 zero-weighted values of the next pixel) and the weights are the small banded (Toeplitz) matrix
// Wz[dy][(jj, ci)][(phy, j, phx, co)] = W[phy + 1 - 2 dy][phx + 1 - 2 (jj - 1 - j)][co][ci].  15 K steps per 512 t-pixels instead of
// 36; the zero blocks of Wz cost tensor FLOPs, which are free here.  Split-fp16 as everywhere (hi*hi + lo*hi + hi*lo):
// B holds [Wz_hi | Wz_lo] side by side, so t_hi * [Wz_hi | Wz_lo] is ONE instruction (N = 96) and t_lo * Wz_hi (N = 48) the
// second; the epilogue adds the two column groups.
//
// A windows in shared memory: un-swizzled K-major core matrices (8 rows x 16 bytes) want 8 consecutive windows' 16-byte
// K chunks adjacent, so the planes are laid out [chunk c][tile row][window xb] x 16 B and a tap dy is the same plane read
// through a descriptor whose start is shifted by one tile row.  Windows overlap (stride 4 pixels = 6 chunks, length 10
// chunks), so the conversion warps write chunks 0-3 of a window a second time as chunks 6-9 of its left neighbour.
//
// Pipeline per persistent CTA (one per SM):  warp 0: TMA of the raw fp32 halo tile (2 stages)  ->  warps 8-15: fp32 ->
// fp16 hi / lo split into the window layout (2 stages)  ->  warp 1: 30 UMMAs per tile  ->  TMEM (2 buffers)  ->
// warps 4-7: scale, bias, sat_u8(rint((x + .5) * 255)), 24 contiguous bytes per thread and image row.
#pragma once
#include "sntc_kernels_tc.cuh"

namespace sntc {

constexpr int TZ_C1 = 12;                          // hidden channels
constexpr int TZ_J = 4;                            // t-pixels per window
constexpr int TZ_ROWS = 8, TZ_XB = 16;             // windows per tile: 8 rows x 16 (= UMMA M = 128)
constexpr int TZ_TX = TZ_J * TZ_XB;                // 64 t-pixels per tile row
constexpr int TZ_RH = TZ_ROWS + 2, TZ_TXH = TZ_TX + 2;
constexpr int TZ_KSTEPS = 5;                       // K = 16 steps per input row: 6 pixels x 12 channels = 72 -> 80
constexpr int TZ_NCH = 2 * TZ_KSTEPS;              // 16-byte chunks per window
constexpr int TZ_CHS = TZ_J * TZ_C1 / 8;           // chunk stride between neighbouring windows (6)
constexpr int TZ_ROWCH = TZ_TXH * TZ_C1 / 8;       // chunks of one raw tile row (99)
constexpr int TZ_NOUT = 48;                        // (phy, j, phx, co)
constexpr uint32_t TZ_CS = TZ_RH * TZ_XB * 16;                      // bytes between chunk planes
constexpr uint32_t TZ_PLANE = (TZ_NCH * TZ_CS + 127u) & ~127u;      // one fp16 plane (hi or lo) of a tile
constexpr uint32_t TZ_RAW = ((uint32_t)TZ_RH * TZ_TXH * TZ_C1 * 4 + 127u) & ~127u;
constexpr uint32_t TZ_WSTEP = 2 * 96 * 16;                          // weights of one K step: [2 chunks][96 columns][8 fp16]
constexpr uint32_t TZ_WBYTES = 3 * TZ_KSTEPS * TZ_WSTEP;
constexpr int TZ_CONV_WARPS = 8;                  // warps 8..15
constexpr int TZ_THREADS = 32 * (8 + TZ_CONV_WARPS);
constexpr uint32_t TZ_ACC_STRIDE = 128;            // TMEM columns per accumulator buffer (96 used)
constexpr size_t TZ_SMEM = 1024 + TZ_WBYTES + 4 * TZ_PLANE + 2 * TZ_RAW + 256;

struct TailTzParams {
  int B, hin, win;               // t [B,hin,win,12] fp32
  int tiles_x, tiles_y, ntiles;
  int reverse;                   // walk the tiles from the last one: what layer 1 wrote last is still in L2
  float inv_scale; float bias[3];
  float* out; int hout, wout;    // optional f32 [B,hout,wout,3]
  uint8_t* out_u8; float* out_crop; int H, W;
  const uint4* w;                // [15 K steps][2 chunks][96 columns][8 fp16]: columns < 48 = hi, >= 48 = lo
  int fast;                      // uint8 is the only destination, W % 8 == 0, 8-byte aligned base: 8-byte stores
  long long* trace;              // debug timeline (SNTC_TC_TRACE=1): block 0, [64 tiles][8] clock64 stamps
};

__device__ __forceinline__ uint32_t tz_pixel(float x) {   // data_lib.floats_to_pixels(training=False), as float_to_pixel()
  const float v = __fmul_rn(__fadd_rn(x, 0.5f), 255.f);
  uint32_t r;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void tz_split8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float c0 = fminf(fmaxf(v[2 * i], -65504.f), 65504.f), c1 = fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f);
    const __half2 hh = __floats2half2_rn(c0, c1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(c0 - hf.x, c1 - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// WLO = false drops the t_hi * w_lo cross term (SNTC_PRECISION_TC_F16X3_SYN2)
template <bool WLO>
__global__ void __launch_bounds__(TZ_THREADS, 1) tail_tz_kernel(const __grid_constant__ CUtensorMap mapT, const TailTzParams P) {
  extern __shared__ uint8_t tz_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tz_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;                                   // weights
  uint8_t* planes = wsm + TZ_WBYTES;                     // [stage][hi | lo]
  uint8_t* raw = planes + 4 * TZ_PLANE;                  // [stage] fp32 halo tile [row][x][12], written by TMA
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + 2 * TZ_RAW);
  uint64_t *raw_full = bars, *raw_empty = bars + 2, *pl_full = bars + 4, *pl_empty = bars + 6, *acc_full = bars + 8, *acc_empty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) tcx::prefetch_tmap(&mapT);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      tcx::mbar_init(&raw_full[i], 1); tcx::mbar_init(&raw_empty[i], TZ_CONV_WARPS);
      tcx::mbar_init(&pl_full[i], TZ_CONV_WARPS); tcx::mbar_init(&pl_empty[i], 1);
      tcx::mbar_init(&acc_full[i], 1); tcx::mbar_init(&acc_empty[i], 4);
    }
    tcx::fence_barrier_init();
  }
  if (warp == 2) tcx::tmem_alloc(tmem_slot, 2 * TZ_ACC_STRIDE);
  // weights -> shared memory; the planes start as zeros (chunk 9 of the last window of a row lies past the raw tile and is never written)
  for (uint32_t i = threadIdx.x; i < TZ_WBYTES / 16; i += TZ_THREADS) reinterpret_cast<uint4*>(wsm)[i] = __ldg(P.w + i);
  for (uint32_t i = threadIdx.x; i < 4 * TZ_PLANE / 16; i += TZ_THREADS) reinterpret_cast<uint4*>(planes)[i] = make_uint4(0u, 0u, 0u, 0u);
  tcx::fence_proxy_async();
  tcx::tc_fence_before();
  __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  tcx::pdl_wait();                                       // t is the previous kernel's output

  auto tile_of = [&](int k) { return P.reverse ? P.ntiles - 1 - k : k; };
  auto tile_origin = [&](int t, int& b, int& ty0, int& tx0) {
    const int txi = t % P.tiles_x, r = t / P.tiles_x;
    tx0 = txi * TZ_TX; ty0 = (r % P.tiles_y) * TZ_ROWS; b = r / P.tiles_y;
  };

  if (warp == 0) {
    // ===== TMA producer: the fp32 halo tile (rows ty0-1 .. ty0+8, pixels tx0-1 .. tx0+64); out-of-image = zeros = conv padding =====
    uint32_t it = 0;
    for (int k = blockIdx.x; k < P.ntiles; k += gridDim.x, ++it) {
      const uint32_t st = it & 1u, ph = (it >> 1) & 1u;
      int b, ty0, tx0;
      tile_origin(tile_of(k), b, ty0, tx0);
      tcx::mbar_wait(&raw_empty[st], ph ^ 1u);
      if (P.trace && blockIdx.x == 0 && lane == 0 && it < 64) P.trace[it * 8 + 0] = clock64();
      if (tcx::elect_one()) {
        tcx::mbar_expect_tx(&raw_full[st], (uint32_t)TZ_RH * TZ_TXH * TZ_C1 * 4);
        tcx::tma_load_4d(raw + st * TZ_RAW, &mapT, &raw_full[st], 0, tx0 - 1, ty0 - 1, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer: 15 K steps x {t_hi * [Wz_hi | Wz_lo], t_lo * Wz_hi} =====
    const uint32_t hi_word = (1u << 14);                               // descriptor version 1, no swizzle
    const uint32_t a_lo_word = ((TZ_CS >> 4) << 16);                   // LBO: next K chunk
    const uint32_t a_hi_word = hi_word | (128u >> 4);                  // SBO: next 8 windows
    const uint32_t b_lo_word = (((96u * 16u) >> 4) << 16);
    const uint32_t b_hi_word = hi_word | (128u >> 4);
    const uint32_t idesc_main = tcx::make_idesc(128, WLO ? 96 : 48), idesc_lo = tcx::make_idesc(128, 48);
    const uint32_t planes_base = tcx::smem_u32(planes), wsm_base = tcx::smem_u32(wsm);
    uint32_t it = 0;
    for (int k = blockIdx.x; k < P.ntiles; k += gridDim.x, ++it) {
      const uint32_t st = it & 1u, ph = (it >> 1) & 1u;
      tcx::mbar_wait(&acc_empty[st], ph ^ 1u);
      tcx::mbar_wait(&pl_full[st], ph);
      tcx::tc_fence_after();
      long long* tr = (P.trace && blockIdx.x == 0 && lane == 0 && it < 64) ? P.trace + it * 8 : nullptr;
      if (tr) tr[3] = clock64();
      const uint32_t tacc = tmem_base + st * TZ_ACC_STRIDE;
      const uint32_t a_hi = planes_base + st * 2 * TZ_PLANE, a_lo = a_hi + TZ_PLANE;
      if (tcx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 3 * TZ_KSTEPS; ++ks) {
          const uint32_t dyi = ks / TZ_KSTEPS, kk = ks % TZ_KSTEPS;
          const uint32_t aoff = 2 * kk * TZ_CS + dyi * (TZ_XB * 16);   // smem tile row = window row + 1 + dy
          const uint32_t wb = wsm_base + ks * TZ_WSTEP;
          const uint64_t d_ah = tcx::desc64((((a_hi + aoff) & 0x3FFFFu) >> 4) | a_lo_word, a_hi_word);
          const uint64_t d_al = tcx::desc64((((a_lo + aoff) & 0x3FFFFu) >> 4) | a_lo_word, a_hi_word);
          const uint64_t d_w = tcx::desc64(((wb & 0x3FFFFu) >> 4) | b_lo_word, b_hi_word);
          tcx::umma_f16(tacc, d_ah, d_w, idesc_main, ks ? 1u : 0u);
          tcx::umma_f16(tacc, d_al, d_w, idesc_lo, 1u);
        }
        tcx::umma_commit(&pl_empty[st]);
        tcx::umma_commit(&acc_full[st]);
      }
      __syncwarp();
      if (tr) tr[4] = clock64();
    }
  } else if (warp >= 8) {
    // ===== conversion: raw fp32 -> fp16 hi / lo in the window layout =====
    const int ct = threadIdx.x - 8 * 32;
    // Tasks: one 32-byte chunk (8 values) each.  Slot s = ct + 256 u; a row's first 96 chunks are dealt to 96 consecutive slots
    // in an order that makes BOTH the LDS.128 of the raw tile and the STS.128 into the window layout conflict-free within every
    // quarter warp: slot 8 k + p of a 48-chunk block takes window w = p with bits 1 and 2 swapped and chunk c = k (w < 4) or
    // k ^ 1 (w >= 4) -- 8 distinct windows (16-byte bank groups of the planes) and chunk indices 6 w + c that cover every
    // residue mod 4 twice (32-byte slots of the raw tile; lanes 4-7 of a quarter fetch their two halves in swapped order).
    // The last 3 chunks of each row (pixel 64, 65: only seen by the last window) are 30 left-over slots.
    constexpr int NMAIN = TZ_RH * 96, NTASK = NMAIN + TZ_RH * (TZ_ROWCH - 96), NIT = (NTASK + 32 * TZ_CONV_WARPS - 1) / (32 * TZ_CONV_WARPS);
    static_assert(TZ_ROWCH - 96 == 3 && TZ_CHS == 6 && TZ_XB == 16, "slot map is written for 16 windows of 6-chunk stride");
    const int hsw = (lane >> 2) & 1;
    uint32_t t_src[NIT], t_dst[NIT];   // byte offsets in the raw tile / in a plane; t_dst bit 0: first store, bit 1: second store
#pragma unroll
    for (int u = 0; u < NIT; ++u) {
      const int sl = ct + u * 32 * TZ_CONV_WARPS;
      int rr, xb, c;
      if (sl < NMAIN) {
        rr = sl / 96;
        const int pos = sl - rr * 96, blk = pos >= 48, l48 = pos - 48 * blk, k = l48 >> 3, pp = l48 & 7;
        const int w = (pp & 1) | ((pp & 2) << 1) | ((pp & 4) >> 1);
        xb = 8 * blk + w; c = w < 4 ? k : (k ^ 1);
      } else {
        const int r = sl - NMAIN;
        rr = r / 3; xb = TZ_XB; c = r - 3 * rr;
      }
      t_src[u] = (uint32_t)rr * (TZ_TXH * TZ_C1 * 4) + (uint32_t)(xb * TZ_CHS + c) * 32;
      t_dst[u] = ((uint32_t)c * TZ_CS + (uint32_t)rr * (TZ_XB * 16) + (uint32_t)xb * 16) | (xb < TZ_XB ? 1u : 0u) | ((c < TZ_NCH - TZ_CHS && xb >= 1) ? 2u : 0u);
      if (sl >= NTASK) { t_src[u] = 0; t_dst[u] = 0; }
    }
    uint32_t it = 0;
    for (int k = blockIdx.x; k < P.ntiles; k += gridDim.x, ++it) {
      const uint32_t st = it & 1u, ph = (it >> 1) & 1u;
      tcx::mbar_wait(&raw_full[st], ph);
      long long* tr = (P.trace && blockIdx.x == 0 && ct == 0 && it < 64) ? P.trace + it * 8 : nullptr;
      if (tr) tr[1] = clock64();
      const uint8_t* src = raw + st * TZ_RAW;
      uint8_t* dhi = planes + st * 2 * TZ_PLANE;
      uint8_t* dlo = dhi + TZ_PLANE;
      // Phase 1: raw tile -> registers -> hi / lo.  The raw stage is released as soon as it has been read, so that the TMA of
      // tile + 2 (latency ~2700 clk) is in flight while this tile is still being stored.
      uint4 hi[NIT], lo[NIT];
      {
        float4 in[NIT][2];
#pragma unroll
        for (int u = 0; u < NIT; ++u) {
          const float4* s4 = reinterpret_cast<const float4*>(src + t_src[u]);
          in[u][0] = s4[hsw]; in[u][1] = s4[1 - hsw];
        }
#pragma unroll
        for (int u = 0; u < NIT; ++u) tz_split8(hsw ? in[u][1] : in[u][0], hsw ? in[u][0] : in[u][1], hi[u], lo[u]);
      }
      tcx::fence_proxy_async();                            // order the reads before the async-proxy overwrite (next TMA)
      __syncwarp();
      if (lane == 0) tcx::mbar_arrive(&raw_empty[st]);     // generic-proxy reads of `raw` are complete (values consumed above)
      // Phase 2: hi / lo -> window layout of this stage's planes (once the MMAs of tile - 2 have finished reading them)
      tcx::mbar_wait(&pl_empty[st], ph ^ 1u);
#pragma unroll
      for (int u = 0; u < NIT; ++u) {
        const uint32_t o = t_dst[u] & ~15u;
        if (t_dst[u] & 1u) {
          *reinterpret_cast<uint4*>(dhi + o) = hi[u];
          *reinterpret_cast<uint4*>(dlo + o) = lo[u];
        }
        if (t_dst[u] & 2u) {                               // the same chunk seen from the window to the left
          const uint32_t o2 = o + TZ_CHS * TZ_CS - 16;
          *reinterpret_cast<uint4*>(dhi + o2) = hi[u];
          *reinterpret_cast<uint4*>(dlo + o2) = lo[u];
        }
      }
      tcx::fence_proxy_async();                          // generic writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) tcx::mbar_arrive(&pl_full[st]);
      if (tr) tr[2] = clock64();
    }
  } else if (warp >= 4) {
    // ===== epilogue: one thread = one window = 8 x 2 output pixels (24 contiguous bytes on each of two image rows) =====
    const int qd = warp & 3;
    const int n = qd * 32 + lane, r = n / TZ_XB, xb = n % TZ_XB;
    uint32_t it = 0;
    for (int k = blockIdx.x; k < P.ntiles; k += gridDim.x, ++it) {
      const uint32_t st = it & 1u, ph = (it >> 1) & 1u;
      int b, ty0, tx0;
      tile_origin(tile_of(k), b, ty0, tx0);
      tcx::mbar_wait(&acc_full[st], ph);
      tcx::tc_fence_after();
      long long* tr = (P.trace && blockIdx.x == 0 && qd == 0 && lane == 0 && it < 64) ? P.trace + it * 8 : nullptr;
      if (tr) tr[5] = clock64();
      uint32_t a0[32], a1[32], a2[32];
      const uint32_t taddr = tmem_base + st * TZ_ACC_STRIDE + ((uint32_t)(qd * 32) << 16);
      tcx::tmem_ld32_nowait(taddr, a0);
      tcx::tmem_ld32_nowait(taddr + 32, a1);
      if (WLO) tcx::tmem_ld32_nowait(taddr + 64, a2);
      tcx::tmem_ld_wait();
      tcx::tc_fence_before();
      __syncwarp();
      if (lane == 0) tcx::mbar_arrive(&acc_empty[st]);   // values are in registers
      float v[TZ_NOUT];
#pragma unroll
      for (int i = 0; i < TZ_NOUT; ++i) {
        float s = __uint_as_float(i < 32 ? a0[i] : a1[i - 32]);
        if (WLO) s += __uint_as_float(i + 48 < 64 ? a1[i + 48 - 32] : a2[i + 48 - 64]);
        const int co = i % 3;
        v[i] = fmaf(s, P.inv_scale, P.bias[co]);
      }
      const int ty = ty0 + r, x0 = tx0 + TZ_J * xb;
      if (ty >= P.hin || x0 >= P.win) { if (tr) tr[6] = clock64(); continue; }
      if (P.fast && x0 + TZ_J <= P.win && 2 * (x0 + TZ_J) <= P.W) {
#pragma unroll
        for (int phy = 0; phy < 2; ++phy) {
          const int oy = 2 * ty + phy;
          if (oy >= P.H) continue;
          uint32_t w6[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const float* p = v + phy * 24 + 4 * j;
            w6[j] = tz_pixel(p[0]) | (tz_pixel(p[1]) << 8) | (tz_pixel(p[2]) << 16) | (tz_pixel(p[3]) << 24);
          }
          uint2* o = reinterpret_cast<uint2*>(P.out_u8 + ((size_t)b * P.H + oy) * ((size_t)P.W * 3) + (size_t)x0 * 6);
          o[0] = make_uint2(w6[0], w6[1]); o[1] = make_uint2(w6[2], w6[3]); o[2] = make_uint2(w6[4], w6[5]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < TZ_NOUT; ++i) {
          const int phy = i / 24, rem = i % 24, jp = rem / 3, co = rem % 3;
          const int x = x0 + jp / 2, oy = 2 * ty + phy, ox = 2 * x + (jp & 1);
          if (x >= P.win) continue;
          if (P.out) P.out[(((size_t)b * P.hout + oy) * P.wout + ox) * 3 + co] = v[i];
          if (oy < P.H && ox < P.W) {
            const size_t qi = (((size_t)b * P.H + oy) * P.W + ox) * 3 + co;
            if (P.out_u8) P.out_u8[qi] = (uint8_t)tz_pixel(v[i]);
            if (P.out_crop) P.out_crop[qi] = v[i];
          }
        }
      }
      if (tr) tr[6] = clock64();
    }
  }
  tcx::tc_fence_before();
  __syncthreads();
  if (warp == 2) tcx::tmem_dealloc(tmem_base, 2 * TZ_ACC_STRIDE);
}

// ------------------------------------------------------------------------------------------------
// host side
struct TailTz {
  bool ok = false;
  float scale = 1.f;
  uint4* d_w = nullptr;
  float bias[3] = {0.f, 0.f, 0.f};
  bool attr_set = false;
};

inline bool tail_tz_supported(const ConvLayer& c) {
  return c.s == 2 && c.k == 5 && c.p == 1 && c.cout == 3 && !c.append_ones && c.cin == TZ_C1;
}

inline bool tail_tz_pack(const ConvLayer& c, const HostWeights& hw, TailTz& t, std::vector<void*>& owned, std::string* err) {
  float wmax = 0.f;
  for (auto& s : c.sources) for (float v : hw.at(s.kernel).second) wmax = std::max(wmax, std::fabs(v));
  int e = 0;
  if (wmax > 0.f) std::frexp(wmax, &e);
  t.scale = std::ldexp(1.f, 12 - e);                               // max |w| * S in [2^11, 2^12): the lo part stays normal
  std::vector<__half> w((size_t)TZ_WBYTES / 2, __float2half(0.f));
  for (int ks = 0; ks < 3 * TZ_KSTEPS; ++ks) {
    const int dy = ks / TZ_KSTEPS - 1, kk = ks % TZ_KSTEPS;
    for (int kc = 0; kc < 2; ++kc) for (int e8 = 0; e8 < 8; ++e8) {
      const int el = kk * 16 + kc * 8 + e8;                        // position in the window's 6-pixel run
      if (el >= (TZ_J + 2) * TZ_C1) continue;
      const int jj = el / TZ_C1, ci = el % TZ_C1;
      for (int nc = 0; nc < TZ_NOUT; ++nc) {
        const int phy = nc / 24, rem = nc % 24, jp = rem / 3, co = rem % 3, j = jp / 2, phx = jp % 2;
        const int jx = jj - 1 - j;                                 // input pixel - output cell
        if (jx < -1 || jx > 1) continue;
        const int ay = phy + c.p - 2 * dy, ax = phx + c.p - 2 * jx;   // o = 2 n + a - p
        if (ay < 0 || ay >= c.k || ax < 0 || ax >= c.k) continue;
        const float v = conv_w(c, hw, ay, ax, co, ci) * t.scale;
        const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
        const size_t base = (size_t)ks * (TZ_WSTEP / 2) + (size_t)kc * 96 * 8;
        w[base + (size_t)nc * 8 + e8] = h;
        w[base + (size_t)(48 + nc) * 8 + e8] = l;
      }
    }
  }
  if (cudaMalloc((void**)&t.d_w, TZ_WBYTES) != cudaSuccess) { *err = "cudaMalloc (window-GEMM tail weights) failed"; return false; }
  owned.push_back(t.d_w);
  if (cudaMemcpy(t.d_w, w.data(), TZ_WBYTES, cudaMemcpyHostToDevice) != cudaSuccess) { *err = "cudaMemcpy (window-GEMM tail weights) failed"; return false; }
  { std::vector<float> bv = pack_bias(c, hw); for (int i = 0; i < 3; ++i) t.bias[i] = bv[i]; }
  t.ok = true;
  return true;
}

struct TailTzOut { float* f32 = nullptr; uint8_t* u8 = nullptr; float* crop = nullptr; int H = 0, W = 0; };

inline int tail_tz_run(TcDriver& drv, const ConvLayer& c, TailTz& t, const float* in, int B, int h, int w, const TailTzOut& o, bool pdl,
                       cudaStream_t s, uint64_t* launches, std::string* err, bool w_lo = true) {
  if (!drv.encode) { *err = "cuTensorMapEncodeTiled unavailable"; return 2; }
  TailTzParams P{};
  P.B = B; P.hin = h; P.win = w;
  P.tiles_x = (w + TZ_TX - 1) / TZ_TX; P.tiles_y = (h + TZ_ROWS - 1) / TZ_ROWS; P.ntiles = P.tiles_x * P.tiles_y * B;
  if (P.ntiles <= 0) return 0;
  static const int reverse = tc_env_int("SNTC_TAIL_TZ_REVERSE", 1);
  P.reverse = reverse;
  P.inv_scale = 1.f / t.scale;
  for (int i = 0; i < 3; ++i) P.bias[i] = t.bias[i];
  P.out = o.f32; P.hout = 2 * h; P.wout = 2 * w; P.out_u8 = o.u8; P.out_crop = o.crop; P.H = o.H; P.W = o.W;
  P.w = t.d_w;
  P.fast = o.u8 && !o.f32 && !o.crop && (o.W % 8) == 0 && (reinterpret_cast<uintptr_t>(o.u8) & 7) == 0;
  CUtensorMap mapT;   // t [B,h,w,12] fp32 as (c, x, y, b); box = one halo tile
  {
    cuuint64_t dims[4] = {(cuuint64_t)TZ_C1, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)TZ_C1 * 4, (cuuint64_t)w * TZ_C1 * 4, (cuuint64_t)h * w * TZ_C1 * 4};
    cuuint32_t box[4] = {(cuuint32_t)TZ_C1, (cuuint32_t)TZ_TXH, (cuuint32_t)TZ_RH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = drv.encode(&mapT, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled(window-GEMM tail) failed: " + std::to_string((int)r); return 2; }
  }
  if (!t.attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tail_tz_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TZ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_tz_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TZ_SMEM);
    if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute(window-GEMM tail): ") + cudaGetErrorString(e); return 2; }
    t.attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(P.ntiles, drv.num_sms));
  cfg.blockDim = dim3(TZ_THREADS); cfg.dynamicSmemBytes = TZ_SMEM; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  static const bool trace_on = tc_env_int("SNTC_TC_TRACE", 0) != 0;
  long long* d_trace = nullptr;
  if (trace_on) { cudaMalloc((void**)&d_trace, 64 * 8 * 8); cudaMemsetAsync(d_trace, 0, 64 * 8 * 8, s); P.trace = d_trace; }
  cudaError_t e = w_lo ? cudaLaunchKernelEx(&cfg, tail_tz_kernel<true>, mapT, P) : cudaLaunchKernelEx(&cfg, tail_tz_kernel<false>, mapT, P);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("tail_tz_kernel launch: ") + cudaGetErrorString(e); return 2; }
  if (launches) (*launches)++;
  if (trace_on) {
    long long hbuf[64 * 8];
    cudaStreamSynchronize(s);
    cudaMemcpy(hbuf, d_trace, sizeof(hbuf), cudaMemcpyDeviceToHost);
    cudaFree(d_trace);
    fprintf(stderr, "[tc-trace] window-GEMM tail: tiles=%d grid=%d\n", P.ntiles, (int)cfg.gridDim.x);
    for (int j = 0; j < 32; ++j) {
      const long long* r = hbuf + j * 8; const long long t0 = hbuf[0];
      fprintf(stderr, "[tc-trace]  tile %2d: tma-issue %7lld  conv: start %7lld done %7lld  mma: data %7lld issued %7lld  epi: full %7lld done %7lld\n", j,
              r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0, r[4] - t0, r[5] - t0, r[6] - t0);
    }
  }
  return 0;
}

}  // namespace sntc
