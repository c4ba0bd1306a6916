// tcgen05 tail of the two-layer synthesis: ConvT(k<=6, s=2, C1 -> <=4 channels) + crop + uint8 epilogue on the
// tensor cores (common/transforms.py:311-313, 350-353 + image_utils.py:22-23, 69-71), sm_100a only.
//
// Cell form (sntc_plan.hpp): u = o + p = 2m + phi, a = phi + 2j, n = m - j, j in {0,1,2}.  All four phases of a
// cell read the same 3x3 input patch, so one GEMM row per cell produces its 2x2 output pixels:
//     D[cell m, (phi_y, phi_x, co)] = sum_{jy,jx,ci} t[m - j, ci] * W[phi + 2j, co, ci]        (N = 4 phases x 4 = 16)
// The FLOPs are tiny; what bounds this op is operand delivery.  Therefore:
//   * HALO REUSE: a CTA TMA-loads ONE (16+2)x(8+2) patch of the fp16 hi/lo planes of t per 128-cell tile into
//     un-swizzled, K-major shared memory laid out [y][k-octet][x][8 channels].  A core matrix (8 rows x 16 B)
//     is then 8 consecutive x-cells, and the A operand of tap (jy,jx) is the SAME patch addressed through a
//     descriptor whose start is shifted by (2-jy, 2-jx) cells (SBO = one patch row, LBO = one k-octet row):
//     9 taps read 11.5 KB of shared memory that was fetched once, instead of 9 separately fetched tiles.
//   * t is stored OCTET-PLANAR in HBM, [B][h][k-octet][w][8] (written by the layer-1 epilogue), so that the patch
//     rows are 160 contiguous bytes for TMA; with pixel-major [..][w][16] the box rows would be 16 bytes and the
//     copy engine, not the tensor pipe, would set the pace (measured: 2100 clk per tile).
//   * split-fp16 with TWO MMAs per tap instead of three: B holds [w_hi | w_lo] side by side (N = 32), so
//     A_hi x [w_hi | w_lo] is one instruction (the A-operand read is what costs); A_lo x w_hi (N = 16) is the second.
//     The epilogue adds the two column groups.  (lo*lo dropped as everywhere, < 2^-22 relative.)
//   * accumulators: 4 TMEM buffers of 32 columns; 8 epilogue warps (two per lane quarter, alternate tiles).
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue.
#pragma once
#include "sntc_kernels_tc.cuh"

namespace sntc {

constexpr int TT_TY = 16, TT_TX = 8;             // cells per tile (128 = UMMA M)
constexpr int TT_PY = TT_TY + 2, TT_PX = TT_TX + 2;
constexpr int TT_NBUF = 4;                       // TMEM accumulator buffers (32 columns each)
constexpr int TT_TAPS = 9;
constexpr uint32_t TT_ROW_BYTES = TT_PX * 16;   // one k-octet row of the patch: 10 cells x 8 fp16
constexpr uint32_t TT_WTAP_KQ_BYTES = 32 * 16;   // weights: per (tap, k-octet): 32 n-rows x 8 fp16

struct TailTcParams {
  int B, hin, win;            // t planes [B,hin,win,CP]
  int KQ;                     // k-octets = CP / 8 (even: a K=16 MMA step spans two octets)
  int p;                      // ConvT padding: o = 2m + phi - p
  int tiles_y, tiles_x, total_tiles;
  int stages;
  float inv_scale; float bias[4]; int cout;
  int hout, wout;
  float* out_f32; uint8_t* out_u8; float* out_crop; int H, W;
  const __half* w;            // [9 taps][KQ][32][8]: n < 16 -> hi of (phase, co), n >= 16 -> lo
  long long* trace;           // debug timeline (SNTC_TC_TRACE=1): block 0, [64 tiles][8]
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tail_tc_kernel(const __grid_constant__ CUtensorMap mapHi, const __grid_constant__ CUtensorMap mapLo, const TailTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t plane_bytes = ((uint32_t)(TT_PY * P.KQ) * TT_ROW_BYTES + 127u) & ~127u;   // [y][kq][x][8] of one plane
  const uint32_t stage_bytes = 2u * plane_bytes;                                           // hi plane | lo plane
  const uint32_t w_bytes = (uint32_t)TT_TAPS * (uint32_t)P.KQ * TT_WTAP_KQ_BYTES;
  uint8_t* wsm = smem + (size_t)P.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(wsm + w_bytes);
  uint64_t* empty_bar = full_bar + P.stages;
  uint64_t* tmem_full_bar = empty_bar + P.stages;      // [TT_NBUF]
  uint64_t* tmem_empty_bar = tmem_full_bar + TT_NBUF;  // [TT_NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + TT_NBUF);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { tcx::prefetch_tmap(&mapHi); tcx::prefetch_tmap(&mapLo); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < P.stages; ++i) { tcx::mbar_init(&full_bar[i], 1); tcx::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < TT_NBUF; ++i) { tcx::mbar_init(&tmem_full_bar[i], 1); tcx::mbar_init(&tmem_empty_bar[i], 4); }
    tcx::fence_barrier_init();
  }
  if (warp == 2) tcx::tmem_alloc(tmem_slot, TT_NBUF * 32);
  // weights -> shared memory (generic-proxy writes, made visible to the tensor-core (async) proxy by the fence)
  for (uint32_t i = threadIdx.x; i < w_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(P.w) + i);
  tcx::fence_proxy_async();
  tcx::tc_fence_before();
  __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = tcx::smem_u32(smem), wsm_base = tcx::smem_u32(wsm);
  const int tiles_per_img = P.tiles_y * P.tiles_x;

  if (warp == 0) {
    // ===== TMA producer: one haloed patch (hi + lo planes, KQ octets each) per tile =====
    uint32_t st = 0, ph = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img, tt = tile - b * tiles_per_img;
      const int ty = tt / P.tiles_x, tx = tt - ty * P.tiles_x;
      const int x0 = tx * TT_TX - 2, y0 = ty * TT_TY - 2;
      tcx::mbar_wait(&empty_bar[st], ph ^ 1u);
      { const int jt = (tile - (int)blockIdx.x) / (int)gridDim.x; if (P.trace && blockIdx.x == 0 && lane == 0 && jt < 64) P.trace[jt * 8 + 0] = clock64(); }
      if (tcx::elect_one()) {
        tcx::mbar_expect_tx(&full_bar[st], 2u * (uint32_t)(TT_PY * P.KQ) * TT_ROW_BYTES);
        uint8_t* sa = smem + (size_t)st * stage_bytes;
        tcx::tma_load_4d(sa, &mapHi, &full_bar[st], 8 * x0, 0, y0, b);                 // (8*x + c8, kq, y, b)
        tcx::tma_load_4d(sa + plane_bytes, &mapLo, &full_bar[st], 8 * x0, 0, y0, b);
      }
      __syncwarp();
      if (++st == (uint32_t)P.stages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // un-swizzled K-major descriptors: start | LBO (K-direction core-matrix stride) | SBO (8-row group stride)
    const uint32_t hi_word = (1u << 14);                                           // version 1, layout NONE
    const uint32_t a_lbo_sbo_lo = ((TT_ROW_BYTES >> 4) << 16);                      // low word: LBO at [16,30): next k-octet row
    const uint32_t a_hi_word = hi_word | (((uint32_t)P.KQ * TT_ROW_BYTES) >> 4);     // high word: SBO at [32,46): next patch row
    const uint32_t b_lbo_lo = ((TT_WTAP_KQ_BYTES >> 4) << 16);
    const uint32_t b_hi_word = hi_word | (uint32_t)(128 >> 4);
    const uint32_t idesc32 = tcx::make_idesc(128, 32), idesc16 = tcx::make_idesc(128, 16);
    const int ksteps = P.KQ / 2;
    uint32_t st = 0, ph = 0, j = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++j) {
      const uint32_t buf = j % TT_NBUF;
      long long* tr = (P.trace && blockIdx.x == 0 && lane == 0 && j < 64) ? P.trace + j * 8 : nullptr;
      tcx::mbar_wait(&tmem_empty_bar[buf], ((j / TT_NBUF) & 1u) ^ 1u);
      if (tr) tr[1] = clock64();
      tcx::mbar_wait(&full_bar[st], ph);
      tcx::tc_fence_after();
      if (tr) tr[2] = clock64();
      const uint32_t tacc = tmem_base + buf * 32;
      const uint32_t sa_hi = smem_base + st * stage_bytes, sa_lo = sa_hi + plane_bytes;
      if (tcx::elect_one()) {
        uint32_t acc = 0;
#pragma unroll
        for (int jy = 0; jy < 3; ++jy)
#pragma unroll
          for (int jx = 0; jx < 3; ++jx) {
            const uint32_t aoff = (uint32_t)(2 - jy) * (uint32_t)P.KQ * TT_ROW_BYTES + (uint32_t)(2 - jx) * 16;
            const uint32_t woff = (uint32_t)(jy * 3 + jx) * (uint32_t)P.KQ * TT_WTAP_KQ_BYTES;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t ah = sa_hi + aoff + (uint32_t)ks * 2 * TT_ROW_BYTES, al = sa_lo + aoff + (uint32_t)ks * 2 * TT_ROW_BYTES;
              const uint32_t wb = wsm_base + woff + (uint32_t)ks * 2 * TT_WTAP_KQ_BYTES;
              const uint64_t d_ah = tcx::desc64(((ah & 0x3FFFFu) >> 4) | a_lbo_sbo_lo, a_hi_word);
              const uint64_t d_al = tcx::desc64(((al & 0x3FFFFu) >> 4) | a_lbo_sbo_lo, a_hi_word);
              const uint64_t d_w = tcx::desc64(((wb & 0x3FFFFu) >> 4) | b_lbo_lo, b_hi_word);
              tcx::umma_f16(tacc, d_ah, d_w, idesc32, acc);      // hi * [w_hi | w_lo] -> columns [0,32)
              tcx::umma_f16(tacc, d_al, d_w, idesc16, 1u);       // lo * w_hi         -> columns [0,16)
              acc = 1;
            }
          }
        tcx::umma_commit(&empty_bar[st]);
        tcx::umma_commit(&tmem_full_bar[buf]);
      }
      __syncwarp();
      if (tr) tr[3] = clock64();
      if (++st == (uint32_t)P.stages) { st = 0; ph ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===== epilogue: one thread = one cell = 2x2 output pixels =====
    const int ew = warp & 3, eh = (warp - 4) >> 2;
    const int r = ew * 32 + lane, gy = r / TT_TX, gx = r % TT_TX;
    const int nc = P.cout;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++j) {
      if ((int)(j & 1u) != eh) continue;
      const uint32_t buf = j % TT_NBUF;
      const int b = tile / tiles_per_img, tt = tile - b * tiles_per_img;
      const int ty = tt / P.tiles_x, tx = tt - ty * P.tiles_x;
      const int my = ty * TT_TY + gy, mx = tx * TT_TX + gx;
      long long* tr = (P.trace && blockIdx.x == 0 && (warp & 3) == 0 && lane == 0 && j < 64) ? P.trace + j * 8 : nullptr;
      tcx::mbar_wait(&tmem_full_bar[buf], (j / TT_NBUF) & 1u);
      tcx::tc_fence_after();
      if (tr) tr[4] = clock64();
      uint32_t raw[32];
      tcx::tmem_ld32_nowait(tmem_base + buf * 32 + ((uint32_t)(ew * 32) << 16), raw);
      tcx::tmem_ld_wait();
      tcx::tc_fence_before();
      __syncwarp();
      if (lane == 0) tcx::mbar_arrive(&tmem_empty_bar[buf]);      // values are in registers: the buffer is free
      if (tr) tr[5] = clock64();
      if (my > P.hin || mx > P.win) { if (tr) tr[6] = clock64(); continue; }
#pragma unroll
      for (int fy = 0; fy < 2; ++fy) {
        const int oy = 2 * my + fy - P.p;
        if (oy < 0 || oy >= P.hout) continue;
#pragma unroll
        for (int fx = 0; fx < 2; ++fx) {
          const int ox = 2 * mx + fx - P.p;
          if (ox < 0 || ox >= P.wout) continue;
          const int n0 = (fy * 2 + fx) * 4;   // 4 columns per phase whatever cout is: compile-time register indices
          float v[4];
#pragma unroll
          for (int co = 0; co < 4; ++co)
            v[co] = co < nc ? fmaf(__uint_as_float(raw[n0 + co]) + __uint_as_float(raw[16 + n0 + co]), P.inv_scale, P.bias[co]) : 0.f;
          if (P.out_f32) {
            float* o = P.out_f32 + (((size_t)b * P.hout + oy) * P.wout + ox) * nc;
#pragma unroll
            for (int co = 0; co < 4; ++co) if (co < nc) o[co] = v[co];
          }
          if (oy < P.H && ox < P.W) {
            const size_t qi = (((size_t)b * P.H + oy) * P.W + ox) * nc;
            if (P.out_u8) {
#pragma unroll
              for (int co = 0; co < 4; ++co) if (co < nc) P.out_u8[qi + co] = float_to_pixel(v[co]);
            }
            if (P.out_crop) {
#pragma unroll
              for (int co = 0; co < 4; ++co) if (co < nc) P.out_crop[qi + co] = v[co];
            }
          }
        }
      }
      if (tr) tr[6] = clock64();
    }
  }
  tcx::tc_fence_before();
  __syncthreads();
  if (warp == 2) tcx::tmem_dealloc(tmem_base, TT_NBUF * 32);
}

// ------------------------------------------------------------------------------------------------
// host side (struct TailTc is declared in sntc_kernels_tc.cuh next to the model state)

// The tail runs on the tensor cores when it is a stride-2 transposed conv with at most 3 taps per dimension and at most
// 4 output channels, fed by fp16 planes with CP = cin rounded up to 16 channels.
inline bool tail_tc_supported(const ConvLayer& c) {
  return c.s == 2 && c.k <= 6 && c.cout <= 4 && !c.append_ones && c.cin <= 64;
}

inline bool tail_tc_pack(const ConvLayer& c, const HostWeights& hw, TailTc& t, std::vector<void*>& owned, std::string* err) {
  t.CP = (c.cin + 15) / 16 * 16;
  t.KQ = t.CP / 8;
  float wmax = 0.f;
  for (auto& s : c.sources) for (float v : hw.at(s.kernel).second) wmax = std::max(wmax, std::fabs(v));
  int e = 0;
  if (wmax > 0.f) std::frexp(wmax, &e);
  t.scale = std::ldexp(1.f, 12 - e);
  std::vector<__half> w((size_t)TT_TAPS * t.KQ * 32 * 8, __float2half(0.f));
  for (int jy = 0; jy < 3; ++jy) for (int jx = 0; jx < 3; ++jx)
    for (int fy = 0; fy < 2; ++fy) for (int fx = 0; fx < 2; ++fx) {
      const int ay = fy + 2 * jy, ax = fx + 2 * jx;
      if (ay >= c.k || ax >= c.k) continue;
      for (int co = 0; co < c.cout; ++co) for (int ci = 0; ci < c.cin; ++ci) {
        const float v = conv_w(c, hw, ay, ax, co, ci) * t.scale;
        const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
        const int n = (fy * 2 + fx) * 4 + co, kq = ci / 8, k8 = ci % 8;
        const size_t base = ((size_t)(jy * 3 + jx) * t.KQ + kq) * 32 * 8;
        w[base + (size_t)n * 8 + k8] = h;
        w[base + (size_t)(16 + n) * 8 + k8] = l;
      }
    }
  if (cudaMalloc((void**)&t.d_w, w.size() * sizeof(__half)) != cudaSuccess) { *err = "cudaMalloc (tail weights) failed"; return false; }
  owned.push_back(t.d_w);
  if (cudaMemcpy(t.d_w, w.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) { *err = "cudaMemcpy (tail weights) failed"; return false; }
  { std::vector<float> b = pack_bias(c, hw); for (int i = 0; i < c.cout && i < 4; ++i) t.bias[i] = b[i]; }
  const int stage_bytes = 2 * (((TT_PY * t.KQ) * (int)TT_ROW_BYTES + 127) / 128 * 128);
  const int fixed = 2048 + TT_TAPS * t.KQ * (int)TT_WTAP_KQ_BYTES + 512;
  t.stages = std::min(8, (200 * 1024 - fixed) / stage_bytes);
  if (t.stages < 2) { *err = "tail: not enough shared memory"; return false; }
  t.ok = true;
  return true;
}

inline bool tail_tc_make_map(TcDriver& drv, CUtensorMap* map, const void* base, int CP, int w, int h, int B, std::string* err) {
  // octet-planar planes [B][h][KQ][w][8]: dims (8*w, KQ, h, B)
  const int KQ = CP / 8;
  cuuint64_t dims[4] = {(cuuint64_t)w * 8, (cuuint64_t)KQ, (cuuint64_t)h, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)w * 16, (cuuint64_t)KQ * w * 16, (cuuint64_t)h * KQ * w * 16};
  cuuint32_t box[4] = {(cuuint32_t)TT_PX * 8, (cuuint32_t)KQ, (cuuint32_t)TT_PY, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = drv.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled(tail) failed: " + std::to_string((int)r); return false; }
  return true;
}

struct TailTcOut { float* f32 = nullptr; uint8_t* u8 = nullptr; float* crop = nullptr; int H = 0, W = 0; };

inline int tail_tc_run(TcDriver& drv, const ConvLayer& c, TailTc& t, const float* h_bias, const __half* in_hi, const __half* in_lo, int B, int h, int w,
                       const TailTcOut& o, cudaStream_t s, uint64_t* launches, std::string* err) {
  CUtensorMap mapHi, mapLo;
  if (!tail_tc_make_map(drv, &mapHi, in_hi, t.CP, w, h, B, err)) return TC_ERROR;
  if (!tail_tc_make_map(drv, &mapLo, in_lo, t.CP, w, h, B, err)) return TC_ERROR;
  TailTcParams P{};
  P.B = B; P.hin = h; P.win = w; P.KQ = t.KQ; P.p = c.p;
  P.tiles_y = (h + 1 + TT_TY - 1) / TT_TY; P.tiles_x = (w + 1 + TT_TX - 1) / TT_TX;
  P.total_tiles = P.tiles_y * P.tiles_x * B;
  P.stages = t.stages;
  P.inv_scale = 1.f / t.scale; P.cout = c.cout;
  for (int i = 0; i < 4; ++i) P.bias[i] = i < c.cout ? h_bias[i] : 0.f;
  P.hout = 2 * h; P.wout = 2 * w;
  P.out_f32 = o.f32; P.out_u8 = o.u8; P.out_crop = o.crop; P.H = o.H; P.W = o.W;
  P.w = t.d_w;
  size_t smem = (size_t)t.stages * 2 * (((size_t)(TT_PY * t.KQ) * TT_ROW_BYTES + 127) / 128 * 128) + (size_t)TT_TAPS * t.KQ * TT_WTAP_KQ_BYTES + 1024 + 512;
  if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM (TMEM allocation is per CTA)
  if (!t.attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tail_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute(tail): ") + cudaGetErrorString(e); return TC_ERROR; }
    t.attr_set = true;
  }
  int grid = std::min(drv.num_sms, P.total_tiles);
  if (grid <= 0) return TC_OK;
  static const bool trace_on = tc_env_int("SNTC_TC_TRACE", 0) != 0;
  long long* d_trace = nullptr;
  if (trace_on) { cudaMalloc((void**)&d_trace, 64 * 8 * 8); cudaMemsetAsync(d_trace, 0, 64 * 8 * 8, s); P.trace = d_trace; }
  tail_tc_kernel<<<grid, TC_THREADS, smem, s>>>(mapHi, mapLo, P);
  if (launches) (*launches)++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("tail_tc_kernel launch: ") + cudaGetErrorString(e); return TC_ERROR; }
  if (trace_on) {
    long long hbuf[64 * 8];
    cudaStreamSynchronize(s);
    cudaMemcpy(hbuf, d_trace, sizeof(hbuf), cudaMemcpyDeviceToHost);
    cudaFree(d_trace);
    fprintf(stderr, "[tc-trace] tail: tiles=%d grid=%d stages=%d KQ=%d\n", P.total_tiles, grid, t.stages, t.KQ);
    for (int j = 0; j < 24; ++j) {
      const long long* r = hbuf + j * 8; const long long t0 = hbuf[0];
      fprintf(stderr, "[tc-trace]  tile %2d: tma-issue %7lld  mma: acc-free %7lld data %7lld issued %7lld  epi: full %7lld ld %7lld done %7lld\n", j,
              r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0, r[4] - t0, r[5] - t0, r[6] - t0);
    }
  }
  return TC_OK;
}

}  // namespace sntc
