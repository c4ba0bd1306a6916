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
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "sntc_plan.hpp"
#include "sntc_kernels_f32.cuh"
#include "sntc_kernels_rate.cuh"

namespace sntc {

enum { TC_OK = 0, TC_NOT_HANDLED = 1, TC_ERROR = 2 };
enum { TC_EPI_PLAIN = 0, TC_EPI_HYPER_FINAL = 1, TC_EPI_TWO_LAYER = 2, TC_EPI_COL2IM = 3 };

constexpr int TC_BM = 128;        // cells per tile (UMMA M)
constexpr int TC_BK = 64;         // fp16 channels per k-block (128 bytes = one swizzle row)
constexpr int TC_EPI_WARPS = 8;    // epilogue warps 4..11: two per TMEM lane quarter (they split the columns)
constexpr int TC_THREADS = 128 + 32 * TC_EPI_WARPS;
constexpr int TC_ACC_COLS = 256;  // TMEM columns per accumulator buffer (two buffers = all 512 columns)
constexpr int TC_MAX_BANDS = 16;
constexpr int TC_TRACE_ITEMS = 64;

// One band of a layer, resident in device memory (tensor maps must be 64-byte aligned).
struct alignas(64) TcBandDev {
  CUtensorMap mapBhi, mapBlo;     // packed K-major band matrix, hi / lo fp16 planes
  int phy0, nphx, phx0, Ty, Tx, mloy, mlox;
  int N, BN, ntiles;
  int item_begin;                 // n-tiles of all bands before this one (static).  First work item of the band: item_begin * m-tile groups
                                  // (band-major order) or slot item_begin of every group (group-major order)
  int oshift;                     // added to the output coordinate (merged bands: s*dlo + p, see ConvLayer::merged)
  int pad_[4];
};

struct TcParams {
  const TcBandDev* bands; int nbands; int total_items;
  int snake;          // tc_unit_item: alternate the direction of successive waves (band-major order only)
  int ipg;            // > 0: items are m-tile-group major (item = group * ipg + slot, bands[].item_begin = first slot of the band:
                      // a unit walks through all bands, so heavy-MMA and heavy-epilogue items alternate); 0: band major
                      // (items of band i start at bands[i].item_begin * groups).  The band table itself never depends on the batch.
  int B, hin, win, s, p;
  int TH, TW, tiles_y, tiles_x;
  int TB;             // images per m-tile (TH * TW * TB = 128): small latent grids pack several images into one tile (tc_choose_patch)
  int tile_step_y, tile_step_x, tile_off;   // m-tile (ty, tx) starts at cell (ty * step_y + off, tx * step_x + off): TH / TW / 0, or overlapping tiles (col2im)
  // TC_EPI_COL2IM (ConvLayer::col2im): the GEMM is the per-input-pixel contraction P[n, (g, a_y, a_x)] of a final layer ConvT(k, s, p) to
  // c2i_cout channels, c2i_kp columns per channel; cout / bias below are those of the FINAL layer, hout = hin * s
  int c2i_k, c2i_s, c2i_p, c2i_kp;
  int kblocks;        // 64-channel blocks per tap
  int last_kmma;      // MMAs (K=16 each) in the last block of a tap (1..4)
  int cout, bn_max, stages;
  int mtiles;         // m-tiles per band (tiles_y*tiles_x*B); a CTA pair takes m-tiles (2i, 2i+1)
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
  // TC_EPI_TWO_LAYER: columns of one output pixel = base[0,C1) (|| res[C1,2C1)); out_f32 = t [B,hout,wout,C1]
  int C1, has_res, tl_act, tl_inverse; const float* gamma; int gamma_stride; const float* beta;
  double* rate_slots; int* rate_slot_img; RateConst rc;   // TC_EPI_HYPER_FINAL: per-(item, CTA, epilogue warp) partial bits_y
  float* sigma_out;   // TC_EPI_HYPER_FINAL: raw sigma [B,hout,wout,Cy] for the stand-alone rate kernel (nullable)
  // TC_EPI_PLAIN extras for the GDN stages of the deep decoders (transforms.py:8-63, tfc.GDN):
  //   plane_xform: the fp16 planes receive f(x) (A_ABS / A_SQUARE = the GDN pooling input) while out_f32 receives x;
  //   gdn_mode: this GEMM *is* the GDN norm pool (1x1, gamma): out = gx * norm | gx / norm, norm = acc + beta (or its sqrt)
  int plane_xform, gdn_mode; const float* gx;
  int sign_in_lo;                            // plane_xform == A_ABS: the lo plane carries sign(x) in its mantissa LSB (store16_planes_abs_sign)
  const __half* gx_hi; const __half* gx_lo;  // gdn_mode with gx == nullptr: x = +-(hi + lo) from such planes (the GEMM's own A operand)
  int vec16;          // fast epilogue: cout % 16 == 0, Cy % 16 == 0 and every epilogue tensor 32-byte aligned
  int rgb_runs;       // final layer to 3 channels whose only destination is the uint8 image: byte-run epilogue (tc_epi_rgb_chunk)
  // TC_EPI_TWO_LAYER constants as kernel parameters: with the loops over (i, j) fully unrolled every gamma / beta / bias
  // is a constant-bank operand of its FFMA (no shared-memory loads in the per-pixel IGDN)
  float tl_gamma[24 * 24]; float tl_beta[24]; float tl_bias[48];
  long long* trace;   // debug timeline (SNTC_TC_TRACE=1): [unit][TC_TRACE_ITEMS][8] clock64 stamps, leader CTA only
  // Which of the two cross terms of the split product are issued (hi*hi always is): bit 0 = a_lo * w_hi, bit 1 = a_hi * w_lo.
  // 3 = the fp32-class 3-pass product (default).  A plane that is not needed is not loaded either.
  // alo_flag (nullable, device): written by split_planes_kernel, non-zero iff some element of the A lo plane is non-zero;
  // when it reads 0 (integer-valued input such as z_hat: exact in fp16) bit 0 is cleared -- the result is bit-identical.
  unsigned pass_mask; const unsigned* alo_flag;
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
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
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

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                 "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
// 256-bit global accesses (sm_100): one full 32-byte sector per lane and request
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void ldg256_nc(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of the (converged) warp; the surrounding code stays warp-uniform, so the compiler keeps the
// operands of the TMA / MMA instructions in uniform registers instead of broadcasting them per instruction
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Programmatic dependent launch: let the next kernel of the stream be scheduled while this one is still running
// (its prologue overlaps our tail), and wait for the previous kernel's results before touching them.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- thread-block cluster / CTA-pair (cta_group::2) forms ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's smem, the transaction bytes are counted on the barrier
// at cluster address `bar_cluster` (the leader CTA's full barrier)
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
    "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
    ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
    "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
    ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both: 128 rows each] * B[smem of both: N/2 rows each]; issued by the leader CTA
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
    ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_f16_t(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 2) umma_f16_2sm(tmem_d, adesc, bdesc, idesc, accumulate); else umma_f16(tmem_d, adesc, bdesc, idesc, accumulate);
}
// commit of the pair's MMAs: arrives on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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
// Epilogue helpers (one thread = one cell row of the accumulator tile)

struct TcItem {
  int band, b, iy0, ix0, n0, nrows, mma_n;
  bool dup;     // second CTA of a pair past the last m-tile: participates in the MMA, stores nothing
};

// CG = CTAs per work item (1, or 2 for a cta_group::2 pair).  Items of a band: (m-tile group, n-tile), n-tile fastest.
// k-th work item of persistent unit `unit0`.  snake != 0: successive waves run in opposite directions -- the items are sorted
// heaviest band first, so the unit that drew the heaviest item of one wave draws the lightest of the next (hyper layer 0 at
// batch 24: 96 items on 74 units, longest unit 45 + 20 -> 30 + 20 k-blocks).  >= total_items: the unit is done (a partial wave
// is the last one).  Which unit computes an item never changes its arithmetic.
__device__ __forceinline__ int tc_unit_item(int unit0, int nunits, int k, int snake) {
  return k * nunits + ((snake && (k & 1)) ? nunits - 1 - unit0 : unit0);
}

template <int CG>
__device__ __forceinline__ TcItem tc_decode_item(const TcParams& P, int item, int rank) {
  TcItem it;
  int bi = 0, nt, mg;
  // unsigned divisions (no sign fix-up code); the per-item decode matters for layers with short items (col2im final layer)
  if (P.ipg > 0) {
    mg = (int)((unsigned)item / (unsigned)P.ipg);
    const int slot = item - mg * P.ipg;
    for (int i = 1; i < P.nbands; ++i)
      if (slot >= P.bands[i].item_begin) bi = i;
    nt = slot - P.bands[bi].item_begin;
  } else if (P.nbands == 1 && P.bands[0].ntiles == 1) {
    nt = 0; mg = item;
  } else {
    const int groups = (P.mtiles + CG - 1) / CG;
    for (int i = 1; i < P.nbands; ++i)
      if (item >= P.bands[i].item_begin * groups) bi = i;
    const unsigned local = (unsigned)(item - P.bands[bi].item_begin * groups), ntl = (unsigned)P.bands[bi].ntiles;
    mg = (int)(local / ntl); nt = (int)(local - (unsigned)mg * ntl);
  }
  const TcBandDev& bd = P.bands[bi];
  int mt = mg * CG + rank;
  it.dup = mt >= P.mtiles;
  if (it.dup) mt = P.mtiles - 1;
  const unsigned q1 = (unsigned)mt / (unsigned)P.tiles_x, bq = q1 / (unsigned)P.tiles_y;
  const int tx = mt - (int)q1 * P.tiles_x, ty = (int)(q1 - bq * (unsigned)P.tiles_y);
  it.band = bi;
  it.b = (int)bq * P.TB;   // first image of the tile
  it.iy0 = ty * P.tile_step_y + P.tile_off; it.ix0 = tx * P.tile_step_x + P.tile_off;
  it.n0 = nt * bd.BN;
  it.nrows = min(bd.BN, bd.N - it.n0);
  it.mma_n = (it.nrows + 15) & ~15;
  return it;
}

// plain / hyper-final epilogue for 8 consecutive columns co..co+7 of output pixel (b, oy, ox)  (cout % 8 == 0)
__device__ __forceinline__ void tc_epi_vec8(const TcParams& P, const float* sbias, int b, int oy, int ox, int co, float* v) {
  if (oy < 0 || oy >= P.hout || ox < 0 || ox >= P.wout) return;
  const size_t pix = ((size_t)b * P.hout + oy) * P.wout + ox;
  float4 b0 = *reinterpret_cast<const float4*>(sbias + co);       // bias staged in shared memory
  float4 b1 = *reinterpret_cast<const float4*>(sbias + co + 4);
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
    } else if (P.idx) {   // sigma half: idx = round(clamp(exp(sigma), 0, S-1))   :274-276
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
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = apply_act(v[i], P.act);
  if (P.out_hi) store8_planes(P.out_hi, P.out_lo, pix * P.cout + co, v);
  if (P.out_f32) {
    float* o = P.out_f32 + pix * P.cout + co;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// 16 consecutive columns co..co+15 (cout % 16 == 0, Cy % 16 == 0, all tensors 32-byte aligned): every global access
// is a full 32-byte sector (256-bit loads / stores).
__device__ __forceinline__ void store16_planes(__half* hi, __half* lo, size_t off, const float* v) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {   // split_f16 on pairs: packed conversions (one F2FP per two values each way), same roundings
    const float c0 = fminf(fmaxf(v[2 * i], -65504.f), 65504.f), c1 = fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f);
    const __half2 hh = __floats2half2_rn(c0, c1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(c0 - hf.x, c1 - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  tcx::stg256(hi + off, h);
  tcx::stg256(lo + off, l);
}
// Planes of |v| whose lo halves carry sign(v) in their mantissa LSB (TcParams::sign_in_lo): the GDN stage that follows pools |x| on
// the tensor cores and rebuilds x = +-(hi + lo) in its epilogue from the very planes its TMA loads just pulled through L2, so x
// never makes a separate fp32 round trip through HBM.  The LSB moves |x| by <= 1 ulp(lo) <= 2^-21 |x| (the split product already
// drops lo*lo, 2^-22).
__device__ __forceinline__ void store16_planes_abs_sign(__half* hi, __half* lo, size_t off, const float* v) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float c0 = fminf(fabsf(v[2 * i]), 65504.f), c1 = fminf(fabsf(v[2 * i + 1]), 65504.f);
    const __half2 hh = __floats2half2_rn(c0, c1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(c0 - hf.x, c1 - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = (*reinterpret_cast<const uint32_t*>(&ll) & 0xFFFEFFFEu) | (__float_as_uint(v[2 * i]) >> 31) | ((__float_as_uint(v[2 * i + 1]) >> 31) << 16);
  }
  tcx::stg256(hi + off, h);
  tcx::stg256(lo + off, l);
}
// x[16] back from such planes
__device__ __forceinline__ void load16_planes_abs_sign(const __half* hi, const __half* lo, size_t off, float* x) {
  uint32_t h[8], l[8];
  tcx::ldg256_nc(hi + off, h);
  tcx::ldg256_nc(lo + off, l);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i])), lf = __half22float2(*reinterpret_cast<const __half2*>(&l[i]));
    const float a0 = hf.x + lf.x, a1 = hf.y + lf.y;
    x[2 * i] = (l[i] & 1u) ? -a0 : a0;
    x[2 * i + 1] = (l[i] & 0x10000u) ? -a1 : a1;
  }
}
__device__ __forceinline__ void store16_f32(float* o, const float* v) {
  tcx::stg256(o, reinterpret_cast<const uint32_t*>(v));
  tcx::stg256(o + 8, reinterpret_cast<const uint32_t*>(v) + 8);
}
__device__ __forceinline__ void load_q16(const void* q, int kind, size_t e, float* out) {
  if (kind == 3) {   // no symbols yet (phase 1 of the two-phase decode)
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = 0.f;
  } else if (kind == 0) {
    tcx::ldg256_nc(reinterpret_cast<const float*>(q) + e, reinterpret_cast<uint32_t*>(out));
    tcx::ldg256_nc(reinterpret_cast<const float*>(q) + e + 8, reinterpret_cast<uint32_t*>(out) + 8);
  } else if (kind == 1) {
    uint32_t r[8];
    tcx::ldg256_nc(reinterpret_cast<const int16_t*>(q) + e, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) { out[2 * i] = (float)(int16_t)(r[i] & 0xFFFFu); out[2 * i + 1] = (float)(int16_t)(r[i] >> 16); }
  } else {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const int8_t*>(q) + e));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      out[4 * i] = (float)(int8_t)(w[i] & 0xFFu); out[4 * i + 1] = (float)(int8_t)((w[i] >> 8) & 0xFFu);
      out[4 * i + 2] = (float)(int8_t)((w[i] >> 16) & 0xFFu); out[4 * i + 3] = (float)(int8_t)(w[i] >> 24);
    }
  }
}
// The band fields the epilogue needs, copied to registers once per work item: the band table lives in global
// memory and the epilogue's own stores would otherwise force a reload (possible aliasing) per column group.
struct TcBandRegs { int N, nphx, phy0, phx0, oshift; };

// GDN norm pool (1x1 GEMM over gamma), 16 columns of one pixel: out = x * norm | x / norm, norm = acc + beta or its root, with
// the mode a COMPILE-TIME constant.  The generic tc_epi_vec16 tests act / gdn_mode / plane_xform per element (~40
// instructions per element); with only 8 epilogue warps per SM that made the stage issue-latency bound (mbt2018 igdn_2:
// 0.24 IPC per scheduler, 10 % tensor pipe, 29 % of HBM: profiles/r01_ncu_full_mbt2018_gdn2_layer3.txt).
template <int GM>
__device__ __forceinline__ void tc_epi_gdn16(const TcParams& P, const float* sbias, size_t off, int co, const uint32_t* raw) {
  float x[16], v[16];
  if (P.gx) {
    tcx::ldg256_nc(P.gx + off, reinterpret_cast<uint32_t*>(x));
    tcx::ldg256_nc(P.gx + off + 8, reinterpret_cast<uint32_t*>(x) + 8);
  } else {
    load16_planes_abs_sign(P.gx_hi, P.gx_lo, off, x);   // the pooling planes themselves (L2-hot: this CTA's TMA just loaded them)
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 bb = *reinterpret_cast<const float4*>(sbias + co + 4 * g);
    v[4 * g] = fmaf(__uint_as_float(raw[4 * g]), P.inv_scale, bb.x); v[4 * g + 1] = fmaf(__uint_as_float(raw[4 * g + 1]), P.inv_scale, bb.y);
    v[4 * g + 2] = fmaf(__uint_as_float(raw[4 * g + 2]), P.inv_scale, bb.z); v[4 * g + 3] = fmaf(__uint_as_float(raw[4 * g + 3]), P.inv_scale, bb.w);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    // sqrt.approx / div.approx (<= 2 ulp): two orders of magnitude below the split-fp16 accumulation error of this path
    float nrm = v[i];
    if (GM == G_MUL_SQRT || GM == G_DIV_SQRT) asm("sqrt.approx.f32 %0, %1;" : "=f"(nrm) : "f"(fmaxf(v[i], 0.f)));
    v[i] = (GM == G_MUL || GM == G_MUL_SQRT) ? x[i] * nrm : __fdividef(x[i], nrm);
  }
  if (P.out_f32) store16_f32(P.out_f32 + off, v);
  if (P.out_hi) store16_planes(P.out_hi, P.out_lo, off, v);
}

// All chunks of one work item of a GDN stage (act = none, no plane transform: checked by the caller).
template <int GM>
__device__ __forceinline__ void tc_epi_gdn_item(const TcParams& P, const TcBandRegs& bd, const TcItem& it, const float* sbias, uint32_t trow,
                                                int my, int mx, bool cell_ok, int eh, int EH) {
  uint32_t raw[32], nxt[32];
  int c = 32 * eh;
  if (c < it.mma_n) tcx::tmem_ld32_nowait(trow + (uint32_t)c, nxt);
  for (; c < it.mma_n; c += 32 * EH) {
    tcx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) raw[i] = nxt[i];
    if (c + 32 * EH < it.mma_n) tcx::tmem_ld32_nowait(trow + (uint32_t)(c + 32 * EH), nxt);
    if (!cell_ok) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = it.n0 + c + 16 * h;
      if (n >= bd.N) break;
      const int ph = n / P.cout, co = n - ph * P.cout;
      const int oy = P.s * my + bd.phy0 + ph / bd.nphx - P.p + bd.oshift, ox = P.s * mx + bd.phx0 + ph % bd.nphx - P.p + bd.oshift;
      if (oy < 0 || oy >= P.hout || ox < 0 || ox >= P.wout) continue;
      tc_epi_gdn16<GM>(P, sbias, (((size_t)it.b * P.hout + oy) * P.wout + ox) * P.cout + co, co, raw + 16 * h);
    }
  }
}

__device__ __forceinline__ float tc_epi_vec16(const TcParams& P, const float* sbias, int b, int oy, int ox, int co, const uint32_t* raw) {
  if (oy < 0 || oy >= P.hout || ox < 0 || ox >= P.wout) return 0.f;
  const size_t pix = ((size_t)b * P.hout + oy) * P.wout + ox;
  float v[16];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 bb = *reinterpret_cast<const float4*>(sbias + co + 4 * g);   // bias staged in shared memory
    v[4 * g] = fmaf(__uint_as_float(raw[4 * g]), P.inv_scale, bb.x); v[4 * g + 1] = fmaf(__uint_as_float(raw[4 * g + 1]), P.inv_scale, bb.y);
    v[4 * g + 2] = fmaf(__uint_as_float(raw[4 * g + 2]), P.inv_scale, bb.z); v[4 * g + 3] = fmaf(__uint_as_float(raw[4 * g + 3]), P.inv_scale, bb.w);
  }
  if (P.epi == TC_EPI_HYPER_FINAL) {
    if (P.out_f32) store16_f32(P.out_f32 + pix * P.cout + co, v);
    if (co < P.Cy) {   // mu half: y_hat = q + mu                     mshyper/models.py:278
      const size_t e = pix * P.Cy + co;
      float y[16];
      load_q16(P.q, P.q_kind, e, y);
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = __fadd_rn(y[i], v[i]);
      if (P.y_hat) store16_f32(P.y_hat + e, y);
      if (P.out_hi) store16_planes(P.out_hi, P.out_lo, e, y);
    } else if (P.idx) {   // sigma half: idx = round(clamp(exp(sigma), 0, S-1))   :274-276
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o[i] = (uint32_t)scale_index(v[4 * i], P.max_index, P.trunc) | ((uint32_t)scale_index(v[4 * i + 1], P.max_index, P.trunc) << 8) |
               ((uint32_t)scale_index(v[4 * i + 2], P.max_index, P.trunc) << 16) | ((uint32_t)scale_index(v[4 * i + 3], P.max_index, P.trunc) << 24);
      *reinterpret_cast<uint4*>(P.idx + pix * P.Cy + (co - P.Cy)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (co >= P.Cy && P.sigma_out) store16_f32(P.sigma_out + pix * P.Cy + (co - P.Cy), v);
    float bits = 0.f;
    if (co >= P.Cy && P.rate_slots) {   // rate term a7: bits of q under NoisyNormal(scale = SCALE_FN(i_c))   :278-279
      float qv[16];
      load_q16(P.q, P.q_kind, pix * P.Cy + (co - P.Cy), qv);
#pragma unroll
      for (int i = 0; i < 16; ++i) bits += noisy_normal_bits(qv[i], v[i], P.rc);
    }
    return bits;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], P.act);
  if (P.gdn_mode != G_NONE) {   // v = beta + f(x) gamma: scale the layer input x by the norm
    float x[16];
    if (P.gx) {
      tcx::ldg256_nc(P.gx + pix * P.cout + co, reinterpret_cast<uint32_t*>(x));
      tcx::ldg256_nc(P.gx + pix * P.cout + co + 8, reinterpret_cast<uint32_t*>(x) + 8);
    } else {
      load16_planes_abs_sign(P.gx_hi, P.gx_lo, pix * P.cout + co, x);
    }
    const bool root = P.gdn_mode == G_MUL_SQRT || P.gdn_mode == G_DIV_SQRT, mul = P.gdn_mode == G_MUL || P.gdn_mode == G_MUL_SQRT;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float nrm = root ? sqrtf(fmaxf(v[i], 0.f)) : v[i];
      v[i] = mul ? x[i] * nrm : x[i] / nrm;
    }
  }
  if (P.out_f32) store16_f32(P.out_f32 + pix * P.cout + co, v);
  if (P.out_hi) {
    if (P.plane_xform == A_ABS && P.sign_in_lo) {
      store16_planes_abs_sign(P.out_hi, P.out_lo, pix * P.cout + co, v);
      return 0.f;
    } else if (P.plane_xform == A_ABS) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fabsf(v[i]);
    } else if (P.plane_xform == A_SQUARE) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = v[i] * v[i];
    }
    store16_planes(P.out_hi, P.out_lo, pix * P.cout + co, v);
  }
  return 0.f;
}


// generic scalar epilogue (final layers with cout = 3, ...)
__device__ __forceinline__ void tc_epi_scalar8(const TcParams& P, const TcBandRegs& bd, const float* sbias, int b, int my, int mx, int n, const float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int nn = n + i;
    if (nn >= bd.N) break;
    const int co = nn % P.cout, ph = nn / P.cout;
    const int oy = P.s * my + bd.phy0 + ph / bd.nphx - P.p + bd.oshift, ox = P.s * mx + bd.phx0 + ph % bd.nphx - P.p + bd.oshift;
    if (oy < 0 || oy >= P.hout || ox < 0 || ox >= P.wout) continue;
    float x = apply_act(fmaf(v[i], P.inv_scale, sbias[co]), P.act);
    const size_t pix = ((size_t)b * P.hout + oy) * P.wout + ox;
    if (P.out_f32) P.out_f32[pix * P.cout + co] = x;
    if ((P.out_u8 || P.out_crop) && oy < P.H && ox < P.W) {
      const size_t qi = (((size_t)b * P.H + oy) * P.W + ox) * P.cout + co;
      if (P.out_u8) P.out_u8[qi] = float_to_pixel(x);
      if (P.out_crop) P.out_crop[qi] = x;
    }
  }
}

// ---- byte-run epilogue of a final layer with 3 output channels (JPEG-like synthesis: ConvT(18, 16, 320 -> 3)) ----
// The columns of a band are (phase_y, phase_x, co): for one cell and one phase_y, the nphx * 3 columns are CONSECUTIVE BYTES of
// one image row.  A thread therefore turns its 32-column accumulator chunk into 32 bytes in registers and writes them as (at
// most two) byte runs with 32-bit stores: the words are re-aligned to the destination with funnel shifts, only the first and
// last word of a run fall back to byte stores.  The scalar path this replaces issues one 1-byte store and four integer
// divisions per element (0.336 ms per 24-image step for 28 MB of pixels: L2 sector writes and issue slots, not HBM).
//
// Elements [lo, hi) of the 32 bytes packed little-endian in w[0..7] go to dst0 + i (dst0 = address of element 0).
__device__ __forceinline__ void tc_store_byte_run(uint8_t* dst0, const uint32_t* w, int lo, int hi) {
  if (lo >= hi) return;
  const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(dst0) & 3u);
  uint8_t* base = dst0 - a;                                   // 4-byte aligned; word k of the destination holds elements 4k - a .. 4k - a + 3
  const uint32_t sh = 8u * a;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const uint32_t wl = k > 0 ? w[k - 1] : 0u, wh = k < 8 ? w[k] : 0u;
    const uint32_t v = __funnelshift_l(wl, wh, sh);           // bytes of elements 4k - a + (0..3)
    const int e0 = 4 * k - (int)a;
    if (e0 >= lo && e0 + 3 < hi) {
      *reinterpret_cast<uint32_t*>(base + 4 * k) = v;
    } else if (e0 + 3 >= lo && e0 < hi) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e0 + j >= lo && e0 + j < hi) base[4 * k + j] = (uint8_t)(v >> (8 * j));
    }
  }
}

// One 32-column chunk (columns n_first .. n_first + 31 of the band) of one cell: requires nphx * 3 >= 32 (at most one row break).
__device__ __forceinline__ void tc_epi_rgb_chunk(const TcParams& P, const TcBandRegs& bd, const float* sbias, int b, int my, int mx, int n_first,
                                                 int n_end, const uint32_t* raw) {
  const int nvalid = min(32, n_end - n_first);                 // n_end: first column past this work item's n-tile
  if (nvalid <= 0) return;
  const int ph0 = n_first / 3, co0 = n_first - 3 * ph0;
  const int py0 = ph0 / bd.nphx, px0 = ph0 - py0 * bd.nphx;
  const float br[3] = {sbias[co0], sbias[co0 == 2 ? 0 : co0 + 1], sbias[co0 == 0 ? 2 : co0 - 1]};   // bias of element i: br[i % 3]
  uint32_t w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = 4 * k + j;
      const float x = apply_act(fmaf(__uint_as_float(raw[i]), P.inv_scale, br[i % 3]), P.act);
      word |= (uint32_t)float_to_pixel(x) << (8 * j);
    }
    w[k] = word;
  }
  const int Hlim = min(P.hout, P.H), Wb = min(P.wout, P.W) * 3;      // crop (image_utils.py:69-71)
  const int lenA = min(nvalid, (bd.nphx - px0) * 3 - co0);          // elements before the row break
  const int oyA = P.s * my + bd.phy0 + py0 - P.p + bd.oshift;
  const int xbA = (P.s * mx + bd.phx0 + px0 - P.p + bd.oshift) * 3 + co0;   // byte column of element 0
  if (oyA >= 0 && oyA < Hlim) {
    uint8_t* row = P.out_u8 + ((size_t)b * P.H + oyA) * (size_t)P.W * 3;
    tc_store_byte_run(row + xbA, w, max(0, -xbA), min(lenA, Wb - xbA));
  }
  if (lenA < nvalid) {                                             // next phase row: phase_x = 0, co = 0 at element lenA
    const int oyB = oyA + 1;
    const int xbB = (P.s * mx + bd.phx0 - P.p + bd.oshift) * 3 - lenA;   // byte column of (virtual) element 0
    if (oyB >= 0 && oyB < Hlim) {
      uint8_t* row = P.out_u8 + ((size_t)b * P.H + oyB) * (size_t)P.W * 3;
      tc_store_byte_run(row + xbB, w, max(lenA, -xbB), min(nvalid, Wb - xbB));
    }
  }
}

// Two-layer synthesis, layer 1: one output pixel = C1 base columns (|| C1 residual columns).
// t = act(base + bias) (+ res + bias'), act = IGDN1 / GDN1 / relu / leaky / none   (common/transforms.py:331-360)
template <int C1, bool RES>
__device__ __forceinline__ void tc_epi_two_layer_pixel(const TcParams& P,
                                                        const uint32_t* raw, float* dst, __half* phi, __half* plo, size_t kq_stride) {
  constexpr int PW = RES ? 2 * C1 : C1;
  float x[C1];
#pragma unroll
  for (int j = 0; j < C1; ++j) x[j] = fmaf(__uint_as_float(raw[j]), P.inv_scale, P.tl_bias[j]);
  float t[C1];
  if (P.tl_act == SNTC_ACT_IGDN1 || P.tl_act == SNTC_ACT_GDN1) {
#pragma unroll
    for (int j = 0; j < C1; ++j) t[j] = P.tl_beta[j];
#pragma unroll
    for (int i = 0; i < C1; ++i) {
      const float a = fabsf(x[i]);
#pragma unroll
      for (int j = 0; j < C1; ++j) t[j] = fmaf(a, P.tl_gamma[i * C1 + j], t[j]);
    }
#pragma unroll
    for (int j = 0; j < C1; ++j) t[j] = P.tl_inverse ? x[j] * t[j] : x[j] / t[j];
  } else {
#pragma unroll
    for (int j = 0; j < C1; ++j) t[j] = apply_act(x[j], P.tl_act);
  }
  if (RES) {
#pragma unroll
    for (int j = 0; j < C1; ++j) t[j] += fmaf(__uint_as_float(raw[C1 + j]), P.inv_scale, P.tl_bias[C1 + j]);
  }
  (void)PW;
  if (dst) {
#pragma unroll
    for (int j = 0; j < C1; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(t[j], t[j + 1], t[j + 2], t[j + 3]);
  }
  if (phi) {   // fp16 hi/lo planes, channels padded to a multiple of 16, one 16-byte octet per store (octet stride kq_stride)
    constexpr int CP = (C1 + 15) / 16 * 16;
#pragma unroll
    for (int g = 0; g < CP / 8; ++g) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c0 = 8 * g + 2 * i;
        __half h0 = __float2half(0.f), l0 = h0, h1 = h0, l1 = h0;
        if (c0 < C1) split_f16(t[c0 < C1 ? c0 : 0], h0, l0);
        if (c0 + 1 < C1) split_f16(t[c0 + 1 < C1 ? c0 + 1 : 0], h1, l1);
        h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      *reinterpret_cast<uint4*>(phi + (size_t)g * kq_stride) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(plo + (size_t)g * kq_stride) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

template <int C1, bool RES>
__device__ __forceinline__ void tc_epi_two_layer(const TcParams& P, const TcBandRegs& bd, const TcItem& it, uint32_t trow, int b, int my, int mx,
                                                 bool cell_ok, int pp0, int pstep) {
  constexpr int PW = RES ? 2 * C1 : C1;
  const int npx = it.nrows / PW;
  uint32_t raw[PW], nxt[PW];
  if (pp0 < npx) {
#pragma unroll
    for (int c = 0; c < PW; c += 4) tcx::tmem_ld4_nowait(trow + (uint32_t)(pp0 * PW + c), nxt + c);
  }
  for (int pp = pp0; pp < npx; pp += pstep) {
    tcx::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < PW; ++c) raw[c] = nxt[c];
    if (pp + pstep < npx) {   // the next pixel's accumulator columns are in flight while this one is processed
#pragma unroll
      for (int c = 0; c < PW; c += 4) tcx::tmem_ld4_nowait(trow + (uint32_t)((pp + pstep) * PW + c), nxt + c);
    }
    if (!cell_ok) continue;
    const int ph = (it.n0 + pp * PW) / PW;
    const int oy = P.s * my + bd.phy0 + ph / bd.nphx - P.p + bd.oshift, ox = P.s * mx + bd.phx0 + ph % bd.nphx - P.p + bd.oshift;
    if (oy < 0 || oy >= P.hout || ox < 0 || ox >= P.wout) continue;
    const size_t pix = ((size_t)b * P.hout + oy) * P.wout + ox;
    constexpr int CP = (C1 + 15) / 16 * 16;
    // planes for the tensor-core tail are octet-planar: [B][hout][CP/8][wout][8]
    const size_t poff = (((size_t)b * P.hout + oy) * (CP / 8) * P.wout + ox) * 8;
    tc_epi_two_layer_pixel<C1, RES>(P, raw, P.out_f32 ? P.out_f32 + pix * C1 : nullptr,
                                    P.out_hi ? P.out_hi + poff : nullptr, P.out_hi ? P.out_lo + poff : nullptr, (size_t)P.wout * 8);
  }
}

// Wide hidden layers (C1 = 48): the same per-pixel stage with the accumulator columns fetched in 16-column chunks -- base
// columns first, residual columns after the IGDN, so at most x[C1] + t[C1] + one chunk are live -- and gamma / beta in
// shared memory (float4 broadcast loads).  Measured on B200 (8 x 1200x1200): 1.21 ms, LSU-bound (a warp-wide LDS.128
// returns 512 B however few distinct addresses it has, i.e. one LSU cycle per FFMA); the constant-bank form of the narrow
// variants is slower here (1.90 ms): 9.2 KB of gamma thrash the constant cache.  Still 2.1x faster than the unfused
// conv + fp32 IGDN kernel it replaces (0.55 + 2.03 ms).
// tcgen05.ld is warp-collective: every lane runs the whole pixel, only the stores are predicated.
template <int C1, bool RES>
__device__ __forceinline__ void tc_epi_two_layer_wide(const TcParams& P, const TcBandRegs& bd, const TcItem& it, const float* sbias, const float* sgamma,
                                                      const float* sbeta, uint32_t trow, int b, int my, int mx, bool cell_ok, int pp0, int pstep) {
  static_assert(C1 % 16 == 0, "16-column TMEM chunks");
  constexpr int PW = RES ? 2 * C1 : C1;
  const int npx = it.nrows / PW;
  for (int pp = pp0; pp < npx; pp += pstep) {
    const uint32_t tcol = trow + (uint32_t)(pp * PW);
    float x[C1], t[C1];
#pragma unroll
    for (int c = 0; c < C1; c += 16) {
      uint32_t r[16];
      tcx::tmem_ld8_nowait(tcol + (uint32_t)c, r);
      tcx::tmem_ld8_nowait(tcol + (uint32_t)c + 8, r + 8);
      tcx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) x[c + i] = fmaf(__uint_as_float(r[i]), P.inv_scale, sbias[c + i]);
    }
    if (P.tl_act == SNTC_ACT_IGDN1 || P.tl_act == SNTC_ACT_GDN1) {
#pragma unroll
      for (int j = 0; j < C1; j += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(sbeta + j);
        t[j] = bb.x; t[j + 1] = bb.y; t[j + 2] = bb.z; t[j + 3] = bb.w;
      }
#pragma unroll
      for (int i = 0; i < C1; ++i) {
        const float a = fabsf(x[i]);
#pragma unroll
        for (int j = 0; j < C1; j += 4) {
          const float4 g = *reinterpret_cast<const float4*>(sgamma + i * C1 + j);
          t[j] = fmaf(a, g.x, t[j]); t[j + 1] = fmaf(a, g.y, t[j + 1]); t[j + 2] = fmaf(a, g.z, t[j + 2]); t[j + 3] = fmaf(a, g.w, t[j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < C1; ++j) t[j] = P.tl_inverse ? x[j] * t[j] : x[j] / t[j];
    } else {
#pragma unroll
      for (int j = 0; j < C1; ++j) t[j] = apply_act(x[j], P.tl_act);
    }
    if (RES) {
#pragma unroll
      for (int c = 0; c < C1; c += 16) {
        uint32_t r[16];
        tcx::tmem_ld8_nowait(tcol + (uint32_t)(C1 + c), r);
        tcx::tmem_ld8_nowait(tcol + (uint32_t)(C1 + c) + 8, r + 8);
        tcx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) t[c + i] += fmaf(__uint_as_float(r[i]), P.inv_scale, sbias[C1 + c + i]);
      }
    }
    const int ph = (it.n0 + pp * PW) / PW;
    const int oy = P.s * my + bd.phy0 + ph / bd.nphx - P.p + bd.oshift, ox = P.s * mx + bd.phx0 + ph % bd.nphx - P.p + bd.oshift;
    if (!cell_ok || oy < 0 || oy >= P.hout || ox < 0 || ox >= P.wout) continue;
    const size_t pix = ((size_t)b * P.hout + oy) * P.wout + ox;
    if (P.out_f32) {
      float* dst = P.out_f32 + pix * C1;
#pragma unroll
      for (int j = 0; j < C1; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(t[j], t[j + 1], t[j + 2], t[j + 3]);
    }
    if (P.out_hi) {   // octet-planar fp16 hi/lo planes for the tensor-core tail: [B][hout][C1/8][wout][8]
      const size_t poff = (((size_t)b * P.hout + oy) * (C1 / 8) * P.wout + ox) * 8;
      const size_t kq_stride = (size_t)P.wout * 8;
#pragma unroll
      for (int g = 0; g < C1 / 8; ++g) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __half h0, l0, h1, l1;
          split_f16(t[8 * g + 2 * i], h0, l0); split_f16(t[8 * g + 2 * i + 1], h1, l1);
          h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint4*>(P.out_hi + poff + (size_t)g * kq_stride) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(P.out_lo + poff + (size_t)g * kq_stride) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

// Gather phase of the col2im epilogue for compile-time (K, S, PD).  ROWS = false: one work item = (interior cell, channel) -> its S x S
// output samples; every tap of the 3 x 3 neighbour cells is read exactly once and every validity test is a compile-time constant
// (k5 s2: 25 loads + 25 adds -> 4 samples; 84 cells x 3 channels = 252 items on the 256 epilogue threads).  ROWS = true: one work
// item = (cell, channel, output row phase fy) -> S samples (k9 s4, one channel per n-tile: 336 items instead of 84).
template <int K, int S, int PD, bool ROWS>
__device__ __forceinline__ void tc_c2i_gather(const TcParams& P, const TcItem& it, const float* sbias, const float* sP, int PS, int et) {
  constexpr int KP = (K * K + 31) / 32 * 32;
  constexpr int NR = ROWS ? S : 1;
  if (it.dup) return;
  const int G = it.nrows / KP, g0 = it.n0 / KP;
  constexpr unsigned iw = 14;                    // col2im tiles are 8 x 16 cells (tc_run_conv): 6 x 14 interior
  const int nwork = (P.TH - 2) * (int)iw * G * NR;
  const size_t row_u8 = (size_t)P.W * P.cout, row_f32 = (size_t)P.wout * P.cout;
  for (int wi = et; wi < nwork; wi += 32 * TC_EPI_WARPS) {
    unsigned q = (unsigned)wi;
    const int fy0 = ROWS ? (int)(q % S) : 0; if (ROWS) q /= S;
    int g;
    if (G == 3) { g = (int)(q % 3u); q /= 3u; } else if (G == 1) { g = 0; } else { g = (int)(q % (unsigned)G); q /= (unsigned)G; }
    const int cx = 1 + (int)(q % iw), cy = 1 + (int)(q / iw);
    const int my = it.iy0 + cy, mx = it.ix0 + cx;
    if (my >= P.hin || mx >= P.win) continue;
    const float* cell = sP + (size_t)(cy * P.TW + cx) * PS + g * KP;
    float acc[ROWS ? 1 : S][S];
#pragma unroll
    for (int a = 0; a < (ROWS ? 1 : S); ++a)
#pragma unroll
      for (int b = 0; b < S; ++b) acc[a][b] = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const float* src = cell + (ptrdiff_t)(dy * P.TW + dx) * PS;
#pragma unroll
        for (int a = 0; a < (ROWS ? 1 : S); ++a) {
          const int ay = (ROWS ? fy0 : a) + PD - S * dy;          // compile-time unless ROWS
          if (ay < 0 || ay >= K) continue;
#pragma unroll
          for (int fx = 0; fx < S; ++fx) {
            const int ax = fx + PD - S * dx;                        // compile-time
            if (ax >= 0 && ax < K) acc[a][fx] += src[ay * K + ax];
          }
        }
      }
    }
    const int co = g0 + g, oy0 = S * my + fy0, ox0 = S * mx;
    const float bias = sbias[co];
    float* of = P.out_f32 ? P.out_f32 + (((size_t)it.b * P.hout + oy0) * P.wout + ox0) * P.cout + co : nullptr;
    const size_t qi = (((size_t)it.b * P.H + oy0) * P.W + ox0) * P.cout + co;
#pragma unroll
    for (int a = 0; a < (ROWS ? 1 : S); ++a) {
#pragma unroll
      for (int fx = 0; fx < S; ++fx) {
        const float x = fmaf(acc[a][fx], P.inv_scale, bias);
        if (of) of[a * row_f32 + (size_t)fx * P.cout] = x;
        if ((P.out_u8 || P.out_crop) && oy0 + a < P.H && ox0 + fx < P.W) {
          const size_t o = qi + a * row_u8 + (size_t)fx * P.cout;
          if (P.out_u8) P.out_u8[o] = float_to_pixel(x);
          if (P.out_crop) P.out_crop[o] = x;
        }
      }
    }
  }
}

// any (k, s, p) with a one-cell halo: one work item per output sample
__device__ __forceinline__ void tc_c2i_gather_any(const TcParams& P, const TcItem& it, const float* sbias, const float* sP, int PS, int et) {
  const int k = P.c2i_k, s = P.c2i_s, p = P.c2i_p, KP = P.c2i_kp;
  const int G = it.nrows / KP, g0 = it.n0 / KP;
  const int IW = (P.TW - 2) * s, nout = (P.TH - 2) * s * IW * G;
  if (!it.dup) {
    for (int e = et; e < nout; e += 32 * TC_EPI_WARPS) {
      const int g = e % G, q = e / G;
      const int oxl = q % IW, oyl = q / IW;
      const int cy = 1 + oyl / s, cx = 1 + oxl / s, fy = oyl % s, fx = oxl % s;
      const int my = it.iy0 + cy, mx = it.ix0 + cx;
      if (my >= P.hin || mx >= P.win) continue;
      float acc = 0.f;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int ay = fy + p - s * dy;
        if (ay < 0 || ay >= k) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int ax = fx + p - s * dx;
          if (ax < 0 || ax >= k) continue;
          acc += sP[(size_t)((cy + dy) * P.TW + (cx + dx)) * PS + g * KP + ay * k + ax];
        }
      }
      const int co = g0 + g, oy = s * my + fy, ox = s * mx + fx;
      const float x = fmaf(acc, P.inv_scale, sbias[co]);
      if (P.out_f32) P.out_f32[(((size_t)it.b * P.hout + oy) * P.wout + ox) * P.cout + co] = x;
      if ((P.out_u8 || P.out_crop) && oy < P.H && ox < P.W) {
        const size_t qi = (((size_t)it.b * P.H + oy) * P.W + ox) * P.cout + co;
        if (P.out_u8) P.out_u8[qi] = float_to_pixel(x);
        if (P.out_crop) P.out_crop[qi] = x;
      }
    }
  }
}

// ---- col2im epilogue (TC_EPI_COL2IM): overlap-add of the per-input-pixel products of a final layer ----
// The accumulator row of tile cell (cy, cx) holds P[n, (g, a_y, a_x)] for the channels g of this n-tile.  The 8 epilogue warps
// stage the tile in shared memory (row stride odd: conflict-free), then every thread sums, for its output samples of the
// INTERIOR cells (1 .. TH-2, 1 .. TW-2), the <= 3 x 3 neighbour terms in a fixed order:
//     out[s m + f, g] = bias[g] + sum_{d in {-1,0,1}^2, a = f + p - s d in [0,k)^2} P[m + d, (g, a)]
// (cells outside the image were zero-filled by TMA, so their P is 0 = no contribution).  Which tile a cell belongs to never
// changes the sum: bit-identical across batch sizes and band splits.
__device__ __forceinline__ void tc_epi_col2im(const TcParams& P, const TcItem& it, const float* sbias, float* sP, uint32_t trow, int r, int eh, int EH, int et) {
  const int PS = P.bn_max + 1;
  {
    uint32_t raw[32];
    for (int c = 32 * eh; c < it.mma_n; c += 32 * EH) {
      tcx::tmem_ld32_nowait(trow + (uint32_t)c, raw);
      tcx::tmem_ld_wait();
      float* dst = sP + (size_t)r * PS + c;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c + i < it.mma_n) dst[i] = __uint_as_float(raw[i]);
    }
  }
  asm volatile("bar.sync 1, %0;" ::"r"(32 * TC_EPI_WARPS) : "memory");
  if (P.c2i_k == 5 && P.c2i_s == 2 && P.c2i_p == 2) tc_c2i_gather<5, 2, 2, false>(P, it, sbias, sP, PS, et);          // tfc SignalConv2D 5x5 up 2 (mbt2018)
  else if (P.c2i_k == 5 && P.c2i_s == 2 && P.c2i_p == 1) tc_c2i_gather<5, 2, 1, false>(P, it, sbias, sP, PS, et);     // Keras ConvT 5x5 s2 (CNNSynthesis)
  else if (P.c2i_k == 9 && P.c2i_s == 4 && P.c2i_p == 4) tc_c2i_gather<9, 4, 4, true>(P, it, sbias, sP, PS, et);     // tfc 9x9 up 4 (bls2017)
  else tc_c2i_gather_any(P, it, sbias, sP, PS, et);
  asm volatile("bar.sync 1, %0;" ::"r"(32 * TC_EPI_WARPS) : "memory");   // the staging tile is free for the next item
}

// ------------------------------------------------------------------------------------------------
// Persistent layer kernel: every CTA (CG = 1) or CTA pair (CG = 2, one cluster = two SMs of a TPC) loops over
// the layer's work items (band, m-tile group, n-tile); the smem ring and the two TMEM accumulators run across
// items, so the epilogue of item i overlaps the MMAs of item i+1.
//
// CG = 2 (cta_group::2): the pair computes D[2 x 128 cells, N].  Each CTA TMA-loads its own 128-cell A tile and
// HALF of the W tile rows into its own smem; the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads
// A and the W halves from both SMs -> per-SM shared-memory operand traffic drops by the W half.  Transaction bytes
// of both CTAs are counted on the leader's full barrier; the MMA commit is multicast to both CTAs' barriers.
template <int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
band_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS,
  // not generic loads) for everything derived from `smem`
  uint8_t* smem = smem_raw + ((1024u - (tcx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_bytes = TC_BM * 128;                        // one A plane tile
  const uint32_t b_slot = (uint32_t)(P.bn_max / CG) * 128;     // smem reserved per W plane tile (this CTA's rows)
  // The ring region is P.stages slots of the FULL stage (A hi | A lo | W hi | W lo); when a plane is not loaded (integer-valued input:
  // no A lo; 2-pass mode: no W lo) the same bytes hold more, smaller stages (see `nst` below).  Measured: bls2017 4K layer_0 (integer
  // symbols) 0.290 -> 0.272 ms; the single-image hyper layer 0 (5 -> 8 stages) does not move (27 us): its pace is not set by the ring.
  const uint32_t ring_bytes = (uint32_t)P.stages * (2 * a_bytes + 2 * b_slot);
  constexpr int TC_MAX_STAGES = 16;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + ring_bytes);
  uint64_t* empty_bar = full_bar + TC_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + TC_MAX_STAGES;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* sconst = reinterpret_cast<float*>(tmem_slot + 4);   // epilogue constants: bias [cout]

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops
  // (and the operands of the TMA / MMA instructions) on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int rank = CG == 2 ? (int)tcx::cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int unit0 = (int)blockIdx.x / CG, nunits = (int)gridDim.x / CG;   // persistent work units (CTAs or pairs)

  tcx::pdl_launch_dependents();
  if (warp == 0 && lane == 0) { tcx::prefetch_tmap(&mapAhi); tcx::prefetch_tmap(&mapAlo); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < TC_MAX_STAGES; ++i) { tcx::mbar_init(&full_bar[i], 1); tcx::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { tcx::mbar_init(&tmem_full_bar[i], 1); tcx::mbar_init(&tmem_empty_bar[i], TC_EPI_WARPS * CG); }
    tcx::fence_barrier_init();
  }
  if (warp == 2) { if (CG == 2) tcx::tmem_alloc_2sm(tmem_slot, 2 * TC_ACC_COLS); else tcx::tmem_alloc(tmem_slot, 2 * TC_ACC_COLS); }
  if (warp >= 4) {   // epilogue constants -> shared memory: bias [cout]
    const int t = threadIdx.x - 128, nb = P.cout, nt = 32 * TC_EPI_WARPS;
    for (int i = t; i < nb; i += nt) sconst[i] = P.bias[i];
    if (P.epi == TC_EPI_TWO_LAYER && P.C1 > 24 && P.gamma) {   // wide two-layer epilogue: gamma [C1][C1] | beta [C1] after the bias
      float* sg = sconst + ((nb + 3) & ~3);
      for (int i = t; i < P.C1 * P.C1; i += nt) sg[i] = P.gamma[(i / P.C1) * P.gamma_stride + (i % P.C1)];
      for (int i = t; i < P.C1; i += nt) sg[P.C1 * P.C1 + i] = P.beta[i];
    }
  }
  tcx::tc_fence_before();
  if (CG == 2) tcx::cluster_sync_all(); else __syncthreads();
  tcx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = tcx::smem_u32(smem);
  tcx::pdl_wait();   // everything above only touched weights / our own smem; from here on we read the previous layer's output
  unsigned pmask = P.pass_mask;
  if (P.alo_flag != nullptr && *reinterpret_cast<const volatile unsigned*>(P.alo_flag) == 0u) pmask &= ~1u;
  const bool use_alo = (pmask & 1u) != 0, use_blo = (pmask & 2u) != 0;
  const uint32_t off_bhi = (use_alo ? 2u : 1u) * a_bytes, off_blo = off_bhi + b_slot;
  const uint32_t stage_bytes = off_bhi + (use_blo ? 2u : 1u) * b_slot;          // a multiple of 1024 (a_bytes = 16 KB, b_slot = k * 2 KB)
  const uint32_t nst = min((uint32_t)TC_MAX_STAGES, ring_bytes / stage_bytes);   // >= P.stages

  if (warp == 0) {
    // ===== TMA producer: the whole warp runs the (uniform) loop, one elected lane issues the copies.
    // Every CTA loads its own A tile and its own rows of W. =====
    uint32_t st = 0, ph = 0;   // smem ring position: runs across work items
    for (int kw = 0;; ++kw) {
      const int item = tc_unit_item(unit0, nunits, kw, P.snake);
      if (item >= P.total_items) break;
      const TcItem it = tc_decode_item<CG>(P, item, rank);
      const TcBandDev& bd = P.bands[it.band];
      const int Ty = bd.Ty, Tx = bd.Tx;
      const int wrows = bd.BN / CG;                                  // W box rows per CTA
      const int wrow0 = it.n0 + rank * (it.mma_n / CG);              // this CTA supplies columns [rank*mma_n/CG, ...)
      const uint32_t tx_bytes = (uint32_t)CG * ((use_alo ? 2u : 1u) * a_bytes + (use_blo ? 2u : 1u) * (uint32_t)wrows * 128);
      const int cx0 = bd.mlox + it.ix0, cy0 = bd.mloy + it.iy0;
      int kcol = 0;
      const int jt = kw;
      long long* tr = (P.trace && leader && jt < TC_TRACE_ITEMS) ? P.trace + ((size_t)unit0 * TC_TRACE_ITEMS + jt) * 8 : nullptr;
      if (tr && lane == 0) { tr[0] = clock64(); tr[7] = item; }
      for (int jy = 0; jy < Ty; ++jy)
        for (int jx = 0; jx < Tx; ++jx)
          for (int kb = 0; kb < P.kblocks; ++kb, kcol += TC_BK) {
            tcx::mbar_wait(&empty_bar[st], ph ^ 1u);
            if (tcx::elect_one()) {
              uint8_t* sa = smem + (size_t)st * stage_bytes;
              const int cx = cx0 - jx, cy = cy0 - jy, cc = kb * TC_BK;
              if (CG == 2) {
                if (leader) tcx::mbar_expect_tx(&full_bar[st], tx_bytes);
                const uint32_t fb = tcx::mapa_u32(tcx::smem_u32(&full_bar[st]), 0);   // the leader's full barrier
                tcx::tma_load_4d_2sm(sa, &mapAhi, fb, cc, cx, cy, it.b);
                if (use_alo) tcx::tma_load_4d_2sm(sa + a_bytes, &mapAlo, fb, cc, cx, cy, it.b);
                tcx::tma_load_2d_2sm(sa + off_bhi, &bd.mapBhi, fb, kcol, wrow0);
                if (use_blo) tcx::tma_load_2d_2sm(sa + off_blo, &bd.mapBlo, fb, kcol, wrow0);
              } else {
                tcx::mbar_expect_tx(&full_bar[st], tx_bytes);
                tcx::tma_load_4d(sa, &mapAhi, &full_bar[st], cc, cx, cy, it.b);
                if (use_alo) tcx::tma_load_4d(sa + a_bytes, &mapAlo, &full_bar[st], cc, cx, cy, it.b);
                tcx::tma_load_2d(sa + off_bhi, &bd.mapBhi, &full_bar[st], kcol, wrow0);
                if (use_blo) tcx::tma_load_2d(sa + off_blo, &bd.mapBlo, &full_bar[st], kcol, wrow0);
              }
            }
            __syncwarp();
            if (++st == nst) { st = 0; ph ^= 1u; }
          }
      if (tr && lane == 0) tr[1] = clock64();
    }
  } else if (warp == 1) {
    if (leader) {
      // ===== MMA issuer (leader CTA only): warp-uniform loop, one elected lane issues the MMAs + commits =====
      // descriptor = fixed high word | (smem address >> 4) in the low word; +2 per K=16 step inside the swizzle row
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      uint32_t st = 0, ph = 0, j = 0;
      for (;; ++j) {
        const int item = tc_unit_item(unit0, nunits, (int)j, P.snake);
        if (item >= P.total_items) break;
        const TcItem it = tc_decode_item<CG>(P, item, 0);
        const TcBandDev& bd = P.bands[it.band];
        const int ntaps = bd.Ty * bd.Tx;
        const uint32_t buf = j & 1u;
        long long* tr = (P.trace && j < TC_TRACE_ITEMS) ? P.trace + ((size_t)unit0 * TC_TRACE_ITEMS + j) * 8 : nullptr;
        tcx::mbar_wait(&tmem_empty_bar[buf], ((j >> 1) & 1u) ^ 1u);   // the epilogue(s) drained this accumulator
        tcx::tc_fence_after();
        if (tr && lane == 0) tr[2] = clock64();
        const uint32_t tacc = tmem_base + buf * TC_ACC_COLS;
        const uint32_t idesc = tcx::make_idesc(TC_BM * CG, it.mma_n);
        uint32_t acc = 0;
        for (int tap = 0; tap < ntaps; ++tap)
          for (int kb = 0; kb < P.kblocks; ++kb) {
            tcx::mbar_wait(&full_bar[st], ph);
            tcx::tc_fence_after();
            if (tr && lane == 0 && tap == 0 && kb == 0) tr[3] = clock64();
            const uint32_t sa = smem_base + st * stage_bytes;
            const uint32_t a_hi = (((sa) & 0x3FFFFu) >> 4) | (1u << 16), a_lo = (((sa + a_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_hi = (((sa + off_bhi) & 0x3FFFFu) >> 4) | (1u << 16), b_lo = (((sa + off_blo) & 0x3FFFFu) >> 4) | (1u << 16);
            const int nm = (kb == P.kblocks - 1) ? P.last_kmma : 4;
            if (tcx::elect_one()) {
              // lo*hi and hi*lo first (small terms), hi*hi last
              if (nm == 4 && pmask == 3u) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { tcx::umma_f16_t<CG>(tacc, tcx::desc64(a_lo + 2 * q, desc_hi), tcx::desc64(b_hi + 2 * q, desc_hi), idesc, q == 0 ? acc : 1u); }
#pragma unroll
                for (int q = 0; q < 4; ++q) tcx::umma_f16_t<CG>(tacc, tcx::desc64(a_hi + 2 * q, desc_hi), tcx::desc64(b_lo + 2 * q, desc_hi), idesc, 1u);
#pragma unroll
                for (int q = 0; q < 4; ++q) tcx::umma_f16_t<CG>(tacc, tcx::desc64(a_hi + 2 * q, desc_hi), tcx::desc64(b_hi + 2 * q, desc_hi), idesc, 1u);
              } else {
                uint32_t a1 = acc;   // accumulate flag of the next MMA: 0 only for the very first MMA of the work item
                if (use_alo) for (int q = 0; q < nm; ++q) { tcx::umma_f16_t<CG>(tacc, tcx::desc64(a_lo + 2 * q, desc_hi), tcx::desc64(b_hi + 2 * q, desc_hi), idesc, a1); a1 = 1u; }
                if (use_blo) for (int q = 0; q < nm; ++q) { tcx::umma_f16_t<CG>(tacc, tcx::desc64(a_hi + 2 * q, desc_hi), tcx::desc64(b_lo + 2 * q, desc_hi), idesc, a1); a1 = 1u; }
                for (int q = 0; q < nm; ++q) { tcx::umma_f16_t<CG>(tacc, tcx::desc64(a_hi + 2 * q, desc_hi), tcx::desc64(b_hi + 2 * q, desc_hi), idesc, a1); a1 = 1u; }
              }
              if (CG == 2) tcx::umma_commit_2sm(&empty_bar[st], 3);   // frees this smem stage in both CTAs
              else tcx::umma_commit(&empty_bar[st]);
            }
            __syncwarp();
            acc = 1;
            if (++st == nst) { st = 0; ph ^= 1u; }
          }
        if (tcx::elect_one()) {   // accumulator complete
          if (CG == 2) tcx::umma_commit_2sm(&tmem_full_bar[buf], 3); else tcx::umma_commit(&tmem_full_bar[buf]);
        }
        __syncwarp();
        if (tr && lane == 0) tr[4] = clock64();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global (every CTA drains its own 128 accumulator rows) =====
    const int ew = warp & 3;                    // TMEM lanes [32*ew, 32*ew+32): a warp may only touch the quarter (warp % 4)
    const int eh = (warp - 4) >> 2;             // which of the quarter's warps: they take alternate column chunks / pixels
    constexpr int EH = TC_EPI_WARPS / 4;
    const int r = ew * 32 + lane;               // row of the tile = cell
    const float* sbias = sconst;
    const int r_x = r % P.TW, r_q = r / P.TW, r_y = r_q % P.TH, r_b = r_q / P.TH;   // tile row -> (image, y, x) of the TMA box
    uint32_t j = 0;
    for (;; ++j) {
      const int item = tc_unit_item(unit0, nunits, (int)j, P.snake);
      if (item >= P.total_items) break;
      TcItem it = tc_decode_item<CG>(P, item, rank);
      it.b += r_b;                                   // this thread's image (tiles of TB > 1 images: rows [b][y][x])
      const TcBandDev& bdg = P.bands[it.band];
      const TcBandRegs bd{bdg.N, bdg.nphx, bdg.phy0, bdg.phx0, bdg.oshift};
      const int nk = bdg.Ty * bdg.Tx * P.kblocks;
      const uint32_t buf = j & 1u;
      const int iy = it.iy0 + r_y, ix = it.ix0 + r_x;
      const bool cell_ok = iy < P.hin && ix < P.win && it.b < P.B && !it.dup;
      const int my = bdg.mloy + iy, mx = bdg.mlox + ix;
      long long* tr = (P.trace && leader && warp == 4 && lane == 0 && j < TC_TRACE_ITEMS) ? P.trace + ((size_t)unit0 * TC_TRACE_ITEMS + j) * 8 : nullptr;
      tcx::mbar_wait(&tmem_full_bar[buf], (j >> 1) & 1u);
      tcx::tc_fence_after();
      if (tr) tr[5] = clock64();
      const uint32_t trow = tmem_base + buf * TC_ACC_COLS + ((uint32_t)(ew * 32) << 16);
      if (P.epi == TC_EPI_COL2IM) {
        tc_epi_col2im(P, it, sbias, sconst + ((P.cout + 3) & ~3), trow, r, eh, EH, (int)threadIdx.x - 128);
      } else if (P.epi == TC_EPI_TWO_LAYER && nk > 0) {
        if (P.C1 == 48) {
          const float* sg = sconst + ((P.cout + 3) & ~3);
          if (P.has_res) tc_epi_two_layer_wide<48, true>(P, bd, it, sbias, sg, sg + 48 * 48, trow, it.b, my, mx, cell_ok, eh, EH);
          else tc_epi_two_layer_wide<48, false>(P, bd, it, sbias, sg, sg + 48 * 48, trow, it.b, my, mx, cell_ok, eh, EH);
        } else
        if (P.C1 == 12) { if (P.has_res) tc_epi_two_layer<12, true>(P, bd, it, trow, it.b, my, mx, cell_ok, eh, EH);
                          else tc_epi_two_layer<12, false>(P, bd, it, trow, it.b, my, mx, cell_ok, eh, EH); }
        else            { if (P.has_res) tc_epi_two_layer<24, true>(P, bd, it, trow, it.b, my, mx, cell_ok, eh, EH);
                          else tc_epi_two_layer<24, false>(P, bd, it, trow, it.b, my, mx, cell_ok, eh, EH); }
      } else {
        if (P.vec16 && nk > 0 && P.gdn_mode != G_NONE && P.act == SNTC_ACT_NONE && P.plane_xform == A_NONE) {
          switch (P.gdn_mode) {   // mode-specialised GDN epilogue
            case G_MUL: tc_epi_gdn_item<G_MUL>(P, bd, it, sbias, trow, my, mx, cell_ok, eh, EH); break;
            case G_DIV: tc_epi_gdn_item<G_DIV>(P, bd, it, sbias, trow, my, mx, cell_ok, eh, EH); break;
            case G_MUL_SQRT: tc_epi_gdn_item<G_MUL_SQRT>(P, bd, it, sbias, trow, my, mx, cell_ok, eh, EH); break;
            default: tc_epi_gdn_item<G_DIV_SQRT>(P, bd, it, sbias, trow, my, mx, cell_ok, eh, EH); break;
          }
        } else if (P.vec16 && nk > 0) {
          // 32-column chunks (one tcgen05.ld.x32 each), the quarter's warps take alternate chunks; the load of the
          // next chunk is in flight while this one is processed
          int co = it.n0 % P.cout, ph = it.n0 / P.cout, c_at = 0;
          int last_ph = -1, oy = 0, ox = 0;
          uint32_t raw[32], nxt[32];
          float bits = 0.f;
          int c = 32 * eh;
          if (c < it.mma_n) tcx::tmem_ld32_nowait(trow + (uint32_t)c, nxt);
          for (; c < it.mma_n; c += 32 * EH) {
            tcx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = nxt[i];
            if (c + 32 * EH < it.mma_n) tcx::tmem_ld32_nowait(trow + (uint32_t)(c + 32 * EH), nxt);
            if (!cell_ok) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int cc = c + 16 * h;
              if (it.n0 + cc >= bd.N) break;
              co += cc - c_at; c_at = cc;
              while (co >= P.cout) { co -= P.cout; ++ph; }
              if (ph != last_ph) {
                last_ph = ph;
                oy = P.s * my + bd.phy0 + ph / bd.nphx - P.p + bd.oshift; ox = P.s * mx + bd.phx0 + ph % bd.nphx - P.p + bd.oshift;
              }
              bits += tc_epi_vec16(P, sbias, it.b, oy, ox, co, raw + 16 * h);
            }
          }
          if (P.rate_slots) {   // one deterministic partial per (item, CTA, epilogue warp)
            double d = (double)bits;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) d += __shfl_down_sync(0xffffffffu, d, off);
            if (lane == 0) {
              const size_t slot = ((size_t)item * CG + rank) * TC_EPI_WARPS + (warp - 4);
              P.rate_slots[slot] = d;
              P.rate_slot_img[slot] = (it.dup || it.b >= P.B) ? -1 : it.b;
            }
          }
        } else if (P.rgb_runs && bd.nphx * 3 >= 32 && nk > 0) {
          // 3-channel final layer, uint8 image only: 32-column chunks -> byte runs (see tc_epi_rgb_chunk)
          uint32_t raw[32], nxt[32];
          int c = 32 * eh;
          if (c < it.mma_n) tcx::tmem_ld32_nowait(trow + (uint32_t)c, nxt);
          for (; c < it.mma_n; c += 32 * EH) {
            tcx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = nxt[i];
            if (c + 32 * EH < it.mma_n) tcx::tmem_ld32_nowait(trow + (uint32_t)(c + 32 * EH), nxt);
            if (cell_ok) tc_epi_rgb_chunk(P, bd, sbias, it.b, my, mx, it.n0 + c, it.n0 + it.nrows, raw);
          }
        } else {
        const bool vec = (P.cout % 8) == 0;
        // (co, phase) of column n = it.n0 + c advance incrementally: no per-group division by cout
        int co = it.n0 % P.cout, ph = it.n0 / P.cout, c_at = 0;
        int last_ph = -1, oy = 0, ox = 0;
        for (int c = 16 * eh; c < it.mma_n; c += 16 * EH) {
          uint32_t raw[16];
          if (nk > 0) {                           // warp-wide TMEM loads: executed by all lanes
            tcx::tmem_ld8_nowait(trow + (uint32_t)c, raw);
            tcx::tmem_ld8_nowait(trow + (uint32_t)c + 8, raw + 8);
            tcx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) raw[i] = 0u;
          }
          if (!cell_ok) continue;
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const int n = it.n0 + c + h8 * 8;
            if (n >= bd.N) break;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[h8 * 8 + i]);
            if (vec) {
              co += (c + h8 * 8) - c_at; c_at = c + h8 * 8;
              while (co >= P.cout) { co -= P.cout; ++ph; }
              if (ph != last_ph) {
                last_ph = ph;
                oy = P.s * my + bd.phy0 + ph / bd.nphx - P.p + bd.oshift; ox = P.s * mx + bd.phx0 + ph % bd.nphx - P.p + bd.oshift;
              }
              tc_epi_vec8(P, sbias, it.b, oy, ox, co, v);
            } else {
              tc_epi_scalar8(P, bd, sbias, it.b, my, mx, n, v);
            }
          }
        }
        }
      }
      // this warp is done reading accumulator `buf`: tell the (leader's) MMA issuer
      tcx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) tcx::mbar_arrive_cluster(tcx::mapa_u32(tcx::smem_u32(&tmem_empty_bar[buf]), 0));
        else tcx::mbar_arrive(&tmem_empty_bar[buf]);
      }
      if (tr) tr[6] = clock64();
    }
  }
  tcx::tc_fence_before();
  if (CG == 2) tcx::cluster_sync_all(); else __syncthreads();   // the peer may still read this CTA's smem / signal its barriers
  if (warp == 2) { if (CG == 2) tcx::tmem_dealloc_2sm(tmem_base, 2 * TC_ACC_COLS); else tcx::tmem_dealloc(tmem_base, 2 * TC_ACC_COLS); }
}

// f32 NHWC -> fp16 hi/lo planes (input of the first tensor-core layer).  lo_flag (nullable, zeroed by the caller) is set
// to 1 when any lo element is non-zero: an all-zero lo plane (integer-valued symbols) lets the layer skip its a_lo * w_hi pass.
__global__ void split_planes_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, size_t n8, unsigned* lo_flag,
                                    int xform = A_NONE) {   // xform: the planes receive |x| / x^2 (pooling input of a GDN stage)
  tcx::pdl_launch_dependents();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool nz = false;
  if (i < n8) {
    float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { split_f16(a_xform(v[k], xform), h[k], l[k]); nz = nz || (__half_as_ushort(l[k]) & 0x7FFFu) != 0; }
    *reinterpret_cast<uint4*>(hi + i * 8) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + i * 8) = *reinterpret_cast<const uint4*>(l);
  }
  if (lo_flag != nullptr && __any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(lo_flag, 1u);
}

// ------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcDriver {
  PFN_encodeTiled encode = nullptr;
  int num_sms = 148;
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
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) num_sms = n;
  }
};

struct TcConv {
  bool ok = false;
  int cin = 0, kblocks = 0, last_kmma = 4, ktot_pad = 0;   // per-tap K padded to kblocks*64
  int bn_max = 16, stages = 2;
  int cg = 1;                                              // CTAs per work item: 2 = cta_group::2 pairs
  bool fused_two_layer = false;                            // layer-1 of a two-layer synthesis with the IGDN(+res) epilogue
  float scale = 1.f;
  __half* d_hi = nullptr; __half* d_lo = nullptr;          // [rows][kmax] K-major
  size_t kmax = 0;                                         // row pitch (elements)
  std::vector<TcBandDev> bands;                            // host copy, sorted by K descending
  std::vector<int> items_per_mtile;                        // n-tiles per band
  TcBandDev* d_bands = nullptr;
  int nbands = 0;
  // Second tiling of the same packed weights with narrow n-tiles (<= 64 columns), for batches whose wide tiling has
  // fewer work items than half the persistent units (a single 768x512 image gives the 3x3 hyper head 18 items for 74
  // CTA pairs): more, shorter items -> the layer's latency drops with the serial K loop of one item.
  struct Tiling { std::vector<TcBandDev> bands; TcBandDev* d_bands = nullptr; int nbands = 0, bn_max = 16, stages = 2; } narrow;
};

inline int tc_env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// Work items of the wide tiling for `mtiles` m-tiles; the narrow tiling is used when they leave half the units idle.
inline bool tc_use_narrow(const TcConv& t, int mtiles, int num_sms) {
  const int env = tc_env_int("SNTC_TC_NARROW", -1);   // 0 / 1 force, default auto
  if (t.narrow.nbands == 0 || env == 0) return false;
  if (env == 1) return true;
  const int groups = (mtiles + t.cg - 1) / t.cg;
  long items = 0;
  for (auto& bd : t.bands) items += (long)groups * bd.ntiles;
  // SNTC_TC_NARROW_WAVES (x 0.01): narrow when the wide tiling has fewer items than this many waves of the persistent units
  static const int waves_pct = tc_env_int("SNTC_TC_NARROW_WAVES", 50);
  return items * 100 <= (long)waves_pct * (num_sms / t.cg);
}

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

// tensor-core tail of a two-layer synthesis (kernel + packing in sntc_kernels_tail_tc.cuh)
struct TailTc {
  bool ok = false;
  int CP = 0, KQ = 0, stages = 0;
  float scale = 1.f;
  __half* d_w = nullptr;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  bool attr_set = false;
};

// GDN norm pool of a deep decoder as a 1x1 band GEMM (gamma [in,out] as a [1,1,Cin,Cout] kernel, beta as the bias)
struct TcGdn {
  bool ok = false;
  ConvLayer conv;
  TcConv tc;
};

struct TcModelState {
  std::vector<TcConv> hyper, syn;     // parallel to Transform::convs
  std::vector<TcGdn> syn_gdn;         // parallel to syn->gdns
  std::vector<TailTc> syn_tail;       // parallel to syn: tensor-core tail of a two-layer synthesis (sntc_kernels_tail_tc.cuh)
  TcDevBuf plane[4];                  // hi/lo ping-pong activation planes
  TcDevBuf yh[2];                     // hi/lo planes of y_hat (output of the fused hyper-synthesis head)
  bool smem_attr_set = false;
  void release() { for (auto& b : plane) b.release(); yh[0].release(); yh[1].release(); }
};

// Widest n-tile (multiple of `unit`, itself a multiple of 16) that keeps >= 3 pipeline stages in 227 KB
// and wastes the least padded columns; the last tile of a band only issues MMAs for its own columns.
inline int tc_bn_max() {
  static int v = -1;
  if (v < 0) { v = tc_env_int("SNTC_TC_BN_MAX", 256); if (v < 16 || v > TC_ACC_COLS) v = 256; }
  return v;
}
// Programmatic dependent launch between the kernels of one decode: the next kernel's prologue (barrier init, TMEM allocation,
// tensor-map prefetch, weight-fragment loads) overlaps the tail of the previous one; every kernel executes griddepcontrol.wait
// before it touches the previous kernel's output.  Measured on B200: +2-4 % at B <= 4 (launch-bound), nothing at B = 24 (one
// CTA per SM owns all shared memory / TMEM, the successor cannot become resident early) -> auto: on for small batches.
// SNTC_TC_PDL=0 / 1 forces it off / on.
inline bool tc_pdl(int batch) {
  static const int env = tc_env_int("SNTC_TC_PDL", -1);
  return env >= 0 ? env != 0 : batch <= 4;
}
inline int tc_cta_group() {   // read when a model is finalized (not cached: tests build both variants in one process)
  const int v = tc_env_int("SNTC_TC_CTA_GROUP", 2);
  return (v != 1 && v != 2) ? 2 : v;
}
// Fewest n-tiles of at most bn_max columns, then the narrowest (balanced) tile that achieves it; tiles are
// multiples of `unit`.  The last tile of a band only issues MMAs for its own columns.
inline int tc_choose_bn(int N, int unit, int bn_cap) {
  int cap = (bn_cap / unit) * unit;
  if (cap < unit) cap = unit;
  int tiles = (N + cap - 1) / cap;
  int bn = ((N + tiles - 1) / tiles + unit - 1) / unit * unit;
  return std::min(bn, cap);
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

inline bool tc_make_map_4d(TcDriver& drv, CUtensorMap* map, void* base, int C, int w, int h, int B, int TW, int TH, std::string* err, int TB = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)w * C * 2, (cuuint64_t)h * w * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = drv.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled(4d) failed: " + std::to_string((int)r); return false; }
  return true;
}

// A conv layer runs on the tensor cores when its input channel count is TMA-addressable.
inline bool tc_conv_supported(const ConvLayer& c) {
  return !c.append_ones && c.cin % 8 == 0 && c.cin >= 64 && (int)c.bands.size() <= TC_MAX_BANDS;
}

// `pixel_cols` > 0: n-tiles must hold whole output pixels of that many columns (fused two-layer epilogue).
inline bool tc_pack_conv(TcDriver& drv, const ConvLayer& c, const HostWeights& hw, TcConv& t, int pixel_cols, std::vector<void*>& owned, std::string* err,
                         int bn_cap_override = 0, int extra_reserve = 0) {
  t.cin = c.cin;
  t.kblocks = (c.cin + TC_BK - 1) / TC_BK;
  int rem = c.cin - (t.kblocks - 1) * TC_BK;
  t.last_kmma = (rem + 15) / 16;
  t.ktot_pad = t.kblocks * TC_BK;
  t.fused_two_layer = pixel_cols > 0;
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
  t.cg = tc_cta_group();
  const int base_unit = 16 * t.cg;                                                   // each CTA of a pair holds BN/2 rows of W
  int unit = base_unit;
  if (pixel_cols > 0) { unit = pixel_cols; while (unit % base_unit) unit += pixel_cols; }   // lcm(pixel_cols, 16*cg)
  // heavy bands first: with items dealt round-robin to the persistent CTAs this balances the tail
  std::vector<size_t> order(c.bands.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return c.bands[a].Ty * c.bands[a].Tx > c.bands[b].Ty * c.bands[b].Tx; });
  std::vector<size_t> row0_of(c.bands.size());
  size_t row0 = 0;
  for (size_t bi = 0; bi < c.bands.size(); ++bi) {
    const Band& b = c.bands[bi];
    row0_of[bi] = row0;
    for (int fy = 0; fy < b.nphy; ++fy) for (int fx = 0; fx < b.nphx; ++fx) for (int co = 0; co < c.cout; ++co) {
      size_t row = row0 + (size_t)(fy * b.nphx + fx) * c.cout + co;
      for (int jy = 0; jy < b.Ty; ++jy) for (int jx = 0; jx < b.Tx; ++jx) {
        int ay = band_tap_index(c, b.phy0, fy, jy), ax = band_tap_index(c, b.phx0, fx, jx);
        if (ay < 0 || ax < 0 || ay >= c.k || ax >= c.k) continue;
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
  t.items_per_mtile.clear();
  // one band table (n-tiling + W tensor maps) per tile-width cap over the same packed weights
  auto build_tiling = [&](int bn_cap, std::vector<TcBandDev>& bands, int& nbands, int& bn_max, int& stages, TcBandDev*& d_bands) -> bool {
    bands.clear();
    bn_max = 16;
    for (size_t oi = 0; oi < order.size(); ++oi) {
      size_t bi = order[oi];
      const Band& b = c.bands[bi];
      if (b.N == 0) continue;
      TcBandDev d;
      memset(&d, 0, sizeof(d));
      d.phy0 = b.phy0; d.nphx = b.nphx; d.phx0 = b.phx0; d.Ty = b.Ty; d.Tx = b.Tx;
      // band index -> (yi, xi) in row-major order of (by, bx)
      d.mloy = c.by[bi / c.bx.size()].mlo; d.mlox = c.bx[bi % c.bx.size()].mlo;
      d.N = b.N;
      d.oshift = c.merged ? c.s * c.dlo + c.p : 0;
      if (b.N <= bn_cap) d.BN = (b.N + base_unit - 1) / base_unit * base_unit;   // single tile: no pixel-alignment constraint
      else d.BN = tc_choose_bn(b.N, unit, bn_cap);
      d.ntiles = (b.N + d.BN - 1) / d.BN;
      bn_max = std::max(bn_max, d.BN);
      uint64_t kcols = (uint64_t)std::max(1, b.Ty * b.Tx) * t.ktot_pad;
      if (!tc_make_map_2d(drv, &d.mapBhi, t.d_hi + row0_of[bi] * kmax, kcols, (uint64_t)b.N, kmax, TC_BK, (uint32_t)(d.BN / t.cg), err)) return false;
      if (!tc_make_map_2d(drv, &d.mapBlo, t.d_lo + row0_of[bi] * kmax, kcols, (uint64_t)b.N, kmax, TC_BK, (uint32_t)(d.BN / t.cg), err)) return false;
      bands.push_back(d);
    }
    nbands = (int)bands.size();
    if (bn_max > TC_ACC_COLS) { *err = "n-tile wider than a TMEM accumulator"; return false; }
    int stage_bytes = 2 * TC_BM * 128 + 2 * (bn_max / t.cg) * 128;
    const int reserve = extra_reserve + 2048 + (c.cout + (pixel_cols >= 48 ? 48 * 48 + 48 + 8 : 700)) * 4;   // alignment slack + barriers + epilogue constants (bias | gamma | beta)
    stages = std::min(8, (227 * 1024 - reserve) / stage_bytes);
    if (stages < 2) { *err = "not enough shared memory for a 2-stage pipeline"; return false; }
    int prefix = 0;
    for (auto& bd : bands) { bd.item_begin = prefix; prefix += bd.ntiles; }
    if (cudaMalloc((void**)&d_bands, sizeof(TcBandDev) * std::max(1, nbands)) != cudaSuccess) { *err = "cudaMalloc (band table) failed"; return false; }
    owned.push_back(d_bands);
    if (nbands > 0 && cudaMemcpy(d_bands, bands.data(), sizeof(TcBandDev) * nbands, cudaMemcpyHostToDevice) != cudaSuccess) { *err = "cudaMemcpy (band table) failed"; return false; }
    return true;
  };
  if (!build_tiling(bn_cap_override > 0 ? bn_cap_override : tc_bn_max(), t.bands, t.nbands, t.bn_max, t.stages, t.d_bands)) return false;
  const int narrow_cap = std::max(64 / unit * unit, unit);
  if (narrow_cap < t.bn_max && !build_tiling(narrow_cap, t.narrow.bands, t.narrow.nbands, t.narrow.bn_max, t.narrow.stages, t.narrow.d_bands)) return false;
  t.ok = t.nbands > 0;
  return true;
}

inline bool tail_tc_supported(const ConvLayer& c);
inline bool tail_tc_pack(const ConvLayer& c, const HostWeights& hw, TailTc& t, std::vector<void*>& owned, std::string* err);

inline bool tc_finalize(TcDriver& drv, TcModelState& st, Transform* hyper, Transform* syn, const HostWeights& hw,
                        std::vector<void*>& owned, std::string* err) {
  if (!drv.encode) { *err = drv.err.empty() ? "cuTensorMapEncodeTiled unavailable" : drv.err; return false; }
  auto pack = [&](Transform* t, std::vector<TcConv>& out) {
    if (!t) return true;
    out.resize(t->convs.size());
    for (size_t i = 0; i < t->convs.size(); ++i) {
      const ConvLayer& c = t->convs[i];
      if (!tc_conv_supported(c)) continue;
      if (c.col2im) {   // final layer in col2im form: pack its 1x1 contraction; n-tiles hold whole channels (<= 96 columns)
        if (!tc_pack_conv(drv, c.c2i[0], hw, out[i], c.c2i_kp, owned, err, 96, TC_BM * 97 * 4)) return false;
        out[i].fused_two_layer = false;
        continue;
      }
      // two-layer synthesis layer 1 followed by the pointwise IGDN(+res) stage: fuse it when C1 is 12 or 24
      int pixel_cols = 0;
      for (size_t oi = 0; oi + 1 < t->ops.size(); ++oi) {
        if ((t->ops[oi].type == OP_CONVT) && t->ops[oi].conv == (int)i) {
          const Op& nx = t->ops[oi + 1];
          if (nx.type == OP_ACT_RES && (c.cout == 24 || c.cout == 48 || c.cout == 96)) pixel_cols = c.cout;
          if (nx.type == OP_GDN && t->gdns[nx.gdn].kind == GDN_1 && (c.cout == 12 || c.cout == 24 || c.cout == 48) && (t->kind == SNTC_T_TWO_LAYER)) pixel_cols = c.cout;
        }
      }
      if (!tc_pack_conv(drv, c, hw, out[i], pixel_cols, owned, err)) return false;
    }
    return true;
  };
  if (!pack(hyper, st.hyper) || !pack(syn, st.syn)) return false;
  if (syn && tc_env_int("SNTC_TC_GDN", 1)) {
    // GDN stages between tensor-core layers (mbt2018 / bls2017 / cnn): norm pool on the tensor cores; SNTC_TC_GDN=0 keeps
    // the FFMA band GEMM
    st.syn_gdn.resize(syn->gdns.size());
    for (size_t oi = 1; oi < syn->ops.size(); ++oi) {
      const Op& op = syn->ops[oi];
      if (op.type != OP_GDN || syn->ops[oi - 1].type != OP_CONVT) continue;
      const int prev = syn->ops[oi - 1].conv;
      if (prev >= (int)st.syn.size() || !st.syn[prev].ok) continue;
      const GdnLayer& g = syn->gdns[op.gdn];
      if (g.C < 64 || g.C % 16 != 0 || !g.d_beta) continue;
      TcGdn& tg = st.syn_gdn[op.gdn];
      ConvLayer& c = tg.conv;
      c.k = 1; c.s = 1; c.p = 0; c.cin = g.C; c.cout = g.C; c.layout = LAYOUT_TFC_IO; c.has_bias = true; c.act = SNTC_ACT_NONE;
      c.sources = {ConvSource{g.gamma, g.beta, g.C}};
      finish_conv(c);
      c.d_bias = g.d_beta;
      if (!tc_pack_conv(drv, c, hw, tg.tc, 0, owned, err)) return false;
      tg.ok = tg.tc.ok;
    }
  }
  const int tail_env = tc_env_int("SNTC_TC_TAIL", -1);   // -1 = auto: hidden widths the warp-MMA tail does not take (C1 > 16)
  if (syn && tail_env != 0) {
    // Tail of a two-layer synthesis on the tensor cores (sntc_kernels_tail_tc.cuh).  Default (auto): hidden widths > 16,
    // where the warp-MMA tail (sntc_kernels_tail_mma.cuh) does not apply -- C1 = 24: 0.60 -> 0.22 ms per 8 x 1200x1200 step
    // against the FFMA tail.  At C1 = 12 it loses to the warp-MMA tail (0.19 vs 0.13 ms): every M=128 x K=16 MMA costs
    // ~95 clk of A-operand delivery however small N is, and the octet-planar fp16 stores slow the layer-1 epilogue.
    // SNTC_TC_TAIL=1 / 0 forces it on / off.
    st.syn_tail.resize(syn->convs.size());
    for (size_t oi = 2; oi < syn->ops.size(); ++oi) {
      const Op& op = syn->ops[oi];
      if (op.type != OP_CONVT_RGB || syn->ops[oi - 2].type != OP_CONVT) continue;
      const int prev = syn->ops[oi - 2].conv;
      if (prev >= (int)st.syn.size() || !st.syn[prev].ok || !st.syn[prev].fused_two_layer) continue;
      const ConvLayer& c = syn->convs[op.conv];
      if (!tail_tc_supported(c)) continue;
      if (tail_env < 0 && c.cin <= 16) continue;   // tail_mma_supported(): measured faster there (0.10 vs 0.17 ms per 24-image step)
      if (!tail_tc_pack(c, hw, st.syn_tail[op.conv], owned, err)) return false;
    }
  }
  if (!st.smem_attr_set) {
    cudaError_t e = cudaFuncSetAttribute(band_gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(band_gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
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
  double* rate_slots = nullptr; int* rate_slot_img = nullptr; RateConst rc{}; size_t rate_slot_cap = 0;   // bits_y partials (nullable)
  float* sigma_out = nullptr;                                                                              // raw sigma for rate_y_flat_kernel (nullable)
  // two-layer fusion: f32 receives t = act(base) (+ res), [B,hout,wout,C1]
  bool two_layer = false; int C1 = 0; bool has_res = false; int tl_act = SNTC_ACT_NONE; bool tl_inverse = true;
  const float* gamma = nullptr; int gamma_stride = 0; const float* beta = nullptr;
  const float* h_gamma = nullptr; const float* h_beta = nullptr;   // host copies ([C1][C1], [C1]) -> kernel parameters
  // GDN stages (see TcParams): pooled planes out / norm-pool GEMM epilogue
  int plane_xform = A_NONE; int gdn_mode = G_NONE; const float* gx = nullptr;
  bool sign_in_lo = false; const __half* gx_hi = nullptr; const __half* gx_lo = nullptr;   // see TcParams
  unsigned pass_mask = 3u; const unsigned* alo_flag = nullptr;   // see TcParams
  // col2im final layer (ConvLayer::col2im): the layer passed to tc_run_conv is its 1x1 contraction c2i[0]; these describe the ConvT
  bool col2im = false; int c2i_k = 0, c2i_s = 1, c2i_p = 0, c2i_kp = 0, c2i_cout = 0; const float* c2i_bias = nullptr;
};

// m-tile = TB images x TH x TW cells (= 128 rows of the TMA box, order [b][y][x]); TH * TW is a multiple of 32 so that an epilogue
// warp (32 consecutive rows) stays inside one image.  Fewest tiles wins; ties keep one image per tile and the historical order.
// 24 latent grids of 8 x 12 (hyper layer 0 of 768x512 images): 8 x 16 x 1 wastes a quarter of every tile (24 tiles), 8 x 4 x 4 tiles
// exactly (18) -- one wave of work items instead of two.  SNTC_TC_BATCH_TILES=0 keeps TB = 1.
inline void tc_choose_patch(int h, int w, int B, int* TH, int* TW, int* TB) {
  static const int cand[8][2] = {{8, 16}, {16, 8}, {4, 32}, {32, 4}, {2, 64}, {64, 2}, {1, 128}, {128, 1}};
  static const bool batch_tiles = tc_env_int("SNTC_TC_BATCH_TILES", 1) != 0;
  long best = -1;
  for (int tb = 1; tb <= (batch_tiles ? 4 : 1); tb *= 2)
    for (auto& c : cand) {
      if (c[0] % tb != 0 && c[1] % tb != 0) continue;
      // split the y extent of the 128-row candidate by tb when possible, else the x extent
      const int th = c[0] % tb == 0 ? c[0] / tb : c[0], tw = c[0] % tb == 0 ? c[1] : c[1] / tb;
      if (th * tw % 32 != 0) continue;
      long cost = (long)((h + th - 1) / th) * ((w + tw - 1) / tw) * ((B + tb - 1) / tb);
      if (best < 0 || cost < best) { best = cost; *TH = th; *TW = tw; *TB = tb; }
    }
}

// Number of bits_y partial slots the hyper-final epilogue of this layer writes for a [B,h,w] input.
inline size_t tc_rate_slots(const TcConv& t, int B, int h, int w, int num_sms) {
  int TH, TW, TB;
  tc_choose_patch(h, w, B, &TH, &TW, &TB);
  const int mtiles = ((h + TH - 1) / TH) * ((w + TW - 1) / TW) * ((B + TB - 1) / TB);
  const int groups = (mtiles + t.cg - 1) / t.cg;
  size_t items = 0;
  for (auto& bd : (tc_use_narrow(t, mtiles, num_sms) ? t.narrow.bands : t.bands)) items += (size_t)groups * bd.ntiles;
  return items * t.cg * TC_EPI_WARPS;
}

// One persistent launch for all bands of one conv layer.  Input: fp16 planes [B,h,w,cin].
inline int tc_run_conv(TcDriver& drv, const ConvLayer& c, TcConv& t, const __half* in_hi, const __half* in_lo, int B, int h, int w,
                       const TcConvOut& o, cudaStream_t s, uint64_t* launches, std::string* err) {
  int TH, TW, TB;
  tc_choose_patch(h, w, B, &TH, &TW, &TB);
  if (o.col2im) { TH = 8; TW = 16; TB = 1; }   // overlapping tiles: 6 x 14 interior cells + a one-cell halo
  CUtensorMap mapAhi, mapAlo;
  if (!tc_make_map_4d(drv, &mapAhi, (void*)in_hi, c.cin, w, h, B, TW, TH, err, TB)) return TC_ERROR;
  if (!tc_make_map_4d(drv, &mapAlo, (void*)in_lo, c.cin, w, h, B, TW, TH, err, TB)) return TC_ERROR;
  TcParams P{};
  P.B = B; P.hin = h; P.win = w; P.s = c.s; P.p = c.p;
  P.TH = TH; P.TW = TW; P.tiles_y = (h + TH - 1) / TH; P.tiles_x = (w + TW - 1) / TW;
  P.tile_step_y = TH; P.tile_step_x = TW; P.tile_off = 0;
  if (o.col2im) {
    P.tile_step_y = TH - 2; P.tile_step_x = TW - 2; P.tile_off = -1;
    P.tiles_y = (h + P.tile_step_y - 1) / P.tile_step_y; P.tiles_x = (w + P.tile_step_x - 1) / P.tile_step_x;
    P.c2i_k = o.c2i_k; P.c2i_s = o.c2i_s; P.c2i_p = o.c2i_p; P.c2i_kp = o.c2i_kp;
  }
  P.TB = TB;
  const int mtiles = P.tiles_x * P.tiles_y * ((B + TB - 1) / TB);
  const int groups = (mtiles + t.cg - 1) / t.cg;
  P.mtiles = mtiles;
  // item order: m-tile-group major for the two-layer layer (its per-pixel IGDN epilogue is as long as the MMAs of a light
  // band: alternating bands keeps both busy, 0.226 -> 0.217 ms), band major elsewhere (hyper layer 0: 0.057 -> 0.052 ms);
  // SNTC_TC_ORDER = 0 / 1 forces band / group major everywhere
  static const int order_env = tc_env_int("SNTC_TC_ORDER", -1);
  const int order_mode = order_env >= 0 ? order_env : (o.two_layer ? 1 : 0);
  // tiling: wide n-tiles, or the narrow ones when the batch is too small to give every unit a wide item
  const bool narrow = tc_use_narrow(t, mtiles, drv.num_sms);
  std::vector<TcBandDev>& bands = narrow ? t.narrow.bands : t.bands;
  TcBandDev* const d_bands = narrow ? t.narrow.d_bands : t.d_bands;
  const int nbands = narrow ? t.narrow.nbands : t.nbands, bn_max = narrow ? t.narrow.bn_max : t.bn_max, stages = narrow ? t.narrow.stages : t.stages;
  // the device band table is static (uploaded once at finalize: no in-stream uploads, so the launches are safe to capture into
  // a CUDA graph and to overlap with programmatic dependent launch); only the item count depends on the batch
  int per_group = 0;
  for (auto& bd : bands) per_group += bd.ntiles;
  if (order_mode) P.ipg = per_group;
  static const int snake_env = tc_env_int("SNTC_TC_SNAKE", 1);
  P.snake = (!order_mode && snake_env) ? 1 : 0;
  const int item = per_group * groups;
  cudaError_t e = cudaSuccess;
  P.bands = d_bands; P.nbands = nbands; P.total_items = item;
  P.kblocks = t.kblocks; P.last_kmma = t.last_kmma;
  P.cout = c.cout; P.bn_max = bn_max; P.stages = stages;
  P.inv_scale = 1.f / t.scale; P.bias = c.d_bias; P.act = c.act;
  P.hout = h * c.s - c.out_crop; P.wout = w * c.s - c.out_crop;
  P.epi = o.hyper_final ? TC_EPI_HYPER_FINAL : (o.two_layer ? TC_EPI_TWO_LAYER : TC_EPI_PLAIN);
  if (o.col2im) {
    if (o.hyper_final || o.two_layer || o.hi || o.gdn_mode != G_NONE || !o.c2i_bias) { *err = "col2im layer: unsupported epilogue combination"; return TC_ERROR; }
    P.epi = TC_EPI_COL2IM; P.hout = h * o.c2i_s; P.wout = w * o.c2i_s; P.cout = o.c2i_cout; P.bias = o.c2i_bias;
  }
  P.out_hi = o.hi; P.out_lo = o.lo; P.out_f32 = o.f32; P.out_u8 = o.u8; P.out_crop = o.crop; P.H = o.H; P.W = o.W;
  P.q = o.q; P.q_kind = o.q_kind; P.Cy = o.Cy; P.max_index = o.max_index; P.trunc = o.trunc ? 1 : 0; P.y_hat = o.y_hat; P.idx = o.idx;
  P.C1 = o.C1; P.has_res = o.has_res ? 1 : 0; P.tl_act = o.tl_act; P.tl_inverse = o.tl_inverse ? 1 : 0;
  P.gamma = o.gamma; P.gamma_stride = o.gamma_stride; P.beta = o.beta;
  if (o.two_layer) {
    if ((o.C1 != 12 && o.C1 != 24 && o.C1 != 48) || (int)c.h_bias.size() < c.cout) { *err = "two-layer epilogue: hidden width must be 12, 24 or 48"; return TC_ERROR; }
    if (o.C1 <= 24) {   // narrow variants: constants as kernel parameters (the wide one stages gamma / beta in shared memory)
      for (int i = 0; i < o.C1 * o.C1; ++i) P.tl_gamma[i] = o.h_gamma ? o.h_gamma[i] : 0.f;
      for (int i = 0; i < o.C1; ++i) P.tl_beta[i] = o.h_beta ? o.h_beta[i] : 0.f;
      for (int i = 0; i < c.cout; ++i) P.tl_bias[i] = c.h_bias[i];
    } else if ((o.tl_act == SNTC_ACT_IGDN1 || o.tl_act == SNTC_ACT_GDN1) && (!o.gamma || !o.beta)) {
      *err = "two-layer epilogue: GDN parameters missing"; return TC_ERROR;
    }
  }
  P.rate_slots = o.rate_slots; P.rate_slot_img = o.rate_slot_img; P.rc = o.rc; P.sigma_out = o.sigma_out;
  P.plane_xform = o.plane_xform; P.gdn_mode = o.gdn_mode; P.gx = o.gx;
  P.sign_in_lo = o.sign_in_lo ? 1 : 0; P.gx_hi = o.gx_hi; P.gx_lo = o.gx_lo;
  if (o.gdn_mode != G_NONE && !o.gx && !(o.gx_hi && o.gx_lo)) { *err = "GDN stage: no source for x"; return TC_ERROR; }
  if (o.sign_in_lo && (o.plane_xform != A_ABS || !o.hi)) { *err = "sign_in_lo needs |x| planes"; return TC_ERROR; }
  P.pass_mask = o.pass_mask & 3u; P.alo_flag = o.alo_flag;
  {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; };
    bool ok = c.cout % 16 == 0 && (!o.hyper_final || o.Cy % 16 == 0) && !o.two_layer && !o.u8 && !o.crop;
    ok = ok && al(o.hi) && al(o.lo) && al(o.f32) && al(o.q) && al(o.y_hat) && al(o.idx) && al(o.gx) && al(o.sigma_out) && al(o.gx_hi) && al(o.gx_lo);
    // n-tiles must start on a 16-column boundary and the mma width is a multiple of 32 only when BN is: chunks are 32 wide,
    // the last one may be half-used
    for (auto& bd : bands) ok = ok && bd.BN % 16 == 0;
    P.vec16 = (ok && !o.col2im) ? 1 : 0;
  }
  P.rgb_runs = (!o.col2im && c.cout == 3 && o.u8 && !o.f32 && !o.crop && !o.hi && !o.hyper_final && !o.two_layer && tc_env_int("SNTC_TC_RGB_RUNS", 1)) ? 1 : 0;
  if ((o.plane_xform != A_NONE || o.gdn_mode != G_NONE) && !P.vec16) {
    *err = "GDN stage on the tensor cores: needs C % 16 == 0 and 32-byte aligned tensors"; return TC_ERROR;
  }
  if (o.sigma_out && !P.vec16) { *err = "rate term: needs Cy % 16 == 0 and 32-byte aligned tensors"; return TC_ERROR; }
  if (o.rate_slots && (!P.vec16 || (size_t)P.total_items * t.cg * TC_EPI_WARPS > o.rate_slot_cap)) {
    *err = "rate term: needs Cy % 16 == 0, 32-byte aligned tensors and a large enough slot buffer"; return TC_ERROR;
  }
  size_t smem = (size_t)stages * (2 * TC_BM * 128 + 2 * (bn_max / t.cg) * 128) + 1024 + 64 * 8 + (size_t)((o.two_layer ? o.C1 * o.C1 + o.C1 : 0) + c.cout + 8) * 4;
  if (o.col2im) smem += (size_t)TC_BM * (bn_max + 1) * 4;   // staging tile of the overlap-add epilogue
  if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: each CTA owns all 512 TMEM columns
  if (smem > 227 * 1024) { *err = "shared memory budget exceeded"; return TC_ERROR; }
  int units = std::min(drv.num_sms / t.cg, P.total_items);
  if (units <= 0) return TC_OK;
  static const bool trace_on = tc_env_int("SNTC_TC_TRACE", 0) != 0;
  long long* d_trace = nullptr;
  const size_t trace_n = (size_t)units * TC_TRACE_ITEMS * 8;
  if (trace_on) {
    if (cudaMalloc((void**)&d_trace, trace_n * 8) != cudaSuccess) { *err = "trace alloc failed"; return TC_ERROR; }
    cudaMemsetAsync(d_trace, 0, trace_n * 8, s);
    P.trace = d_trace;
  }
  if (t.cg == 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * units)); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = tc_pdl(B) ? 2 : 1;
    e = cudaLaunchKernelEx(&cfg, band_gemm_tc_kernel<2>, mapAhi, mapAlo, P);
    if (e != cudaSuccess) { *err = std::string("band_gemm_tc_kernel<2> launch: ") + cudaGetErrorString(e); return TC_ERROR; }
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)units); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = tc_pdl(B) ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, band_gemm_tc_kernel<1>, mapAhi, mapAlo, P);
    if (e != cudaSuccess) { *err = std::string("band_gemm_tc_kernel<1> launch: ") + cudaGetErrorString(e); return TC_ERROR; }
  }
  if (launches) (*launches)++;
  e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("band_gemm_tc_kernel launch: ") + cudaGetErrorString(e); return TC_ERROR; }
  if (trace_on) {   // debug only: synchronous dump of the per-item timeline of a few units
    std::vector<long long> h(trace_n);
    cudaStreamSynchronize(s);
    cudaMemcpy(h.data(), d_trace, trace_n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_trace);
    fprintf(stderr, "[tc-trace] layer cin=%d cout=%d k=%d s=%d items=%d units=%d cg=%d stages=%d bn_max=%d\n", c.cin, c.cout, c.k, c.s, P.total_items, units, t.cg, stages, bn_max);
    for (int u : {0, units / 2, units - 1}) {
      const long long* b = h.data() + (size_t)u * TC_TRACE_ITEMS * 8;
      long long t0 = b[0];
      for (int j = 0; j < TC_TRACE_ITEMS && b[j * 8 + 5] != 0; ++j) {
        const long long* r = b + j * 8;
        fprintf(stderr, "[tc-trace]  unit %3d item %5lld: tma %7lld..%7lld  mma start %7lld first-data %7lld issued %7lld  epi %7lld..%7lld\n", u, r[7],
                r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0, r[4] - t0, r[5] - t0, r[6] - t0);
      }
    }
  }
  return TC_OK;
}

}  // namespace sntc
