// libsntc.so -- C ABI implementation (see include/sntc.h).  One translation unit: host plan logic
// (sntc_plan.hpp), fp32 CUDA-core kernels (sntc_kernels_f32.cuh), tcgen05 kernels (sntc_kernels_tc.cuh).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the entry points are resolved with dlsym (no link-time dependency on libnccl)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <memory>

#include "../../include/sntc.h"
#include "sntc_plan.hpp"
#include "sntc_kernels_f32.cuh"
#include "sntc_kernels_rate.cuh"
#include "sntc_kernels_tc.cuh"
#include "sntc_kernels_tail_tc.cuh"
#include "sntc_kernels_tail_mma.cuh"
#include "sntc_kernels_tail_tz.cuh"
#include "sntc_kernels_msssim.cuh"
#include "sntc_kernels_lpips.cuh"
#include "sntc_kernels_vjp.cuh"
#include "sntc_coder.hpp"

using namespace sntc;

static thread_local std::string g_err;

static inline bool is_tc(int precision) { return precision == SNTC_PRECISION_TC_F16X3 || precision == SNTC_PRECISION_TC_F16X3_SYN2; }

static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail(SNTC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));              \
  } while (0)

#define TRY(expr)            \
  do {                       \
    int r__ = (expr);        \
    if (r__ != SNTC_OK) return r__; \
  } while (0)

struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return SNTC_OK;
    if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return fail(SNTC_E_CUDA, std::string("cudaFree: ") + cudaGetErrorString(e)); }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; return fail(SNTC_E_CUDA, std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e)); }
    cap = want;
    return SNTC_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct sntc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
  uint64_t kinds[SNTC_LAUNCH_KINDS] = {0, 0, 0, 0, 0, 0};   // by kernel family (sntc_launch_counts); [0] unused (= launches)
  cudaDeviceProp prop{};
  TcDriver tc;   // cuTensorMapEncodeTiled entry point etc.
  DevBuf ms_ws;  // scratch of sntc_image_msssim
};

struct GraphKey { long long v[6]; const void* p[20]; };
struct GraphEntry { GraphKey key{}; cudaGraphExec_t exec = nullptr; bool failed = false; uint64_t launches = 0; uint64_t kinds[SNTC_LAUNCH_KINDS] = {0, 0, 0, 0, 0, 0}; };

struct ProfRec { std::string label; cudaEvent_t a = nullptr, b = nullptr; double macs = 0; };
struct ProfAgg { std::string label; float ms = 0; int n = 0; double macs = 0; };

// Decoder backward (sntc_*_vjp): an fp32 plan of the transform (every conv a band GEMM, every op output kept) and, per conv,
// its input-gradient as a forward stride-1 layer (sntc_plan.hpp make_backward_conv), packed for the tensor cores when the
// model's precision asks for them.
struct VjpPlan {
  bool ok = false; std::string why;
  Transform fwd;
  std::vector<ConvLayer> bconv;         // parallel to fwd.convs
  std::vector<TcConv> btc;              // parallel to fwd.convs (tensor-core precision only)
  std::vector<float*> d_gamma_t;        // parallel to fwd.gdns: gamma^T [out][in] for the wide GDN adjoint (C > 64)
  std::vector<DevBuf> acts;             // forward output of every op
  DevBuf g[2], s2d, pl[2], ones, st_x, st_g, st_gin, st_out;
  void release() {
    for (auto& b : acts) b.release();
    for (DevBuf* b : {&g[0], &g[1], &s2d, &pl[0], &pl[1], &ones, &st_x, &st_g, &st_gin, &st_out}) b->release();
  }
};

struct sntc_model {
  sntc_ctx* ctx = nullptr;
  bool prof_on = false;
  std::vector<ProfRec> prof;            // appended while profiling is enabled
  std::vector<cudaEvent_t> prof_pool;   // recycled events
  std::vector<ProfAgg> prof_agg;
  bool prof_agg_valid = false;
  sntc_model_desc desc{};
  Transform hyper, syn;
  bool has_hyper = false, has_syn = true;
  std::vector<VarSpec> vars;
  HostWeights hw;
  bool finalized = false;
  std::vector<void*> owned;            // device allocations (weights)
  DevBuf ws_a, ws_b, ws_c;             // ping-pong activations + misc
  DevBuf st_z, st_q, st_u8, st_idx, st_yhat, st_f32, st_orig;   // staging for host tensors
  DevBuf d_hs, d_yhat, d_ssd;
  DevBuf d_rate_slots, d_rate_img, d_rate_zslots, d_rate;   // rate term: partial slots, slot->image map, per-image [bits_y, bits_z]
  float* d_prior = nullptr;             // NoisyDeepFactorized parameters [Cz][DF_STRIDE] (softplus / tanh applied)
  double* h_rate = nullptr; int h_rate_cap = 0;   // pinned
  DevBuf d_flag;                        // split_planes_kernel's "lo plane is non-zero" flag (one unsigned)
  DevBuf d_mu;                          // two-phase decode: mu of the last sntc_decode_hyper [B,hy,wy,Cy] f32
  int ph_B = 0, ph_hy = 0, ph_wy = 0;   // geometry of that call (0 = none pending)
  TcModelState tc;                      // tensor-core plan state (tensor maps, fp16 planes)
  std::vector<TailMma> tail_mma;       // parallel to syn.convs: warp-MMA tail of a two-layer synthesis (tc precision only)
  std::vector<TailTz> tail_tz;         // parallel to syn.convs: window-GEMM tcgen05 tail (hidden width 12; takes precedence over tail_mma)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool ev_valid = false;
  bool graphs_on = true;                // sntc_model_enable_graphs
  std::vector<GraphEntry> graphs;       // CUDA graphs of small-batch decodes, keyed by shapes + pointers (decode_impl)
  unsigned long long* h_ssd = nullptr;  // pinned
  int h_ssd_cap = 0;
  bool want_vjp = false;                // sntc_model_enable_vjp: finalize also builds the backward plans
  VjpPlan vjp_hyper, vjp_syn;
};

// ------------------------------------------------------------------------------------------------
extern "C" int sntc_version(void) { return SNTC_VERSION; }
extern "C" const char* sntc_last_error(void) { return g_err.c_str(); }

extern "C" int sntc_create(int device, sntc_ctx** out) {
  if (!out) return fail(SNTC_E_INVALID, "sntc_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SNTC_E_CUDA, std::string("sntc_create: no CUDA device (") + cudaGetErrorString(e) + "); libsntc has no CPU fallback");
  if (device < 0 || device >= n) return fail(SNTC_E_INVALID, "sntc_create: device index out of range");
  CU_TRY(cudaSetDevice(device));
  auto ctx = std::make_unique<sntc_ctx>();
  ctx->device = device;
  CU_TRY(cudaGetDeviceProperties(&ctx->prop, device));
  if (ctx->prop.major != 10)
    return fail(SNTC_E_CUDA, "sntc_create: device is sm_" + std::to_string(ctx->prop.major * 10 + ctx->prop.minor) +
                                 "; libsntc is built for sm_100a (B200) only");
  CU_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->tc.init();   // resolves the driver entry points lazily; failure is reported when the TC path is requested
  *out = ctx.release();
  return SNTC_OK;
}

extern "C" int sntc_destroy(sntc_ctx* ctx) {
  if (!ctx) return SNTC_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  ctx->ms_ws.release();
  delete ctx;
  return SNTC_OK;
}

extern "C" int sntc_sync(sntc_ctx* ctx) {
  if (!ctx) return fail(SNTC_E_INVALID, "sntc_sync: ctx is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  CU_TRY(cudaGetLastError());
  return SNTC_OK;
}

extern "C" void* sntc_stream(sntc_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int sntc_device_name(sntc_ctx* ctx, char* buf, size_t n) {
  if (!ctx || !buf || n == 0) return fail(SNTC_E_INVALID, "sntc_device_name: bad argument");
  snprintf(buf, n, "%s", ctx->prop.name);
  return SNTC_OK;
}

extern "C" int sntc_device_pci_bus_id(sntc_ctx* ctx, char* buf, size_t n) {
  if (!ctx || !buf || n < 13) return fail(SNTC_E_INVALID, "sntc_device_pci_bus_id: bad argument (buffer must hold 13 bytes)");
  CU_TRY(cudaDeviceGetPCIBusId(buf, (int)n, ctx->device));
  return SNTC_OK;
}

extern "C" uint64_t sntc_launch_count(sntc_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int sntc_launch_counts(sntc_ctx* ctx, uint64_t out[SNTC_LAUNCH_KINDS]) {
  if (!ctx || !out) return fail(SNTC_E_INVALID, "sntc_launch_counts: bad argument");
  for (int i = 0; i < SNTC_LAUNCH_KINDS; ++i) out[i] = ctx->kinds[i];
  out[SNTC_LAUNCH_TOTAL] = ctx->launches;
  return SNTC_OK;
}

// ------------------------------------------------------------------------------------------------
// memory / event helpers
extern "C" int sntc_malloc(sntc_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return fail(SNTC_E_INVALID, "sntc_malloc: bad argument");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaMalloc(out, bytes ? bytes : 1));
  return SNTC_OK;
}
extern "C" int sntc_free(sntc_ctx* ctx, void* p) {
  if (!ctx) return fail(SNTC_E_INVALID, "sntc_free: ctx is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaFree(p));
  return SNTC_OK;
}
extern "C" int sntc_host_alloc(sntc_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return fail(SNTC_E_INVALID, "sntc_host_alloc: bad argument");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return SNTC_OK;
}
extern "C" int sntc_host_alloc_flags(sntc_ctx* ctx, size_t bytes, unsigned flags, void** out) {
  if (!ctx || !out) return fail(SNTC_E_INVALID, "sntc_host_alloc_flags: bad argument");
  if (flags & ~(unsigned)SNTC_HOST_WRITE_COMBINED) return fail(SNTC_E_INVALID, "sntc_host_alloc_flags: unknown flag");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaHostAlloc(out, bytes ? bytes : 1, (flags & SNTC_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return SNTC_OK;
}
extern "C" int sntc_host_free(sntc_ctx* ctx, void* p) {
  if (!ctx) return fail(SNTC_E_INVALID, "sntc_host_free: ctx is NULL");
  CU_TRY(cudaFreeHost(p));
  return SNTC_OK;
}
static cudaStream_t pick_stream(sntc_ctx* ctx, void* stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

extern "C" int sntc_memcpy_h2d(sntc_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream) {
  if (!ctx) return fail(SNTC_E_INVALID, "sntc_memcpy_h2d: ctx is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, pick_stream(ctx, stream)));
  return SNTC_OK;
}
extern "C" int sntc_memcpy_d2h(sntc_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream) {
  if (!ctx) return fail(SNTC_E_INVALID, "sntc_memcpy_d2h: ctx is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, pick_stream(ctx, stream)));
  return SNTC_OK;
}
extern "C" int sntc_memset(sntc_ctx* ctx, void* dst, int value, size_t bytes, void* stream) {
  if (!ctx) return fail(SNTC_E_INVALID, "sntc_memset: ctx is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaMemsetAsync(dst, value, bytes, pick_stream(ctx, stream)));
  return SNTC_OK;
}
extern "C" int sntc_stream_create(sntc_ctx* ctx, void** out) {
  if (!ctx || !out) return fail(SNTC_E_INVALID, "sntc_stream_create: bad argument");
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st;
  CU_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *out = st;
  return SNTC_OK;
}
extern "C" int sntc_stream_destroy(sntc_ctx* ctx, void* stream) {
  if (!ctx || !stream) return fail(SNTC_E_INVALID, "sntc_stream_destroy: bad argument");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaStreamDestroy((cudaStream_t)stream));
  return SNTC_OK;
}
extern "C" int sntc_stream_wait_event(sntc_ctx* ctx, void* stream, void* event) {
  if (!ctx || !event) return fail(SNTC_E_INVALID, "sntc_stream_wait_event: bad argument");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaStreamWaitEvent(pick_stream(ctx, stream), (cudaEvent_t)event, 0));
  return SNTC_OK;
}
extern "C" int sntc_stream_sync(sntc_ctx* ctx, void* stream) {
  if (!ctx) return fail(SNTC_E_INVALID, "sntc_stream_sync: ctx is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaStreamSynchronize(pick_stream(ctx, stream)));
  return SNTC_OK;
}
extern "C" int sntc_event_create(sntc_ctx* ctx, void** out) {
  if (!ctx || !out) return fail(SNTC_E_INVALID, "sntc_event_create: bad argument");
  CU_TRY(cudaSetDevice(ctx->device));
  cudaEvent_t ev;
  CU_TRY(cudaEventCreate(&ev));
  *out = ev;
  return SNTC_OK;
}
extern "C" int sntc_event_destroy(sntc_ctx* ctx, void* ev) {
  if (!ctx) return fail(SNTC_E_INVALID, "sntc_event_destroy: ctx is NULL");
  CU_TRY(cudaEventDestroy((cudaEvent_t)ev));
  return SNTC_OK;
}
extern "C" int sntc_event_record(sntc_ctx* ctx, void* ev, void* stream) {
  if (!ctx || !ev) return fail(SNTC_E_INVALID, "sntc_event_record: bad argument");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaEventRecord((cudaEvent_t)ev, pick_stream(ctx, stream)));
  return SNTC_OK;
}
extern "C" int sntc_event_elapsed_ms(sntc_ctx* ctx, void* start, void* stop, float* ms) {
  if (!ctx || !start || !stop || !ms) return fail(SNTC_E_INVALID, "sntc_event_elapsed_ms: bad argument");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaEventSynchronize((cudaEvent_t)stop));
  CU_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return SNTC_OK;
}

// ------------------------------------------------------------------------------------------------
// model
static void mark_final_op(Transform& t, int precision) {
  // The last conv of a synthesis transform with <= 3 output channels and only tiny bands runs on the
  // cell kernel (all phases per thread); otherwise it stays a band GEMM with the pixel epilogue.
  if (t.ops.empty()) return;
  Op& last = t.ops.back();
  if (last.type != OP_CONVT) return;
  ConvLayer& c = t.convs[last.conv];
  int maxN = 0;
  for (auto& b : c.bands) maxN = std::max(maxN, b.N);
  if (c.cout <= 3 && maxN < 64 && (c.s == 1 || c.s == 2 || c.s == 4)) {
    // tensor-core precision and a TMA-addressable input (deep decoders: 192 / 256 channels): all s*s output residues of a
    // cell become ONE band (N = s*s*cout) of the tcgen05 band GEMM; SNTC_TC_FINAL=0 keeps the CUDA-core cell kernel
    // SNTC_TC_COL2IM (default 1): wide inputs take the col2im form instead (one 1x1 GEMM per input pixel + overlap-add epilogue)
    if (is_tc(precision) && tc_conv_supported(c) && col2im_eligible(c) && tc_env_int("SNTC_TC_FINAL", 1) && tc_env_int("SNTC_TC_COL2IM", 1)) {
      c.col2im = true;
      c.c2i.clear();
      c.c2i.push_back(make_col2im_conv(c));
      c.c2i_kp = c.c2i[0].c2i_kp;
    } else if (is_tc(precision) && tc_conv_supported(c) && c.s > 1 && tc_env_int("SNTC_TC_FINAL", 1)) finish_conv_merged(c);
    else last.type = OP_CONVT_RGB;
  }
}

extern "C" int sntc_model_create(sntc_ctx* ctx, const sntc_model_desc* desc, sntc_model** out) {
  if (!ctx || !desc || !out) return fail(SNTC_E_INVALID, "sntc_model_create: bad argument");
  *out = nullptr;
  if (desc->struct_size != (int32_t)sizeof(sntc_model_desc))
    return fail(SNTC_E_INVALID, "sntc_model_create: desc.struct_size mismatch (ABI)");
  if (desc->precision != SNTC_PRECISION_FP32 && !is_tc(desc->precision))
    return fail(SNTC_E_INVALID, "sntc_model_create: unknown precision");
  auto m = std::make_unique<sntc_model>();
  m->ctx = ctx;
  m->desc = *desc;
  if (m->desc.num_scales <= 0) m->desc.num_scales = 64;
  try {
    m->has_hyper = desc->hyper.kind != SNTC_T_NONE;
    if (m->has_hyper) {
      if (desc->hyper.kind < SNTC_T_HYPER_SYNTHESIS || desc->hyper.kind > SNTC_T_HYPER_SMALL)
        return fail(SNTC_E_INVALID, "sntc_model_create: hyper.kind is not a hyper-synthesis class");
      m->hyper = build_transform(desc->hyper, "hyper_synthesis");
    }
    m->has_syn = desc->synthesis.kind != SNTC_T_NONE;   // hyper-only models serve the standalone transform call
    if (!m->has_syn && !m->has_hyper) return fail(SNTC_E_INVALID, "sntc_model_create: model has no transform");
    if (m->has_syn) {
      if (desc->synthesis.kind < SNTC_T_JPEG_LIKE_SYNTHESIS)
        return fail(SNTC_E_INVALID, "sntc_model_create: synthesis.kind is not a synthesis class");
      m->syn = build_transform(desc->synthesis, "synthesis");
    }
  } catch (const std::exception& e) {
    return fail(SNTC_E_INVALID, std::string("sntc_model_create: ") + e.what());
  }
  if (m->has_hyper && m->has_syn && m->hyper.out_channels != 2 * m->syn.in_channels)
    return fail(SNTC_E_INVALID, "sntc_model_create: hyper-synthesis must output 2*Cy channels (mu || sigma)");
  if (m->has_syn && m->syn.in_channels % 4 != 0) return fail(SNTC_E_UNSUPPORTED, "sntc_model_create: latent channels must be a multiple of 4");
  if (m->has_syn && m->syn.out_channels > 3) return fail(SNTC_E_UNSUPPORTED, "sntc_model_create: more than 3 image channels");
  if (m->has_syn) mark_final_op(m->syn, desc->precision);
  for (Transform* t : {&m->hyper, &m->syn})
    for (auto& c : t->convs)
      if (!c.append_ones && c.cin % 4 != 0)
        return fail(SNTC_E_UNSUPPORTED, "sntc_model_create: layer input channels must be a multiple of 4");
  if (m->has_hyper) for (auto& v : transform_variables(m->hyper)) m->vars.push_back(v);
  if (m->has_syn) for (auto& v : transform_variables(m->syn)) m->vars.push_back(v);
  if (desc->prior != SNTC_PRIOR_NONE) {
    if (desc->prior != SNTC_PRIOR_DEEP_FACTORIZED) return fail(SNTC_E_INVALID, "sntc_model_create: unknown prior");
    if (!m->has_hyper) return fail(SNTC_E_INVALID, "sntc_model_create: a hyper-latent prior needs a hyper-synthesis transform");
    // tfc.NoisyDeepFactorized(batch_shape=(Cz,)) with the default num_filters=(3,3,3)   mshyper/models.py:135
    const int f[5] = {1, 3, 3, 3, 1};
    const int64_t Cz = m->hyper.in_channels;
    for (int i = 0; i < 4; ++i) {
      m->vars.push_back({"prior.matrix_" + std::to_string(i), {Cz, f[i + 1], f[i]}});
      m->vars.push_back({"prior.bias_" + std::to_string(i), {Cz, f[i + 1], 1}});
      if (i < 3) m->vars.push_back({"prior.factor_" + std::to_string(i), {Cz, f[i + 1], 1}});
    }
  }
  *out = m.release();
  return SNTC_OK;
}

extern "C" int sntc_model_destroy(sntc_model* m) {
  if (!m) return SNTC_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  for (void* p : m->owned) cudaFree(p);
  for (DevBuf* b : {&m->ws_a, &m->ws_b, &m->ws_c, &m->st_z, &m->st_q, &m->st_u8, &m->st_idx, &m->st_yhat, &m->st_f32,
                    &m->st_orig, &m->d_hs, &m->d_yhat, &m->d_ssd, &m->d_rate_slots, &m->d_rate_img, &m->d_rate_zslots, &m->d_rate, &m->d_mu, &m->d_flag})
    b->release();
  m->tc.release();
  m->vjp_hyper.release(); m->vjp_syn.release();
  for (auto& g : m->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  for (auto& e : m->ev) if (e) cudaEventDestroy(e);
  for (auto& r : m->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto& e : m->prof_pool) cudaEventDestroy(e);
  if (m->h_ssd) cudaFreeHost(m->h_ssd);
  if (m->h_rate) cudaFreeHost(m->h_rate);
  delete m;
  return SNTC_OK;
}

extern "C" int sntc_model_enable_graphs(sntc_model* m, int on) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_model_enable_graphs: model is NULL");
  m->graphs_on = on != 0;
  return SNTC_OK;
}

extern "C" int sntc_model_num_variables(sntc_model* m) { return m ? (int)m->vars.size() : 0; }

extern "C" int sntc_model_variable(sntc_model* m, int i, const char** name, int64_t shape[4], int* ndim) {
  if (!m || i < 0 || i >= (int)m->vars.size()) return fail(SNTC_E_INVALID, "sntc_model_variable: index out of range");
  if (name) *name = m->vars[i].name.c_str();
  if (ndim) *ndim = (int)m->vars[i].shape.size();
  if (shape) for (size_t d = 0; d < m->vars[i].shape.size() && d < 4; ++d) shape[d] = m->vars[i].shape[d];
  return SNTC_OK;
}

extern "C" int sntc_model_load_weights(sntc_model* m, const char* name, const float* host, const int64_t* shape, int ndim) {
  if (!m || !name || !host || !shape) return fail(SNTC_E_INVALID, "sntc_model_load_weights: bad argument");
  if (m->finalized) return fail(SNTC_E_STATE, "sntc_model_load_weights: model already finalized");
  for (auto& v : m->vars) {
    if (v.name != name) continue;
    if ((int)v.shape.size() != ndim) return fail(SNTC_E_INVALID, std::string("sntc_model_load_weights: rank mismatch for ") + name);
    size_t n = 1;
    for (int d = 0; d < ndim; ++d) {
      if (shape[d] != v.shape[d])
        return fail(SNTC_E_INVALID, std::string("sntc_model_load_weights: shape mismatch for ") + name + " (dim " +
                                      std::to_string(d) + ": got " + std::to_string(shape[d]) + ", expected " + std::to_string(v.shape[d]) + ")");
      n *= (size_t)shape[d];
    }
    for (size_t i = 0; i < n; ++i)
      if (!std::isfinite(host[i])) return fail(SNTC_E_INVALID, std::string("sntc_model_load_weights: non-finite value in ") + name);
    m->hw[name] = {std::vector<int64_t>(shape, shape + ndim), std::vector<float>(host, host + n)};
    return SNTC_OK;
  }
  return fail(SNTC_E_INVALID, std::string("sntc_model_load_weights: unknown variable ") + name);
}

static int upload(sntc_model* m, const void* host, size_t bytes, void** dptr) {
  CU_TRY(cudaMalloc(dptr, bytes ? bytes : 4));
  m->owned.push_back(*dptr);
  CU_TRY(cudaMemcpy(*dptr, host, bytes, cudaMemcpyHostToDevice));
  return SNTC_OK;
}

extern "C" int sntc_model_enable_vjp(sntc_model* m, int on) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_model_enable_vjp: model is NULL");
  if (m->finalized) return fail(SNTC_E_STATE, "sntc_model_enable_vjp: call it before sntc_model_finalize (the backward plans are packed from the host weights)");
  m->want_vjp = on != 0;
  return SNTC_OK;
}

// Builds the backward plan of one transform from the host weights (called by sntc_model_finalize).
static int vjp_build(sntc_model* m, const sntc_transform_desc& d, const char* prefix, VjpPlan& P) {
  try { P.fwd = build_transform(d, prefix); } catch (const std::exception& e) { return fail(SNTC_E_INVALID, std::string("vjp plan: ") + e.what()); }
  for (auto& op : P.fwd.ops)
    if (op.type == OP_STASH || op.type == OP_ACT_RES_D2S) { P.why = "TwoLayerResSynthesis(res_type='d2s') has no backward on this path"; return SNTC_OK; }
  for (auto& c : P.fwd.convs) {
    std::vector<float> b = pack_bias(c, m->hw), w = pack_band_weights(c, m->hw);
    TRY(upload(m, b.data(), b.size() * 4, (void**)&c.d_bias));
    TRY(upload(m, w.data(), w.size() * 4, (void**)&c.d_w));
  }
  P.d_gamma_t.assign(P.fwd.gdns.size(), nullptr);
  for (size_t gi = 0; gi < P.fwd.gdns.size(); ++gi) {
    GdnLayer& g = P.fwd.gdns[gi];
    const auto& beta = m->hw.at(g.beta).second;
    const auto& gamma = m->hw.at(g.gamma).second;
    g.Npad = (g.C + 3) / 4 * 4;
    std::vector<float> gp((size_t)g.C * g.Npad, 0.f), gt((size_t)g.C * g.Npad, 0.f);
    for (int i = 0; i < g.C; ++i) for (int j = 0; j < g.C; ++j) {
      gp[(size_t)i * g.Npad + j] = gamma[(size_t)i * g.C + j];
      gt[(size_t)j * g.Npad + i] = gamma[(size_t)i * g.C + j];
    }
    TRY(upload(m, beta.data(), beta.size() * 4, (void**)&g.d_beta));
    TRY(upload(m, gp.data(), gp.size() * 4, (void**)&g.d_gamma));
    TRY(upload(m, gt.data(), gt.size() * 4, (void**)&P.d_gamma_t[gi]));
  }
  P.bconv.clear();
  for (auto& c : P.fwd.convs) P.bconv.push_back(make_backward_conv(c));
  const bool tc = is_tc(m->desc.precision);
  P.btc.assign(P.bconv.size(), TcConv{});
  for (size_t i = 0; i < P.bconv.size(); ++i) {
    ConvLayer& bc = P.bconv[i];
    std::vector<float> b = pack_bias(bc, m->hw);
    TRY(upload(m, b.data(), b.size() * 4, (void**)&bc.d_bias));
    if (tc && tc_conv_supported(bc)) {
      std::string err;
      if (!m->ctx->tc.encode) return fail(SNTC_E_CUDA, "vjp plan (tensor-core path): cuTensorMapEncodeTiled unavailable");
      if (!tc_pack_conv(m->ctx->tc, bc, m->hw, P.btc[i], 0, m->owned, &err)) return fail(SNTC_E_CUDA, "vjp plan (tensor-core path): " + err);
    }
    if (!P.btc[i].ok) {
      std::vector<float> w = pack_band_weights(bc, m->hw);
      TRY(upload(m, w.data(), w.size() * 4, (void**)&bc.d_w));
    }
  }
  P.ok = true;
  return SNTC_OK;
}

extern "C" int sntc_model_finalize(sntc_model* m) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_model_finalize: model is NULL");
  if (m->finalized) return SNTC_OK;
  for (auto& v : m->vars)
    if (!m->hw.count(v.name)) return fail(SNTC_E_STATE, "sntc_model_finalize: missing variable " + v.name);
  CU_TRY(cudaSetDevice(m->ctx->device));
  for (Transform* t : {&m->hyper, &m->syn}) {
    for (auto& op : t->ops) {
      if (op.type == OP_CONVT || op.type == OP_CONVT_RGB) {
        ConvLayer& c = t->convs[op.conv];
        std::vector<float> b = pack_bias(c, m->hw);
        TRY(upload(m, b.data(), b.size() * 4, (void**)&c.d_bias));
        c.h_bias = b;
        if (op.type == OP_CONVT) {
          if (!c.merged) {   // merged final layers exist on the tensor-core path only
            std::vector<float> w = pack_band_weights(c, m->hw);
            TRY(upload(m, w.data(), w.size() * 4, (void**)&c.d_w));
          }
        } else {
          std::vector<float> w = pack_rgb_weights(c, m->hw);
          TRY(upload(m, w.data(), w.size() * 4, (void**)&c.d_w_rgb));
          if (c.k == 5 && c.s == 2 && c.cin == 12 && c.cout == 3 && c.p == 1) {
            c.h_w_tail.assign((size_t)25 * 12 * 3 + 3, 0.f);
            for (int tap = 0; tap < 25; ++tap) for (int ci = 0; ci < 12; ++ci) for (int co = 0; co < 3; ++co)
              c.h_w_tail[((size_t)tap * 12 + ci) * 3 + co] = w[((size_t)tap * c.cin_pad + ci) * 4 + co];
            for (int co = 0; co < 3; ++co) c.h_w_tail[(size_t)25 * 12 * 3 + co] = b[co];
          }
        }
      }
      if (op.gdn >= 0) {
        GdnLayer& g = t->gdns[op.gdn];
        const auto& beta = m->hw.at(g.beta).second;
        const auto& gamma = m->hw.at(g.gamma).second;
        g.Npad = (g.C + 3) / 4 * 4;
        std::vector<float> gp((size_t)g.C * g.Npad, 0.f);
        for (int i = 0; i < g.C; ++i) for (int j = 0; j < g.C; ++j) gp[(size_t)i * g.Npad + j] = gamma[(size_t)i * g.C + j];
        if (g.C <= 64) { g.h_beta = beta; g.h_gamma = gamma; }
        TRY(upload(m, beta.data(), beta.size() * 4, (void**)&g.d_beta));
        TRY(upload(m, gp.data(), gp.size() * 4, (void**)&g.d_gamma));
      }
    }
  }
  if (is_tc(m->desc.precision)) {
    std::string err;
    if (!tc_finalize(m->ctx->tc, m->tc, m->has_hyper ? &m->hyper : nullptr, m->has_syn ? &m->syn : nullptr, m->hw, m->owned, &err))
      return fail(SNTC_E_CUDA, "sntc_model_finalize (tensor-core path): " + err);
    // warp-MMA tail of a two-layer synthesis (sntc_kernels_tail_mma.cuh); SNTC_TAIL_MMA=0 keeps the FFMA tail
    if (m->has_syn && tc_env_int("SNTC_TAIL_MMA", 1)) {
      m->tail_mma.assign(m->syn.convs.size(), TailMma{});
      for (auto& op : m->syn.ops) {
        if (op.type != OP_CONVT_RGB || !tail_mma_supported(m->syn.convs[op.conv])) continue;
        if (!tail_mma_pack(m->syn.convs[op.conv], m->hw, m->tail_mma[op.conv], m->owned, &err))
          return fail(SNTC_E_CUDA, "sntc_model_finalize (warp-MMA tail): " + err);
      }
    }
    // window-GEMM tcgen05 tail (sntc_kernels_tail_tz.cuh) for hidden width 12; SNTC_TAIL_TZ=0 keeps the warp-MMA tail
    if (m->has_syn && tc_env_int("SNTC_TAIL_TZ", 1)) {
      m->tail_tz.assign(m->syn.convs.size(), TailTz{});
      for (auto& op : m->syn.ops) {
        if (op.type != OP_CONVT_RGB || !tail_tz_supported(m->syn.convs[op.conv])) continue;
        if (!tail_tz_pack(m->syn.convs[op.conv], m->hw, m->tail_tz[op.conv], m->owned, &err))
          return fail(SNTC_E_CUDA, "sntc_model_finalize (window-GEMM tail): " + err);
      }
    }
  }
  if (m->desc.prior == SNTC_PRIOR_DEEP_FACTORIZED) {
    const int Cz = m->hyper.in_channels;
    std::vector<float> pk((size_t)Cz * DF_STRIDE, 0.f);
    auto sp = [](float x) { return x > 20.f ? x : std::log1p(std::exp(x)); };
    const int off_m[4] = {0, 9, 24, 39}, off_b[4] = {3, 18, 33, 42}, off_f[3] = {6, 21, 36};
    const int f[5] = {1, 3, 3, 3, 1};
    for (int i = 0; i < 4; ++i) {
      const auto& M = m->hw.at("prior.matrix_" + std::to_string(i)).second;
      const auto& Bv = m->hw.at("prior.bias_" + std::to_string(i)).second;
      const int fo = f[i + 1], fi = f[i];
      for (int c = 0; c < Cz; ++c) {
        for (int a = 0; a < fo * fi; ++a) pk[(size_t)c * DF_STRIDE + off_m[i] + a] = sp(M[(size_t)c * fo * fi + a]);
        for (int a = 0; a < fo; ++a) pk[(size_t)c * DF_STRIDE + off_b[i] + a] = Bv[(size_t)c * fo + a];
        if (i < 3) {
          const auto& F = m->hw.at("prior.factor_" + std::to_string(i)).second;
          for (int a = 0; a < fo; ++a) pk[(size_t)c * DF_STRIDE + off_f[i] + a] = std::tanh(F[(size_t)c * fo + a]);
        }
      }
    }
    TRY(upload(m, pk.data(), pk.size() * 4, (void**)&m->d_prior));
  }
  for (auto& e : m->ev) CU_TRY(cudaEventCreate(&e));
  if (m->want_vjp) {
    if (m->has_hyper) TRY(vjp_build(m, m->desc.hyper, "hyper_synthesis", m->vjp_hyper));
    if (m->has_syn) TRY(vjp_build(m, m->desc.synthesis, "synthesis", m->vjp_syn));
  }
  m->hw.clear();
  m->finalized = true;
  return SNTC_OK;
}

// ------------------------------------------------------------------------------------------------
// tensor checks
static size_t tensor_elems(const sntc_tensor* t) {
  size_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= (size_t)t->shape[i];
  return n;
}

static int check_tensor(const sntc_tensor* t, const char* what, int code, int bits, int ndim) {
  if (!t) return fail(SNTC_E_INVALID, std::string(what) + ": tensor is NULL");
  if (!t->data && tensor_elems(t) != 0) return fail(SNTC_E_INVALID, std::string(what) + ": data is NULL");
  if (t->ndim != ndim || !t->shape) return fail(SNTC_E_INVALID, std::string(what) + ": expected rank " + std::to_string(ndim));
  if (t->dtype_lanes != 1 || t->dtype_code != code || t->dtype_bits != bits)
    return fail(SNTC_E_INVALID, std::string(what) + ": wrong dtype (code " + std::to_string(t->dtype_code) + ", bits " +
                                  std::to_string(t->dtype_bits) + ")");
  for (int i = 0; i < ndim; ++i)
    if (t->shape[i] < 0) return fail(SNTC_E_INVALID, std::string(what) + ": negative dimension");
  if (t->strides) {
    int64_t expect = 1;
    for (int i = ndim - 1; i >= 0; --i) {
      if (t->shape[i] != 1 && t->strides[i] != expect) return fail(SNTC_E_INVALID, std::string(what) + ": tensor must be dense NHWC");
      expect *= t->shape[i];
    }
  }
  if (t->device_type != SNTC_DL_CPU && t->device_type != SNTC_DL_CUDA && t->device_type != SNTC_DL_CUDA_HOST)
    return fail(SNTC_E_INVALID, std::string(what) + ": unsupported device type");
  return SNTC_OK;
}

static bool on_device(const sntc_tensor* t) { return t->device_type == SNTC_DL_CUDA; }
static void* tdata(const sntc_tensor* t) { return (char*)t->data + t->byte_offset; }

static int expect_shape(const sntc_tensor* t, const char* what, int64_t a, int64_t b, int64_t c, int64_t d) {
  if (t->shape[0] != a || t->shape[1] != b || t->shape[2] != c || t->shape[3] != d)
    return fail(SNTC_E_INVALID, std::string(what) + ": shape [" + std::to_string(t->shape[0]) + "," + std::to_string(t->shape[1]) + "," +
                                  std::to_string(t->shape[2]) + "," + std::to_string(t->shape[3]) + "] != expected [" + std::to_string(a) +
                                  "," + std::to_string(b) + "," + std::to_string(c) + "," + std::to_string(d) + "]");
  return SNTC_OK;
}

// Resolve an input tensor to a device pointer (staging host tensors).
static int stage_in(sntc_model* m, const sntc_tensor* t, size_t bytes, DevBuf& st, cudaStream_t s, const void** dptr) {
  if (on_device(t)) {
    if (t->device_id != m->ctx->device) return fail(SNTC_E_INVALID, "tensor lives on a different GPU than the context");
    *dptr = tdata(t);
    return SNTC_OK;
  }
  TRY(st.ensure(bytes));
  CU_TRY(cudaMemcpyAsync(st.p, tdata(t), bytes, cudaMemcpyHostToDevice, s));
  *dptr = st.p;
  return SNTC_OK;
}
// Resolve an output tensor to a device pointer (a staging buffer for host tensors; copied back by unstage_out).
static int stage_out(sntc_model* m, const sntc_tensor* t, size_t bytes, DevBuf& st, void** dptr) {
  if (on_device(t)) {
    if (t->device_id != m->ctx->device) return fail(SNTC_E_INVALID, "tensor lives on a different GPU than the context");
    *dptr = tdata(t);
    return SNTC_OK;
  }
  TRY(st.ensure(bytes));
  *dptr = st.p;
  return SNTC_OK;
}
static int unstage_out(const sntc_tensor* t, size_t bytes, const void* dptr, cudaStream_t s, bool* need_sync) {
  if (on_device(t)) return SNTC_OK;
  CU_TRY(cudaMemcpyAsync(tdata(t), dptr, bytes, cudaMemcpyDeviceToHost, s));
  *need_sync = true;
  return SNTC_OK;
}

// ------------------------------------------------------------------------------------------------
// per-layer profiling
static cudaEvent_t prof_event(sntc_model* m) {
  if (!m->prof_pool.empty()) { cudaEvent_t e = m->prof_pool.back(); m->prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
struct ProfScope {
  sntc_model* m; cudaStream_t s; int idx = -1;
  ProfScope(sntc_model* m_, cudaStream_t s_, const std::string& label, double macs) : m(m_), s(s_) {
    if (!m->prof_on) return;
    ProfRec r; r.label = label; r.macs = macs; r.a = prof_event(m); r.b = prof_event(m);
    cudaEventRecord(r.a, s);
    m->prof.push_back(r); idx = (int)m->prof.size() - 1; m->prof_agg_valid = false;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(m->prof[idx].b, s); }
};
static double conv_macs(const ConvLayer& c, int B, int h, int w) {
  return (double)B * h * w * c.k * c.k * (c.cin + (c.append_ones ? 1 : 0)) * c.cout;
}

extern "C" int sntc_profile_enable(sntc_model* m, int on) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_profile_enable: model is NULL");
  CU_TRY(cudaSetDevice(m->ctx->device));
  if (on) {
    CU_TRY(cudaStreamSynchronize(m->ctx->stream));
    for (auto& r : m->prof) { m->prof_pool.push_back(r.a); m->prof_pool.push_back(r.b); }
    m->prof.clear(); m->prof_agg.clear(); m->prof_agg_valid = false;
  }
  m->prof_on = on != 0;
  return SNTC_OK;
}
static int prof_aggregate(sntc_model* m) {
  if (m->prof_agg_valid) return SNTC_OK;
  CU_TRY(cudaSetDevice(m->ctx->device));
  m->prof_agg.clear();
  for (auto& r : m->prof) {
    CU_TRY(cudaEventSynchronize(r.b));
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
    ProfAgg* a = nullptr;
    for (auto& x : m->prof_agg) if (x.label == r.label) { a = &x; break; }
    if (!a) { m->prof_agg.push_back({r.label, 0.f, 0, r.macs}); a = &m->prof_agg.back(); }
    a->ms += ms; a->n += 1;
  }
  m->prof_agg_valid = true;
  return SNTC_OK;
}
extern "C" int sntc_profile_count(sntc_model* m) {
  if (!m || prof_aggregate(m) != SNTC_OK) return 0;
  return (int)m->prof_agg.size();
}
extern "C" int sntc_profile_get(sntc_model* m, int i, const char** label, float* total_ms, int* intervals, double* macs) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_profile_get: model is NULL");
  TRY(prof_aggregate(m));
  if (i < 0 || i >= (int)m->prof_agg.size()) return fail(SNTC_E_INVALID, "sntc_profile_get: index out of range");
  if (label) *label = m->prof_agg[i].label.c_str();
  if (total_ms) *total_ms = m->prof_agg[i].ms;
  if (intervals) *intervals = m->prof_agg[i].n;
  if (macs) *macs = m->prof_agg[i].macs;
  return SNTC_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 execution of a transform
struct FinalOut {
  float* full = nullptr;     // [B, hout, wout, C] f32 (transform-level call)
  uint8_t* u8 = nullptr;     // [B, H, W, C] cropped pixels
  float* crop = nullptr;     // [B, H, W, C] cropped floats
  int H = 0, W = 0;
};

static int launch_band_gemm(sntc_ctx* ctx, const BandGemmParams& P, cudaStream_t s) {
  int M = P.B * P.cnty * P.cntx;
  if (M <= 0 || P.N <= 0) return SNTC_OK;
  if (P.N >= 96) {
    dim3 grid((M + 127) / 128, (P.N + 127) / 128);
    band_gemm_f32_kernel<128, 128, 8, 8, 8><<<grid, 256, 0, s>>>(P);
  } else {
    dim3 grid((M + 63) / 64, (P.N + 31) / 32);
    band_gemm_f32_kernel<64, 32, 16, 4, 4><<<grid, 128, 0, s>>>(P);
  }
  ctx->launches++; ctx->kinds[SNTC_LAUNCH_BAND_F32]++;
  CU_TRY(cudaGetLastError());
  return SNTC_OK;
}

static int run_conv_f32(sntc_ctx* ctx, const ConvLayer& c, const float* in, int B, int h, int w, float* out,
                        const FinalOut* fin, cudaStream_t s) {
  for (size_t yi = 0, bi = 0; yi < c.by.size(); ++yi)
    for (size_t xi = 0; xi < c.bx.size(); ++xi, ++bi) {
      const Band& b = c.bands[bi];
      BandGemmParams P{};
      P.x = in; P.B = B; P.hin = h; P.win = w; P.cin = c.cin_pad; P.a_transform = A_NONE;
      P.w = c.d_w + b.w_off; P.K = b.K; P.N = b.N; P.Npad = b.Npad;
      P.bias = c.d_bias; P.cout = c.cout;
      P.s = c.s; P.p = c.p; P.phy0 = b.phy0; P.nphx = b.nphx; P.phx0 = b.phx0; P.Ty = b.Ty; P.Tx = b.Tx;
      cell_range(c.by[yi], c.s, c.p, h, &P.mloy, &P.cnty);
      cell_range(c.bx[xi], c.s, c.p, w, &P.mlox, &P.cntx);
      P.act = c.act;
      P.out = out; P.hout = h * c.s - c.out_crop; P.wout = w * c.s - c.out_crop; P.cstride = c.cout;
      if (fin) { P.out_u8 = fin->u8; P.out_crop = fin->crop; P.H = fin->H; P.W = fin->W; }
      P.gx = nullptr; P.gdn_mode = G_NONE;
      TRY(launch_band_gemm(ctx, P, s));
    }
  return SNTC_OK;
}

template <int K, int PD, int C1, int TR>
static int launch_tail(sntc_ctx* ctx, const ConvLayer& c, const float* in, int B, int h, int w, const FinalOut* fin, cudaStream_t s) {
  constexpr int T = (K + 1) / 2, RY = 8;
  constexpr int NY = ((PD + RY - 1) >> 1) - (PD >> 1) + T, NX = ((PD + 1) >> 1) - (PD >> 1) + T;
  constexpr int TILE_Y = (RY / 2) * (TR - 1) + NY, TILE_X = 31 + NX;
  TailParams Q{};
  Q.x = in; Q.B = B; Q.hin = h; Q.win = w; Q.w = c.d_w_rgb; Q.bias = c.d_bias;
  Q.hout = 2 * h; Q.wout = 2 * w;
  if (fin) { Q.out = fin->full; Q.out_u8 = fin->u8; Q.out_crop = fin->crop; Q.H = fin->H; Q.W = fin->W; }
  size_t smem = ((size_t)TILE_Y * TILE_X * C1 + (size_t)K * K * C1 * 4) * 4;
  static bool attr_done = false;
  if (!attr_done && smem > 48 * 1024) {
    CU_TRY(cudaFuncSetAttribute(tail_s2_kernel<K, PD, C1, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((Q.wout + 63) / 64, (Q.hout + RY * TR - 1) / (RY * TR), B);
  tail_s2_kernel<K, PD, C1, TR><<<grid, 32 * TR, smem, s>>>(Q);
  ctx->launches++; ctx->kinds[SNTC_LAUNCH_FINAL_F32]++;
  CU_TRY(cudaGetLastError());
  return SNTC_OK;
}

static int run_rgb_f32(sntc_ctx* ctx, const ConvLayer& c, const float* in, int B, int h, int w, const FinalOut* fin, cudaStream_t s) {
  if (c.s == 2 && c.k == 5 && c.p == 1 && c.cout == 3 && c.cin == c.cin_pad) {   // the two-layer tail (Keras k5 s2)
    if (c.cin == 12 && !c.h_w_tail.empty()) {
      TailParams Q{};
      Q.x = in; Q.B = B; Q.hin = h; Q.win = w; Q.w = c.d_w_rgb; Q.bias = c.d_bias; Q.hout = 2 * h; Q.wout = 2 * w;
      if (fin) { Q.out = fin->full; Q.out_u8 = fin->u8; Q.out_crop = fin->crop; Q.H = fin->H; Q.W = fin->W; }
      TailWeights<5, 12> Wt;
      memcpy(&Wt, c.h_w_tail.data(), sizeof(Wt));
      dim3 grid((Q.wout + 63) / 64, (Q.hout + 31) / 32, B);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = grid; cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr; cfg.numAttrs = tc_pdl(B) ? 1 : 0;
      CU_TRY(cudaLaunchKernelEx(&cfg, tail_s2_const_kernel<5, 1, 12, 4>, Q, Wt));
      ctx->launches++; ctx->kinds[SNTC_LAUNCH_FINAL_F32]++;
      CU_TRY(cudaGetLastError());
      return SNTC_OK;
    }
    if (c.cin == 12) return launch_tail<5, 1, 12, 4>(ctx, c, in, B, h, w, fin, s);
    if (c.cin == 24) return launch_tail<5, 1, 24, 2>(ctx, c, in, B, h, w, fin, s);
    if (c.cin == 48) return launch_tail<5, 1, 48, 2>(ctx, c, in, B, h, w, fin, s);
  }
  RgbCellParams P{};
  P.x = in; P.B = B; P.hin = h; P.win = w; P.cin = c.cin_pad; P.w = c.d_w_rgb; P.bias = c.d_bias;
  P.cout = c.cout; P.k = c.k; P.s = c.s; P.p = c.p;
  P.cnty = floor_div(c.s * h - 1 + c.p, c.s) + 1;
  P.cntx = floor_div(c.s * w - 1 + c.p, c.s) + 1;
  int cc = (40 * 1024) / (c.k * c.k * 16);
  cc = std::max(4, cc / 4 * 4);
  cc = std::min(cc, c.cin_pad);
  P.cc = cc;
  P.hout = h * c.s; P.wout = w * c.s;
  if (fin) { P.out = fin->full; P.out_u8 = fin->u8; P.out_crop = fin->crop; P.H = fin->H; P.W = fin->W; }
  size_t smem = (size_t)c.k * c.k * cc * 16;
  dim3 grid((P.cnty * P.cntx + 127) / 128, B);
  if (c.s == 1) convt_rgb_cell_kernel<1><<<grid, 128, smem, s>>>(P);
  else if (c.s == 2) convt_rgb_cell_kernel<2><<<grid, 128, smem, s>>>(P);
  else if (c.s == 4) convt_rgb_cell_kernel<4><<<grid, 128, smem, s>>>(P);
  else return fail(SNTC_E_UNSUPPORTED, "rgb cell kernel: stride must be 1, 2 or 4");
  ctx->launches++; ctx->kinds[SNTC_LAUNCH_FINAL_F32]++;
  CU_TRY(cudaGetLastError());
  return SNTC_OK;
}

// act_res_kernel keeps gamma [C][C] | beta [C] | a 128-pixel x tile in dynamic shared memory: above 48 KB (C = 64: 49 920 B)
// the launch needs the opt-in attribute
static int launch_act_res(sntc_ctx* ctx, const ActResParams& P, cudaStream_t s) {
  const size_t smem = ((size_t)P.C * P.C + P.C + 128 * (P.C + 1)) * 4;
  static size_t attr_bytes = 48 * 1024;
  if (smem > attr_bytes) {
    CU_TRY(cudaFuncSetAttribute(act_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes = smem;
  }
  act_res_kernel<<<(unsigned)((P.npix + 127) / 128), 128, smem, s>>>(P);
  ctx->launches++;
  CU_TRY(cudaGetLastError());
  return SNTC_OK;
}

static int run_gdn_f32(sntc_ctx* ctx, const GdnLayer& g, const float* in, size_t npix, float* out, cudaStream_t s) {
  if (g.C <= 64) {
    ActResParams P{};
    P.in = in; P.in_stride = g.C; P.out = out; P.npix = npix; P.C = g.C;
    P.act = g.inverse ? SNTC_ACT_IGDN1 : SNTC_ACT_GDN1; P.has_res = 0;
    P.beta = g.d_beta; P.gamma = g.d_gamma; P.gamma_stride = g.Npad; P.inverse = g.inverse;
    if (g.kind != GDN_1) return fail(SNTC_E_UNSUPPORTED, "classic GDN with C <= 64 is not on any decode path");
    return launch_act_res(ctx, P, s);
  }
  // norm = beta + f(x) @ gamma as a 1x1 band GEMM; epilogue multiplies / divides x
  BandGemmParams P{};
  P.x = in; P.B = 1; P.hin = 1; P.win = (int)npix; P.cin = g.C;
  P.a_transform = g.kind == GDN_1 ? A_ABS : A_SQUARE;
  P.w = g.d_gamma; P.K = g.C; P.N = g.C; P.Npad = g.Npad; P.bias = g.d_beta; P.cout = g.C;
  P.s = 1; P.p = 0; P.phy0 = 0; P.nphx = 1; P.phx0 = 0; P.Ty = 1; P.Tx = 1;
  P.mloy = 0; P.cnty = 1; P.mlox = 0; P.cntx = (int)npix;
  P.act = SNTC_ACT_NONE; P.out = out; P.hout = 1; P.wout = (int)npix; P.cstride = g.C;
  P.gx = in;
  P.gdn_mode = g.kind == GDN_1 ? (g.inverse ? G_MUL : G_DIV) : (g.inverse ? G_MUL_SQRT : G_DIV_SQRT);
  return launch_band_gemm(ctx, P, s);
}

// What the executor currently holds: an fp32 tensor and/or its fp16 hi/lo planes.
struct Cur { const float* f32 = nullptr; const __half* hi = nullptr; const __half* lo = nullptr; };

// Fusion of the entropy-model glue into the last hyper-synthesis GEMM (tensor-core path only).
struct HyperFuse {
  const void* q = nullptr; int q_kind = 0; int Cy = 0; float max_index = 63.f; bool trunc = false;
  float* y_hat = nullptr; uint8_t* idx = nullptr;
  bool no_planes = false;   // phase 1 of the two-phase decode: only fp32 mu + idx leave the epilogue
  bool want_rate = false; RateConst rc{}; double* rate_slots = nullptr; int* rate_slot_img = nullptr; size_t rate_nslots = 0;
  float* sigma_out = nullptr;   // out (want_rate, default): raw sigma for rate_y_flat_kernel; SNTC_RATE_FUSED=1 sums in the epilogue instead
  bool done = false; const __half* yh_hi = nullptr; const __half* yh_lo = nullptr;   // out: planes of y_hat
};

static bool op_on_tc(sntc_model* m, Transform& t, bool is_hyper, size_t i) {
  if (!is_tc(m->desc.precision)) return false;
  const Op& op = t.ops[i];
  if (op.type != OP_CONVT) return false;
  const std::vector<TcConv>& tc = is_hyper ? m->tc.hyper : m->tc.syn;
  return op.conv < (int)tc.size() && tc[op.conv].ok;
}

static bool gdn_on_tc(sntc_model* m, Transform& t, bool is_hyper, size_t i) {
  if (is_hyper || !is_tc(m->desc.precision) || i >= t.ops.size()) return false;
  const Op& op = t.ops[i];
  return op.type == OP_GDN && op.gdn < (int)m->tc.syn_gdn.size() && m->tc.syn_gdn[op.gdn].ok;
}

// tc_run_conv for a layer of a transform: a final layer in col2im form runs its 1x1 contraction with the overlap-add epilogue.
static int run_tc_layer(sntc_ctx* ctx, const ConvLayer& c, TcConv& tcv, const __half* hi, const __half* lo, int B, int h, int w, TcConvOut o,
                        cudaStream_t s, std::string* err) {
  if (!c.col2im) return tc_run_conv(ctx->tc, c, tcv, hi, lo, B, h, w, o, s, &ctx->launches, err);
  o.col2im = true; o.c2i_k = c.k; o.c2i_s = c.s; o.c2i_p = c.p; o.c2i_kp = c.c2i_kp; o.c2i_cout = c.cout; o.c2i_bias = c.d_bias;
  return tc_run_conv(ctx->tc, c.c2i[0], tcv, hi, lo, B, h, w, o, s, &ctx->launches, err);
}

// Runs `t` on `cur` [B,h,w,Cin].  Conv layers run on the tensor cores when the model was created with
// SNTC_PRECISION_TC_F16X3 (and the layer is TMA-addressable), otherwise on the fp32 CUDA-core kernels;
// pointwise stages and the tiny final conv always run on CUDA cores.  Intermediates ping-pong in the
// model workspace.  The final op writes to fin (pixel epilogue / fin->full) or, for the hyper transform
// with `hf`, straight into y_hat / idx.
static int run_transform(sntc_model* m, Transform& t, bool is_hyper, Cur cur, int B, int h, int w, const FinalOut* fin,
                         HyperFuse* hf, cudaStream_t s) {
  sntc_ctx* ctx = m->ctx;
  size_t max_f32 = 0, max_pl = 0;
  {
    int ch = h, cw = w;
    for (size_t i = 0; i < t.ops.size(); ++i) {
      const Op& op = t.ops[i];
      if (op.type == OP_CONVT || op.type == OP_CONVT_RGB) {
        const ConvLayer& c = t.convs[op.conv];
        if (c.append_ones) max_f32 = std::max(max_f32, (size_t)B * ch * cw * c.cin_pad * 4);
        if (op_on_tc(m, t, is_hyper, i)) max_pl = std::max(max_pl, (size_t)B * ch * cw * c.cin * 2);
        ch *= c.s; cw *= c.s;
        max_f32 = std::max(max_f32, (size_t)B * ch * cw * c.cout * 4);
        max_pl = std::max(max_pl, (size_t)B * ch * cw * std::max(c.cout, 32) * 2);   // >= the 16-padded planes of the tail input
      } else if (op.type == OP_STASH) {
        ch = h; cw = w;
      }
    }
  }
  TRY(m->ws_a.ensure(max_f32));
  TRY(m->ws_b.ensure(max_f32));
  const bool any_tc = is_tc(m->desc.precision);
  if (any_tc)
    for (auto& pb : m->tc.plane)
      if (!pb.ensure(max_pl)) return fail(SNTC_E_CUDA, "cudaMalloc failed for the fp16 activation planes");
  int ch = h, cw = w, cc = t.in_channels;
  const Cur cur0 = cur;                      // the transform's input (OP_STASH goes back to it)
  const float* stash = nullptr; int stash_c = 0;
  int flip = 0, pflip = 0;
  auto next_buf = [&]() { float* p = (float*)(flip ? m->ws_b.p : m->ws_a.p); flip ^= 1; return p; };
  auto next_planes = [&](__half** hi, __half** lo) { *hi = (__half*)m->tc.plane[pflip * 2].p; *lo = (__half*)m->tc.plane[pflip * 2 + 1].p; pflip ^= 1; };
  for (size_t i = 0; i < t.ops.size(); ++i) {
    Op& op = t.ops[i];
    const bool last = i + 1 == t.ops.size();
    if (op.type == OP_CONVT || op.type == OP_CONVT_RGB) {
      const ConvLayer& c = t.convs[op.conv];
      std::string lbl = c.sources[0].kernel.substr(0, c.sources[0].kernel.size() - 7);
      if (op_on_tc(m, t, is_hyper, i)) {
        // ---- tensor-core band GEMM ----
        bool split_here = false;
        if (!cur.hi) {   // first tensor-core layer of a chain: split the fp32 input into fp16 planes
          if (!cur.f32) return fail(SNTC_E_STATE, "executor: no input for the tensor-core layer");
          __half *hi, *lo;
          next_planes(&hi, &lo);
          size_t n8 = (size_t)B * ch * cw * c.cin / 8;
          ProfScope ps(m, s, lbl + ".split_input", 0);
          TRY(m->d_flag.ensure(4));
          CU_TRY(cudaMemsetAsync(m->d_flag.p, 0, 4, s));
          split_planes_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, s>>>(cur.f32, hi, lo, n8, (unsigned*)m->d_flag.p);
          ctx->launches++;
          CU_TRY(cudaGetLastError());
          cur.hi = hi; cur.lo = lo;
          split_here = true;
        }
        TcConv& tcv = (is_hyper ? m->tc.hyper : m->tc.syn)[op.conv];
        TcConvOut o;
        Cur nxt;
        bool skip_next = false;
        if (split_here) o.alo_flag = (const unsigned*)m->d_flag.p;   // integer-valued input (z_hat): the a_lo pass is skipped
        if (!is_hyper && m->desc.precision == SNTC_PRECISION_TC_F16X3_SYN2) o.pass_mask = 1u;   // a_lo*w_hi + a_hi*w_hi
        const bool next_tc = !last && op_on_tc(m, t, is_hyper, i + 1);
        if (last && hf) {
          o.hyper_final = true; o.q = hf->q; o.q_kind = hf->q_kind; o.Cy = hf->Cy; o.max_index = hf->max_index; o.trunc = hf->trunc;
          o.y_hat = hf->y_hat; o.idx = hf->idx;
          static const bool rate_fused = tc_env_int("SNTC_RATE_FUSED", 0) != 0;
          if (hf->want_rate && !rate_fused) {   // raw sigma leaves the epilogue (fp32), the bits are summed by rate_y_flat_kernel
            TRY(m->d_hs.ensure((size_t)B * ch * c.s * cw * c.s * hf->Cy * 4));
            o.sigma_out = (float*)m->d_hs.p; hf->sigma_out = o.sigma_out;
          } else if (hf->want_rate) {   // bits_y partials: one slot per (work item, CTA, epilogue warp)
            const size_t ns = tc_rate_slots(tcv, B, ch, cw, ctx->tc.num_sms);
            TRY(m->d_rate_slots.ensure(ns * 8));
            TRY(m->d_rate_img.ensure(ns * 4));
            CU_TRY(cudaMemsetAsync(m->d_rate_img.p, 0xFF, ns * 4, s));
            o.rate_slots = (double*)m->d_rate_slots.p; o.rate_slot_img = (int*)m->d_rate_img.p; o.rate_slot_cap = ns; o.rc = hf->rc;
            hf->rate_slots = o.rate_slots; hf->rate_slot_img = o.rate_slot_img; hf->rate_nslots = ns;
          }
          if (!hf->no_planes) {
            size_t pl = (size_t)B * ch * c.s * cw * c.s * hf->Cy * 2;
            if (!m->tc.yh[0].ensure(pl) || !m->tc.yh[1].ensure(pl)) return fail(SNTC_E_CUDA, "cudaMalloc failed for the y_hat planes");
            o.hi = (__half*)m->tc.yh[0].p; o.lo = (__half*)m->tc.yh[1].p;
          }
          hf->done = true; hf->yh_hi = o.hi; hf->yh_lo = o.lo;
        } else if (last) {
          if (fin) { o.f32 = fin->full; o.u8 = fin->u8; o.crop = fin->crop; o.H = fin->H; o.W = fin->W; }
        } else if (tcv.fused_two_layer) {
          // the pointwise stage that follows (IGDN / activation, + residual) is this GEMM's epilogue
          const Op& nx = t.ops[i + 1];
          o.two_layer = true;
          o.has_res = nx.type == OP_ACT_RES;
          o.C1 = o.has_res ? c.cout / 2 : c.cout;
          if (nx.gdn >= 0) {
            const GdnLayer& g = t.gdns[nx.gdn];
            o.gamma = g.d_gamma; o.gamma_stride = g.Npad; o.beta = g.d_beta; o.tl_inverse = g.inverse;
            o.h_gamma = g.h_gamma.data(); o.h_beta = g.h_beta.data();
            o.tl_act = g.inverse ? SNTC_ACT_IGDN1 : SNTC_ACT_GDN1;
          } else {
            o.tl_act = nx.act;
          }
          // the tail conv that follows runs on the tensor cores: hand it fp16 planes (channels padded to 16), no fp32 copy
          const bool tail_tc = !is_hyper && i + 2 < t.ops.size() && t.ops[i + 2].type == OP_CONVT_RGB &&
                               t.ops[i + 2].conv < (int)m->tc.syn_tail.size() && m->tc.syn_tail[t.ops[i + 2].conv].ok &&
                               m->tc.syn_tail[t.ops[i + 2].conv].CP == (o.C1 + 15) / 16 * 16;
          if (tail_tc) {
            __half *hi, *lo;
            next_planes(&hi, &lo);
            o.hi = hi; o.lo = lo; nxt.hi = hi; nxt.lo = lo;
          } else {
            float* dst = next_buf();
            o.f32 = dst; nxt.f32 = dst;
          }
          skip_next = true;
        } else if (next_tc) {
          __half *hi, *lo;
          next_planes(&hi, &lo);
          o.hi = hi; o.lo = lo; nxt.hi = hi; nxt.lo = lo;
        } else if (gdn_on_tc(m, t, is_hyper, i + 1)) {
          // a GDN stage on the tensor cores follows: x as fp32 (scaled by the norm there) + planes of the pooling input
          const GdnLayer& g = t.gdns[t.ops[i + 1].gdn];
          __half *hi, *lo;
          next_planes(&hi, &lo);
          o.hi = hi; o.lo = lo; o.plane_xform = g.kind == GDN_1 ? A_ABS : A_SQUARE;
          nxt.hi = hi; nxt.lo = lo;
          // GDN1: no fp32 copy of x -- the |x| planes carry sign(x) in the LSB of lo and the GDN epilogue rebuilds x from them
          // (SNTC_TC_GDN_SIGN=0: x as a separate fp32 tensor, as the classic (x^2) form always does)
          static const bool gdn_sign = tc_env_int("SNTC_TC_GDN_SIGN", 1) != 0;
          if (g.kind == GDN_1 && gdn_sign) {
            o.sign_in_lo = true;
          } else {
            float* dst = next_buf();
            o.f32 = dst; nxt.f32 = dst;
          }
        } else {
          float* dst = next_buf();
          o.f32 = dst; nxt.f32 = dst;
        }
        std::string err;
        // Experiment, OFF by default (SNTC_L2_CHUNK_MB > 0 enables it): a two-layer synthesis whose hidden tensor t does not fit the
        // L2 (24 x 768x512: 113 MB of fp32 against 126 MB) runs layer 1 and the window-GEMM tail over batch chunks, so the tail reads
        // what layer 1 just wrote from L2 instead of HBM.  Bit-identical (images are independent), but measured slower: 2 chunks
        // 0.807 -> 0.821 ms per step, 3 chunks 0.846 -- every extra launch costs ~8 us of prologue / fill / drain, more than the
        // L2 hits return.
        if (skip_next && o.f32 && !o.hi && fin && i + 3 == t.ops.size() && t.ops[i + 2].type == OP_CONVT_RGB &&
            t.ops[i + 2].conv < (int)m->tail_tz.size() && m->tail_tz[t.ops[i + 2].conv].ok) {
          static const int chunk_mb = tc_env_int("SNTC_L2_CHUNK_MB", 0);
          const size_t per_img = (size_t)ch * c.s * cw * c.s * o.C1 * 4;
          const int Bmax = chunk_mb > 0 ? (int)std::max<size_t>(1, ((size_t)chunk_mb << 20) / per_img) : B;
          const int nchunks = (B + Bmax - 1) / Bmax, Bc = (B + nchunks - 1) / nchunks;   // equal chunks
          if (Bc < B) {
            const ConvLayer& c2 = t.convs[t.ops[i + 2].conv];
            const std::string lbl2 = c2.sources[0].kernel.substr(0, c2.sources[0].kernel.size() - 7);
            const int h1 = ch * c.s, w1 = cw * c.s, h2 = h1 * c2.s, w2 = w1 * c2.s;
            for (int b0 = 0; b0 < B; b0 += Bc) {
              const int bc = std::min(Bc, B - b0);
              TcConvOut oc = o;
              oc.f32 = o.f32 + (size_t)b0 * h1 * w1 * o.C1;
              const size_t in_off = (size_t)b0 * ch * cw * c.cin;
              {
                ProfScope ps(m, s, lbl + "+activation", conv_macs(c, bc, ch, cw));
                if (tc_run_conv(ctx->tc, c, tcv, cur.hi + in_off, cur.lo + in_off, bc, ch, cw, oc, s, &ctx->launches, &err) != TC_OK)
                  return fail(SNTC_E_CUDA, "tensor-core path: " + err);
                ctx->kinds[SNTC_LAUNCH_BAND_TC]++;
              }
              TailTzOut to;
              to.f32 = fin->full ? fin->full + (size_t)b0 * h2 * w2 * c2.cout : nullptr;
              to.u8 = fin->u8 ? fin->u8 + (size_t)b0 * fin->H * fin->W * c2.cout : nullptr;
              to.crop = fin->crop ? fin->crop + (size_t)b0 * fin->H * fin->W * c2.cout : nullptr;
              to.H = fin->H; to.W = fin->W;
              ProfScope ps(m, s, lbl2, conv_macs(c2, bc, h1, w1));
              if (tail_tz_run(ctx->tc, c2, m->tail_tz[t.ops[i + 2].conv], oc.f32, bc, h1, w1, to, tc_pdl(B), s, &ctx->launches, &err,
                              m->desc.precision != SNTC_PRECISION_TC_F16X3_SYN2) != 0)
                return fail(SNTC_E_CUDA, "window-GEMM tail: " + err);
              ctx->kinds[SNTC_LAUNCH_TAIL_TC]++;
            }
            ch = h2; cw = w2; cc = c2.cout;
            cur = Cur{};
            i += 2;
            continue;
          }
        }
        ProfScope ps(m, s, skip_next ? lbl + "+activation" : lbl, conv_macs(c, B, ch, cw));
        if (run_tc_layer(ctx, c, tcv, cur.hi, cur.lo, B, ch, cw, o, s, &err) != TC_OK)
          return fail(SNTC_E_CUDA, "tensor-core path: " + err);
        ctx->kinds[SNTC_LAUNCH_BAND_TC]++;
        ch *= c.s; cw *= c.s; cc = c.cout;
        cur = nxt;
        if (skip_next) { cc = o.C1; ++i; }
        continue;
      }
      if (op.type == OP_CONVT_RGB && cur.hi && !is_hyper && op.conv < (int)m->tc.syn_tail.size() && m->tc.syn_tail[op.conv].ok) {
        // ---- tensor-core tail: stride-2 conv to <= 4 channels + crop + uint8 ----
        TailTc& tt = m->tc.syn_tail[op.conv];
        TailTcOut to;
        if (fin) { to.f32 = fin->full; to.u8 = fin->u8; to.crop = fin->crop; to.H = fin->H; to.W = fin->W; }
        std::string err;
        ProfScope ps(m, s, lbl, conv_macs(c, B, ch, cw));
        if (tail_tc_run(ctx->tc, c, tt, tt.bias, cur.hi, cur.lo, B, ch, cw, to, s, &ctx->launches, &err) != TC_OK)
          return fail(SNTC_E_CUDA, "tensor-core tail: " + err);
        ctx->kinds[SNTC_LAUNCH_TAIL_TC]++;
        ch *= c.s; cw *= c.s; cc = c.cout;
        cur = Cur{};
        continue;
      }
      if (op.type == OP_CONVT_RGB && cur.f32 && !is_hyper && op.conv < (int)m->tail_tz.size() && m->tail_tz[op.conv].ok) {
        // ---- window-GEMM tcgen05 tail: stride-2 conv 12 -> 3 channels + crop + uint8 ----
        TailTzOut to;
        if (fin) { to.f32 = fin->full; to.u8 = fin->u8; to.crop = fin->crop; to.H = fin->H; to.W = fin->W; }
        std::string err;
        ProfScope ps(m, s, lbl, conv_macs(c, B, ch, cw));
        if (tail_tz_run(ctx->tc, c, m->tail_tz[op.conv], cur.f32, B, ch, cw, to, tc_pdl(B), s, &ctx->launches, &err,
                        m->desc.precision != SNTC_PRECISION_TC_F16X3_SYN2) != 0)
          return fail(SNTC_E_CUDA, "window-GEMM tail: " + err);
        ctx->kinds[SNTC_LAUNCH_TAIL_TC]++;
        ch *= c.s; cw *= c.s; cc = c.cout;
        cur = Cur{};
        continue;
      }
      if (op.type == OP_CONVT_RGB && cur.f32 && !is_hyper && op.conv < (int)m->tail_mma.size() && m->tail_mma[op.conv].ok) {
        // ---- warp-MMA tail (split-fp16 mma.sync): stride-2 conv to 3 channels + crop + uint8 ----
        TailMmaOut to;
        if (fin) { to.f32 = fin->full; to.u8 = fin->u8; to.crop = fin->crop; to.H = fin->H; to.W = fin->W; }
        std::string err;
        ProfScope ps(m, s, lbl, conv_macs(c, B, ch, cw));
        if (tail_mma_run(ctx->tc, c, m->tail_mma[op.conv], cur.f32, B, ch, cw, to, tc_pdl(B), s, &ctx->launches, &err,
                         m->desc.precision != SNTC_PRECISION_TC_F16X3_SYN2) != 0)
          return fail(SNTC_E_CUDA, "warp-MMA tail: " + err);
        ctx->kinds[SNTC_LAUNCH_TAIL_MMA]++;
        ch *= c.s; cw *= c.s; cc = c.cout;
        cur = Cur{};
        continue;
      }
      // ---- fp32 CUDA-core kernels ----
      if (!cur.f32) return fail(SNTC_E_STATE, "executor: fp32 layer follows a tensor-core layer that produced planes only");
      if (c.append_ones) {
        float* tmp = next_buf();
        size_t n = (size_t)B * ch * cw * c.cin_pad;
        append_ones_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cur.f32, c.cin, tmp, c.cin_pad, (size_t)B * ch * cw);
        ctx->launches++;
        CU_TRY(cudaGetLastError());
        cur.f32 = tmp;
      }
      float* dst = last ? (fin ? fin->full : nullptr) : next_buf();
      ProfScope ps(m, s, lbl, conv_macs(c, B, ch, cw));
      if (op.type == OP_CONVT) TRY(run_conv_f32(ctx, c, cur.f32, B, ch, cw, dst, last ? fin : nullptr, s));
      else TRY(run_rgb_f32(ctx, c, cur.f32, B, ch, cw, fin, s));
      ch *= c.s; cw *= c.s; cc = c.cout;
      cur = Cur{}; cur.f32 = dst;
    } else if (op.type == OP_GDN) {
      const GdnLayer& gl = t.gdns[op.gdn];
      if (!cur.f32 && !(cur.hi && gdn_on_tc(m, t, is_hyper, i) && gl.kind == GDN_1))
        return fail(SNTC_E_STATE, "executor: GDN needs an fp32 input");
      if (cur.hi && gdn_on_tc(m, t, is_hyper, i)) {
        // ---- norm pool on the tensor cores: [pixels x C] (planes of |x| / x^2) * gamma, epilogue x * norm ----
        TcGdn& tg = m->tc.syn_gdn[op.gdn];
        TcConvOut o;
        Cur nxt;
        o.gx = cur.f32;
        if (!cur.f32) { o.gx_hi = cur.hi; o.gx_lo = cur.lo; }   // planes with sign(x) in the LSB of lo (see the conv above)
        o.gdn_mode = gl.kind == GDN_1 ? (gl.inverse ? G_MUL : G_DIV) : (gl.inverse ? G_MUL_SQRT : G_DIV_SQRT);
        if (!last && op_on_tc(m, t, is_hyper, i + 1)) {
          __half *hi, *lo;
          next_planes(&hi, &lo);
          o.hi = hi; o.lo = lo; nxt.hi = hi; nxt.lo = lo;
        } else {
          float* dst = next_buf();
          o.f32 = dst; nxt.f32 = dst;
        }
        std::string err;
        ProfScope ps(m, s, gl.beta.substr(0, gl.beta.size() - 5), (double)B * ch * cw * gl.C * gl.C);
        // a 1x1 "image" row of B*ch*cw pixels would overflow the TMA box arithmetic for large batches: keep [B,ch,cw]
        if (tc_run_conv(ctx->tc, tg.conv, tg.tc, cur.hi, cur.lo, B, ch, cw, o, s, &ctx->launches, &err) != TC_OK)
          return fail(SNTC_E_CUDA, "tensor-core GDN: " + err);
        ctx->kinds[SNTC_LAUNCH_BAND_TC]++;
        cur = nxt;
        continue;
      }
      float* dst = next_buf();
      ProfScope ps(m, s, gl.beta.substr(0, gl.beta.size() - 5), (double)B * ch * cw * gl.C * gl.C);
      TRY(run_gdn_f32(ctx, gl, cur.f32, (size_t)B * ch * cw, dst, s));
      cur = Cur{}; cur.f32 = dst;
    } else if (op.type == OP_STASH) {
      // residual branch of TwoLayerResSynthesis(res_type="d2s") is complete up to its last depth_to_space: keep it, restart from the input
      if (!cur.f32) return fail(SNTC_E_STATE, "executor: the d2s residual branch must end in an fp32 tensor");
      const size_t bytes = (size_t)B * ch * cw * cc * 4;
      TRY(m->ws_c.ensure(bytes));
      CU_TRY(cudaMemcpyAsync(m->ws_c.p, cur.f32, bytes, cudaMemcpyDeviceToDevice, s));
      stash = (const float*)m->ws_c.p; stash_c = cc;
      cur = cur0;   // the caller's fp32 tensor and / or its planes (the y_hat planes of a decode live in their own buffers, not in the ping-pong pair)
      ch = h; cw = w; cc = t.in_channels;
    } else if (op.type == OP_ACT_RES_D2S) {
      if (!cur.f32 || !stash) return fail(SNTC_E_STATE, "executor: d2s residual stage without its inputs");
      const int C = cc;
      if (stash_c != 4 * C || (ch & 1) || (cw & 1)) return fail(SNTC_E_STATE, "executor: d2s residual geometry");
      if (C > 64) return fail(SNTC_E_UNSUPPORTED, "two-layer hidden width > 64");
      float* dst = next_buf();
      ActResParams P{};
      P.in = cur.f32; P.in_stride = cc; P.out = dst; P.npix = (size_t)B * ch * cw; P.C = C; P.act = op.act; P.has_res = 2;
      P.res_ext = stash; P.H1 = ch; P.W1 = cw;
      if (op.gdn >= 0) { const GdnLayer& g = t.gdns[op.gdn]; P.beta = g.d_beta; P.gamma = g.d_gamma; P.gamma_stride = g.Npad; P.inverse = g.inverse; }
      ProfScope ps(m, s, "synthesis.activation+res_d2s", (double)P.npix * C * C);
      TRY(launch_act_res(ctx, P, s));
      cur = Cur{}; cur.f32 = dst;
    } else if (op.type == OP_ACT_RES) {
      if (!cur.f32) return fail(SNTC_E_STATE, "executor: activation stage needs an fp32 input");
      int C = cc / 2;
      float* dst = next_buf();
      ActResParams P{};
      P.in = cur.f32; P.in_stride = cc; P.out = dst; P.npix = (size_t)B * ch * cw; P.C = C; P.act = op.act; P.has_res = 1;
      if (op.gdn >= 0) { const GdnLayer& g = t.gdns[op.gdn]; P.beta = g.d_beta; P.gamma = g.d_gamma; P.gamma_stride = g.Npad; P.inverse = g.inverse; }
      if (C > 64) return fail(SNTC_E_UNSUPPORTED, "two-layer hidden width > 64");
      ProfScope ps(m, s, "synthesis.activation+res", (double)P.npix * C * C);
      TRY(launch_act_res(ctx, P, s));
      cur = Cur{}; cur.f32 = dst; cc = C;
    }
  }
  return SNTC_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int sntc_hyper_synthesis(sntc_model* m, const sntc_tensor* z_hat, sntc_tensor* out, void* stream) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_hyper_synthesis: model is NULL");
  if (!m->finalized) return fail(SNTC_E_STATE, "sntc_hyper_synthesis: model not finalized");
  if (!m->has_hyper) return fail(SNTC_E_STATE, "sntc_hyper_synthesis: model has no hyperprior");
  TRY(check_tensor(z_hat, "z_hat", SNTC_DL_FLOAT, 32, 4));
  TRY(check_tensor(out, "out", SNTC_DL_FLOAT, 32, 4));
  int B = (int)z_hat->shape[0], h = (int)z_hat->shape[1], w = (int)z_hat->shape[2];
  if (z_hat->shape[3] != m->hyper.in_channels) return fail(SNTC_E_INVALID, "z_hat: wrong channel count");
  TRY(expect_shape(out, "out", B, (int64_t)h * m->hyper.upsample, (int64_t)w * m->hyper.upsample, m->hyper.out_channels));
  if (tensor_elems(z_hat) == 0) return SNTC_OK;
  CU_TRY(cudaSetDevice(m->ctx->device));
  cudaStream_t s = pick_stream(m->ctx, stream);
  const void* dz; void* dout; bool need_sync = false;
  TRY(stage_in(m, z_hat, tensor_elems(z_hat) * 4, m->st_z, s, &dz));
  TRY(stage_out(m, out, tensor_elems(out) * 4, m->d_hs, &dout));
  FinalOut fin; fin.full = (float*)dout;
  { Cur c0; c0.f32 = (const float*)dz; TRY(run_transform(m, m->hyper, true, c0, B, h, w, &fin, nullptr, s)); }
  TRY(unstage_out(out, tensor_elems(out) * 4, dout, s, &need_sync));
  if (need_sync) CU_TRY(cudaStreamSynchronize(s));
  return SNTC_OK;
}

extern "C" int sntc_synthesis(sntc_model* m, const sntc_tensor* y_hat, sntc_tensor* out, void* stream) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_synthesis: model is NULL");
  if (!m->finalized) return fail(SNTC_E_STATE, "sntc_synthesis: model not finalized");
  if (!m->has_syn) return fail(SNTC_E_STATE, "sntc_synthesis: model has no synthesis transform");
  TRY(check_tensor(y_hat, "y_hat", SNTC_DL_FLOAT, 32, 4));
  TRY(check_tensor(out, "out", SNTC_DL_FLOAT, 32, 4));
  int B = (int)y_hat->shape[0], h = (int)y_hat->shape[1], w = (int)y_hat->shape[2];
  if (y_hat->shape[3] != m->syn.in_channels) return fail(SNTC_E_INVALID, "y_hat: wrong channel count");
  TRY(expect_shape(out, "out", B, (int64_t)h * m->syn.upsample, (int64_t)w * m->syn.upsample, m->syn.out_channels));
  if (tensor_elems(y_hat) == 0) return SNTC_OK;
  CU_TRY(cudaSetDevice(m->ctx->device));
  cudaStream_t s = pick_stream(m->ctx, stream);
  const void* dy; void* dout; bool need_sync = false;
  TRY(stage_in(m, y_hat, tensor_elems(y_hat) * 4, m->st_q, s, &dy));
  TRY(stage_out(m, out, tensor_elems(out) * 4, m->st_f32, &dout));
  FinalOut fin; fin.full = (float*)dout;
  { Cur c0; c0.f32 = (const float*)dy; TRY(run_transform(m, m->syn, false, c0, B, h, w, &fin, nullptr, s)); }
  TRY(unstage_out(out, tensor_elems(out) * 4, dout, s, &need_sync));
  if (need_sync) CU_TRY(cudaStreamSynchronize(s));
  return SNTC_OK;
}

// ------------------------------------------------------------------------------------------------
// decoder backward (f4; mshyper/models.py:401-408): grad_in = J(x)^T grad_out of one transform
static int launch_act_bwd(sntc_ctx* ctx, const ActBwdParams& P, cudaStream_t s) {
  if (P.npix == 0) return SNTC_OK;
  if (P.C <= 64) {
    const size_t smem = ((size_t)P.C * P.C + P.C + 2 * 128 * (P.C + 1)) * 4;
    static size_t attr_bytes = 48 * 1024;
    if (smem > attr_bytes) { CU_TRY(cudaFuncSetAttribute(act_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_bytes = smem; }
    act_bwd_kernel<<<(unsigned)((P.npix + 127) / 128), 128, smem, s>>>(P);
  } else {
    if (P.res_copy || !(P.act == SNTC_ACT_IGDN1 || P.act == SNTC_ACT_GDN1 || P.act == SNTC_ACT_IGDN_CLASSIC))
      return fail(SNTC_E_UNSUPPORTED, "vjp: pointwise stage wider than 64 channels that is not a GDN");
    const size_t smem = (size_t)2 * GB_PT * P.C * 4;
    if (smem > 200 * 1024) return fail(SNTC_E_UNSUPPORTED, "vjp: GDN wider than 800 channels");
    static size_t attr_bytes = 48 * 1024;
    if (smem > attr_bytes) { CU_TRY(cudaFuncSetAttribute(gdn_bwd_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_bytes = smem; }
    gdn_bwd_wide_kernel<<<(unsigned)((P.npix + GB_PT - 1) / GB_PT), 256, smem, s>>>(P);
  }
  ctx->launches++;
  CU_TRY(cudaGetLastError());
  return SNTC_OK;
}

static int vjp_run(sntc_model* m, VjpPlan& P, bool is_hyper, const float* x, const float* gout, float* gin, float* out, int B, int h, int w, cudaStream_t s) {
  sntc_ctx* ctx = m->ctx;
  Transform& t = P.fwd;
  // forward layers that the model itself runs on the tensor cores do so here too (same packed weights, fp32 result kept)
  Transform& mt = is_hyper ? m->hyper : m->syn;
  std::vector<TcConv>& ftc = is_hyper ? m->tc.hyper : m->tc.syn;
  const size_t nops = t.ops.size();
  if (P.acts.size() < nops) P.acts.resize(nops);
  std::vector<int> ih(nops), iw(nops), ic(nops);
  // ---- forward on the fp32 kernels, every op output kept ----
  const float* cur = x;
  int ch = h, cw = w, cc = t.in_channels;
  for (size_t i = 0; i < nops; ++i) {
    const Op& op = t.ops[i];
    ih[i] = ch; iw[i] = cw; ic[i] = cc;
    if (op.type == OP_CONVT) {
      const ConvLayer& c = t.convs[op.conv];
      const float* in = cur;
      if (c.append_ones) {
        const size_t n = (size_t)B * ch * cw * c.cin_pad;
        TRY(P.ones.ensure(n * 4));
        append_ones_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cur, c.cin, (float*)P.ones.p, c.cin_pad, (size_t)B * ch * cw);
        ctx->launches++;
        CU_TRY(cudaGetLastError());
        in = (const float*)P.ones.p;
      }
      TRY(P.acts[i].ensure((size_t)B * ch * c.s * cw * c.s * c.cout * 4));
      ProfScope ps(m, s, "vjp.forward." + c.sources[0].kernel.substr(0, c.sources[0].kernel.size() - 7), conv_macs(c, B, ch, cw));
      if (is_tc(m->desc.precision) && op.conv < (int)ftc.size() && ftc[op.conv].ok && !c.append_ones) {
        const size_t n = (size_t)B * ch * cw * c.cin;
        TRY(P.pl[0].ensure(n * 2)); TRY(P.pl[1].ensure(n * 2));
        split_planes_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, s>>>(in, (__half*)P.pl[0].p, (__half*)P.pl[1].p, n / 8, nullptr);
        ctx->launches++;
        CU_TRY(cudaGetLastError());
        TcConvOut o;
        o.f32 = (float*)P.acts[i].p;
        std::string err;
        if (run_tc_layer(ctx, mt.convs[op.conv], ftc[op.conv], (const __half*)P.pl[0].p, (const __half*)P.pl[1].p, B, ch, cw, o, s, &err) != TC_OK)
          return fail(SNTC_E_CUDA, "vjp forward (tensor-core path): " + err);
        ctx->kinds[SNTC_LAUNCH_BAND_TC]++;
      } else if (is_tc(m->desc.precision) && !is_hyper && op.conv < (int)m->tail_tz.size() && m->tail_tz[op.conv].ok) {
        TailTzOut to;                     // the two-layer tail (12 -> 3 channels): window-GEMM tcgen05 kernel, fp32 output
        to.f32 = (float*)P.acts[i].p;
        std::string err;
        if (tail_tz_run(ctx->tc, mt.convs[op.conv], m->tail_tz[op.conv], in, B, ch, cw, to, false, s, &ctx->launches, &err, true) != 0)
          return fail(SNTC_E_CUDA, "vjp forward (window-GEMM tail): " + err);
        ctx->kinds[SNTC_LAUNCH_TAIL_TC]++;
      } else {
        TRY(run_conv_f32(ctx, c, in, B, ch, cw, (float*)P.acts[i].p, nullptr, s));
      }
      ch *= c.s; cw *= c.s; cc = c.cout;
    } else if (op.type == OP_GDN) {
      const GdnLayer& g = t.gdns[op.gdn];
      TRY(P.acts[i].ensure((size_t)B * ch * cw * g.C * 4));
      if (is_tc(m->desc.precision) && !is_hyper && op.gdn < (int)m->tc.syn_gdn.size() && m->tc.syn_gdn[op.gdn].ok) {
        // the model's tensor-core norm pool: planes of |x| (x^2) -> [pixels x C] * gamma, epilogue x * norm
        TcGdn& tg = m->tc.syn_gdn[op.gdn];
        const size_t n = (size_t)B * ch * cw * g.C;
        TRY(P.pl[0].ensure(n * 2)); TRY(P.pl[1].ensure(n * 2));
        split_planes_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, s>>>(cur, (__half*)P.pl[0].p, (__half*)P.pl[1].p, n / 8, nullptr,
                                                                           g.kind == GDN_1 ? A_ABS : A_SQUARE);
        ctx->launches++;
        CU_TRY(cudaGetLastError());
        TcConvOut o;
        o.f32 = (float*)P.acts[i].p; o.gx = cur;
        o.gdn_mode = g.kind == GDN_1 ? (g.inverse ? G_MUL : G_DIV) : (g.inverse ? G_MUL_SQRT : G_DIV_SQRT);
        std::string err;
        ProfScope ps(m, s, "vjp.forward." + g.beta.substr(0, g.beta.size() - 5), (double)B * ch * cw * g.C * g.C);
        if (tc_run_conv(ctx->tc, tg.conv, tg.tc, (const __half*)P.pl[0].p, (const __half*)P.pl[1].p, B, ch, cw, o, s, &ctx->launches, &err) != TC_OK)
          return fail(SNTC_E_CUDA, "vjp forward (tensor-core GDN): " + err);
        ctx->kinds[SNTC_LAUNCH_BAND_TC]++;
      } else {
        ProfScope ps(m, s, "vjp.forward." + g.beta.substr(0, g.beta.size() - 5), (double)B * ch * cw * g.C * g.C);
        TRY(run_gdn_f32(ctx, g, cur, (size_t)B * ch * cw, (float*)P.acts[i].p, s));
      }
    } else if (op.type == OP_ACT_RES) {
      const int C = cc / 2;
      if (C > 64) return fail(SNTC_E_UNSUPPORTED, "two-layer hidden width > 64");
      TRY(P.acts[i].ensure((size_t)B * ch * cw * C * 4));
      ActResParams Q{};
      Q.in = cur; Q.in_stride = cc; Q.out = (float*)P.acts[i].p; Q.npix = (size_t)B * ch * cw; Q.C = C; Q.act = op.act; Q.has_res = 1;
      if (op.gdn >= 0) { const GdnLayer& g = t.gdns[op.gdn]; Q.beta = g.d_beta; Q.gamma = g.d_gamma; Q.gamma_stride = g.Npad; Q.inverse = g.inverse; }
      TRY(launch_act_res(ctx, Q, s));
      cc = C;
    } else {
      return fail(SNTC_E_UNSUPPORTED, "vjp: op without a backward");
    }
    cur = (const float*)P.acts[i].p;
  }
  if (out) CU_TRY(cudaMemcpyAsync(out, cur, (size_t)B * ch * cw * cc * 4, cudaMemcpyDeviceToDevice, s));
  // ---- backward ----
  const float* g = gout;
  int flip = 0;
  for (size_t ii = nops; ii-- > 0;) {
    const Op& op = t.ops[ii];
    const float* xin = ii > 0 ? (const float*)P.acts[ii - 1].p : x;
    const size_t in_bytes = (size_t)B * ih[ii] * iw[ii] * ic[ii] * 4;
    float* dst;
    if (ii == 0) dst = gin;
    else { TRY(P.g[flip].ensure(in_bytes)); dst = (float*)P.g[flip].p; flip ^= 1; }
    if (op.type == OP_CONVT) {
      const ConvLayer& c = t.convs[op.conv];
      ConvLayer& bc = P.bconv[op.conv];
      TcConv& btc = P.btc[op.conv];
      const int T = bc.k, hm = ih[ii] + T - 1, wm = iw[ii] + T - 1;
      const int Cp = btc.ok ? bc.cin : bc.cin_pad;
      S2dGradParams Q{};
      Q.g = g; Q.a = c.act != SNTC_ACT_NONE ? (const float*)P.acts[ii].p : nullptr; Q.act = c.act;
      Q.B = B; Q.H = ih[ii] * c.s; Q.W = iw[ii] * c.s; Q.Cf = c.cout; Q.s = c.s; Q.p = c.p; Q.hm = hm; Q.wm = wm; Q.Cpad = Cp;
      const size_t n = (size_t)B * hm * wm * Cp;
      if (btc.ok) {
        TRY(P.pl[0].ensure(n * 2)); TRY(P.pl[1].ensure(n * 2));
        Q.hi = (__half*)P.pl[0].p; Q.lo = (__half*)P.pl[1].p;
      } else {
        TRY(P.s2d.ensure(n * 4));
        Q.out = (float*)P.s2d.p;
      }
      const std::string lbl = "vjp.backward." + c.sources[0].kernel.substr(0, c.sources[0].kernel.size() - 7);
      {
        ProfScope ps(m, s, lbl + ".s2d", 0);
        s2d_grad_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(Q);
        ctx->launches++;
        CU_TRY(cudaGetLastError());
      }
      ProfScope ps(m, s, lbl, conv_macs(bc, B, hm, wm));
      if (btc.ok) {
        TcConvOut o;
        o.f32 = dst;
        std::string err;
        if (tc_run_conv(ctx->tc, bc, btc, Q.hi, Q.lo, B, hm, wm, o, s, &ctx->launches, &err) != TC_OK)
          return fail(SNTC_E_CUDA, "vjp (tensor-core path): " + err);
        ctx->kinds[SNTC_LAUNCH_BAND_TC]++;
      } else {
        TRY(run_conv_f32(ctx, bc, Q.out, B, hm, wm, dst, nullptr, s));
      }
    } else if (op.type == OP_GDN || op.type == OP_ACT_RES) {
      ActBwdParams Q{};
      const bool res = op.type == OP_ACT_RES;
      const int C = res ? ic[ii] / 2 : ic[ii];
      Q.x = xin; Q.x_stride = ic[ii]; Q.g = g; Q.g_stride = C; Q.out = dst; Q.out_stride = ic[ii];
      Q.npix = (size_t)B * ih[ii] * iw[ii]; Q.C = C; Q.res_copy = res ? 1 : 0; Q.act = op.act;
      if (op.gdn >= 0) {
        const GdnLayer& gl = t.gdns[op.gdn];
        Q.act = gl.kind == GDN_CLASSIC ? SNTC_ACT_IGDN_CLASSIC : (gl.inverse ? SNTC_ACT_IGDN1 : SNTC_ACT_GDN1);
        Q.inverse = gl.inverse ? 1 : 0; Q.classic = gl.kind == GDN_CLASSIC ? 1 : 0;
        Q.beta = gl.d_beta; Q.gamma = gl.d_gamma; Q.gamma_stride = gl.Npad; Q.gamma_t = P.d_gamma_t[op.gdn];
      }
      ProfScope ps(m, s, "vjp.backward.activation", (double)Q.npix * C * C * 2);
      TRY(launch_act_bwd(ctx, Q, s));
    }
    g = dst;
  }
  return SNTC_OK;
}

static int vjp_entry(sntc_model* m, bool is_hyper, const sntc_tensor* x, const sntc_tensor* grad_out, sntc_tensor* grad_in, sntc_tensor* out, void* stream) {
  const char* fn = is_hyper ? "sntc_hyper_synthesis_vjp" : "sntc_synthesis_vjp";
  if (!m) return fail(SNTC_E_INVALID, std::string(fn) + ": model is NULL");
  if (!m->finalized) return fail(SNTC_E_STATE, std::string(fn) + ": model not finalized");
  if (is_hyper ? !m->has_hyper : !m->has_syn) return fail(SNTC_E_STATE, std::string(fn) + ": model has no such transform");
  VjpPlan& P = is_hyper ? m->vjp_hyper : m->vjp_syn;
  if (!m->want_vjp) return fail(SNTC_E_STATE, std::string(fn) + ": call sntc_model_enable_vjp(model, 1) before sntc_model_finalize");
  if (!P.ok) return fail(SNTC_E_UNSUPPORTED, std::string(fn) + ": " + (P.why.empty() ? std::string("no backward plan") : P.why));
  TRY(check_tensor(x, "x", SNTC_DL_FLOAT, 32, 4));
  TRY(check_tensor(grad_out, "grad_out", SNTC_DL_FLOAT, 32, 4));
  TRY(check_tensor(grad_in, "grad_in", SNTC_DL_FLOAT, 32, 4));
  if (out) TRY(check_tensor(out, "out", SNTC_DL_FLOAT, 32, 4));
  const Transform& t = P.fwd;
  const int B = (int)x->shape[0], h = (int)x->shape[1], w = (int)x->shape[2];
  if (x->shape[3] != t.in_channels) return fail(SNTC_E_INVALID, "x: wrong channel count");
  TRY(expect_shape(grad_in, "grad_in", B, h, w, t.in_channels));
  TRY(expect_shape(grad_out, "grad_out", B, (int64_t)h * t.upsample, (int64_t)w * t.upsample, t.out_channels));
  if (out) TRY(expect_shape(out, "out", B, (int64_t)h * t.upsample, (int64_t)w * t.upsample, t.out_channels));
  if (tensor_elems(x) == 0) return SNTC_OK;
  CU_TRY(cudaSetDevice(m->ctx->device));
  cudaStream_t s = pick_stream(m->ctx, stream);
  const void *dx, *dg; void *dgin, *dout = nullptr; bool need_sync = false;
  TRY(stage_in(m, x, tensor_elems(x) * 4, P.st_x, s, &dx));
  TRY(stage_in(m, grad_out, tensor_elems(grad_out) * 4, P.st_g, s, &dg));
  TRY(stage_out(m, grad_in, tensor_elems(grad_in) * 4, P.st_gin, &dgin));
  if (out) TRY(stage_out(m, out, tensor_elems(out) * 4, P.st_out, &dout));
  TRY(vjp_run(m, P, is_hyper, (const float*)dx, (const float*)dg, (float*)dgin, (float*)dout, B, h, w, s));
  TRY(unstage_out(grad_in, tensor_elems(grad_in) * 4, dgin, s, &need_sync));
  if (out) TRY(unstage_out(out, tensor_elems(out) * 4, dout, s, &need_sync));
  if (need_sync) CU_TRY(cudaStreamSynchronize(s));
  return SNTC_OK;
}

extern "C" int sntc_synthesis_vjp(sntc_model* m, const sntc_tensor* y_hat, const sntc_tensor* grad_out, sntc_tensor* grad_in, sntc_tensor* out, void* stream) {
  return vjp_entry(m, false, y_hat, grad_out, grad_in, out, stream);
}
extern "C" int sntc_hyper_synthesis_vjp(sntc_model* m, const sntc_tensor* z_hat, const sntc_tensor* grad_out, sntc_tensor* grad_in, sntc_tensor* out, void* stream) {
  return vjp_entry(m, true, z_hat, grad_out, grad_in, out, stream);
}

static int decode_impl(sntc_model* m, const sntc_tensor* z_hat, const sntc_tensor* q_y, int H, int W,
                       sntc_tensor* out_u8, sntc_tensor* out_idx, sntc_tensor* out_yhat, sntc_tensor* out_f32,
                       const sntc_tensor* original_u8, sntc_image_metrics* metrics, sntc_image_rate* rate, void* stream) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_decode: model is NULL");
  if (rate && !m->has_hyper) return fail(SNTC_E_INVALID, "sntc_decode_rd: the rate term is implemented for the mean-scale hyperprior model only");
  if (!m->finalized) return fail(SNTC_E_STATE, "sntc_decode: model not finalized (weights missing?)");
  if (!m->has_syn) return fail(SNTC_E_STATE, "sntc_decode: model has no synthesis transform");
  if (!q_y) return fail(SNTC_E_INVALID, "sntc_decode: q_y is NULL");
  if (!q_y->shape || q_y->ndim != 4) return fail(SNTC_E_INVALID, "q_y: expected rank 4");
  int q_kind;
  if (q_y->dtype_code == SNTC_DL_FLOAT && q_y->dtype_bits == 32) q_kind = 0;
  else if (q_y->dtype_code == SNTC_DL_INT && q_y->dtype_bits == 16) q_kind = 1;
  else if (q_y->dtype_code == SNTC_DL_INT && q_y->dtype_bits == 8) q_kind = 2;
  else return fail(SNTC_E_INVALID, "q_y: dtype must be float32, int16 or int8");
  TRY(check_tensor(q_y, "q_y", q_y->dtype_code, q_y->dtype_bits, 4));
  const int q_bytes = q_kind == 0 ? 4 : (q_kind == 1 ? 2 : 1);
  const int B = (int)q_y->shape[0], hy = (int)q_y->shape[1], wy = (int)q_y->shape[2], Cy = (int)q_y->shape[3];
  if (Cy != m->syn.in_channels) return fail(SNTC_E_INVALID, "q_y: channel count does not match the synthesis transform");
  const int Hp = hy * m->syn.upsample, Wp = wy * m->syn.upsample, Co = m->syn.out_channels;
  if (H <= 0 || W <= 0 || H > Hp || W > Wp)
    return fail(SNTC_E_INVALID, "sntc_decode: image size " + std::to_string(H) + "x" + std::to_string(W) + " does not fit the latent grid (" +
                                  std::to_string(Hp) + "x" + std::to_string(Wp) + " padded)");
  int hz = 0, wz = 0;
  if (m->has_hyper) {
    TRY(check_tensor(z_hat, "z_hat", SNTC_DL_FLOAT, 32, 4));
    hz = (int)z_hat->shape[1]; wz = (int)z_hat->shape[2];
    if (z_hat->shape[0] != B || z_hat->shape[3] != m->hyper.in_channels || hz * m->hyper.upsample != hy || wz * m->hyper.upsample != wy)
      return fail(SNTC_E_INVALID, "z_hat: shape does not match q_y through the hyper-synthesis (x" + std::to_string(m->hyper.upsample) + ")");
  } else {
    if (z_hat) return fail(SNTC_E_INVALID, "sntc_decode: factorized model takes no z_hat");
    if (out_idx) return fail(SNTC_E_INVALID, "sntc_decode: factorized model has no scale indexes");
  }
  TRY(check_tensor(out_u8, "out_u8", SNTC_DL_UINT, 8, 4));
  TRY(expect_shape(out_u8, "out_u8", B, H, W, Co));
  if (out_idx) { TRY(check_tensor(out_idx, "out_idx", SNTC_DL_UINT, 8, 4)); TRY(expect_shape(out_idx, "out_idx", B, hy, wy, Cy)); }
  if (out_yhat) { TRY(check_tensor(out_yhat, "out_yhat", SNTC_DL_FLOAT, 32, 4)); TRY(expect_shape(out_yhat, "out_yhat", B, hy, wy, Cy)); }
  if (out_f32) { TRY(check_tensor(out_f32, "out_f32", SNTC_DL_FLOAT, 32, 4)); TRY(expect_shape(out_f32, "out_f32", B, H, W, Co)); }
  if ((original_u8 != nullptr) != (metrics != nullptr)) return fail(SNTC_E_INVALID, "sntc_decode: original_u8 and metrics go together");
  if (original_u8) { TRY(check_tensor(original_u8, "original_u8", SNTC_DL_UINT, 8, 4)); TRY(expect_shape(original_u8, "original_u8", B, H, W, Co)); }
  if (B == 0) return SNTC_OK;

  sntc_ctx* ctx = m->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = pick_stream(ctx, stream);
  bool need_sync = false;
  const size_t n_lat = (size_t)B * hy * wy * Cy, n_img = (size_t)B * H * W * Co;

  const void* d_q; const void* d_z = nullptr; void* d_u8; void* d_idx = nullptr; void* d_f32 = nullptr; const void* d_orig = nullptr;
  TRY(stage_in(m, q_y, n_lat * q_bytes, m->st_q, s, &d_q));
  if (m->has_hyper) TRY(stage_in(m, z_hat, tensor_elems(z_hat) * 4, m->st_z, s, &d_z));
  TRY(stage_out(m, out_u8, n_img, m->st_u8, &d_u8));
  if (out_idx) TRY(stage_out(m, out_idx, n_lat, m->st_idx, &d_idx));
  if (out_f32) TRY(stage_out(m, out_f32, n_img * 4, m->st_f32, &d_f32));
  if (original_u8) TRY(stage_in(m, original_u8, n_img, m->st_orig, s, &d_orig));
  // y_hat always materialises on the device (it is the synthesis input)
  void* d_yhat;
  if (out_yhat && on_device(out_yhat)) d_yhat = tdata(out_yhat);
  else { TRY(m->d_yhat.ensure(n_lat * 4)); d_yhat = m->d_yhat.p; }

  Cur ycur;
  // Everything that is enqueued on the stream for the decode itself; `capturing`: the stream is being captured into a CUDA graph
  // (no event records: events recorded inside a capture cannot be timed).
  auto enqueue = [&](bool capturing) -> int {
  if (!capturing) { CU_TRY(cudaEventRecord(m->ev[0], s)); m->ev_valid = true; }
  RateConst rc{};
  if (rate) {
    // Fixed configs for the ScaleIndexedEntropyModel, float32 like the reference computes them   mshyper/models.py:27-32
    const float smin = 0.11f, smax = 256.f;
    rc.max_index = (float)(m->desc.num_scales - 1);
    rc.log_scale_min = std::log(smin);
    rc.scale_factor = (std::log(smax) - std::log(smin)) / (float)(m->desc.num_scales - 1);
    TRY(m->d_rate.ensure((size_t)B * 16));
    CU_TRY(cudaMemsetAsync(m->d_rate.p, 0, (size_t)B * 16, s));
    if (m->h_rate_cap < B) {
      if (m->h_rate) cudaFreeHost(m->h_rate);
      CU_TRY(cudaHostAlloc((void**)&m->h_rate, (size_t)B * 16, cudaHostAllocDefault));
      m->h_rate_cap = B;
    }
    if (m->d_prior) {   // hyper_latent_bits under NoisyDeepFactorized   :249-252
      const size_t per = (size_t)hz * wz * m->hyper.in_channels;
      const int bpi = (int)std::min<size_t>((per + 256 * 4 - 1) / (256 * 4), 256);
      TRY(m->d_rate_zslots.ensure((size_t)B * bpi * 8));
      ProfScope ps(m, s, "rate.bits_z", 0);
      rate_z_kernel<<<dim3(bpi, B), 256, 0, s>>>((const float*)d_z, per, m->hyper.in_channels, m->d_prior, (double*)m->d_rate_zslots.p);
      rate_reduce_kernel<<<B, 256, 0, s>>>((const double*)m->d_rate_zslots.p, nullptr, 0, bpi, (double*)m->d_rate.p + 1, 2);
      ctx->launches += 2;
      CU_TRY(cudaGetLastError());
    }
  }
  if (m->has_hyper) {
    bool fused = false;
    const bool tc = is_tc(m->desc.precision);
    if (tc && op_on_tc(m, m->hyper, true, m->hyper.ops.size() - 1)) {
      // tensor-core path: split/exp/clamp/round and q + mu are the epilogue of the last hyper-synthesis GEMM
      HyperFuse hf;
      hf.q = d_q; hf.q_kind = q_kind; hf.Cy = Cy; hf.max_index = (float)(m->desc.num_scales - 1);
      hf.trunc = m->desc.index_rounding == SNTC_INDEX_TRUNC;
      const bool syn_tc = op_on_tc(m, m->syn, false, 0);
      hf.y_hat = (out_yhat || !syn_tc) ? (float*)d_yhat : nullptr;   // fp32 y_hat only if someone reads it
      hf.idx = (uint8_t*)d_idx;
      hf.want_rate = rate != nullptr; hf.rc = rc;
      Cur c0; c0.f32 = (const float*)d_z;
      TRY(run_transform(m, m->hyper, true, c0, B, hz, wz, nullptr, &hf, s));
      fused = hf.done;
      if (fused) { ycur.f32 = hf.y_hat; ycur.hi = hf.yh_hi; ycur.lo = hf.yh_lo; }
      if (fused && rate && hf.sigma_out) {
        const size_t per = (size_t)hy * wy * Cy;
        const int bpi = (int)std::min<size_t>((per / 4 + 256 * 4 - 1) / (256 * 4), 512);
        TRY(m->d_rate_slots.ensure((size_t)B * bpi * 8));
        ProfScope ps(m, s, "rate.bits_y", 0);
        rate_y_flat_kernel<<<dim3(bpi, B), 256, 0, s>>>(hf.sigma_out, d_q, q_kind, per, rc, (double*)m->d_rate_slots.p);
        rate_reduce_kernel<<<B, 256, 0, s>>>((const double*)m->d_rate_slots.p, nullptr, 0, bpi, (double*)m->d_rate.p, 2);
        ctx->launches += 2;
        CU_TRY(cudaGetLastError());
      } else if (fused && rate) {
        rate_reduce_kernel<<<B, 256, 0, s>>>(hf.rate_slots, hf.rate_slot_img, (int)hf.rate_nslots, 0, (double*)m->d_rate.p, 2);
        ctx->launches++;
        CU_TRY(cudaGetLastError());
      }
    }
    if (!fused) {
      TRY(m->d_hs.ensure(n_lat * 2 * 4));
      FinalOut fin; fin.full = (float*)m->d_hs.p;
      Cur c0; c0.f32 = (const float*)d_z;
      TRY(run_transform(m, m->hyper, true, c0, B, hz, wz, &fin, nullptr, s));
      if (!capturing) CU_TRY(cudaEventRecord(m->ev[1], s));
      ProfScope ps(m, s, "dequant_index", 0);
      DequantParams P{};
      P.hs = (const float*)m->d_hs.p; P.q = d_q; P.q_kind = q_kind; P.npix = (size_t)B * hy * wy; P.C = Cy;
      P.max_index = (float)(m->desc.num_scales - 1); P.trunc = m->desc.index_rounding == SNTC_INDEX_TRUNC;
      P.y_hat = (float*)d_yhat; P.idx = (uint8_t*)d_idx;
      size_t n = P.npix * (Cy / 4);
      dequant_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(P);
      ctx->launches++;
      CU_TRY(cudaGetLastError());
      ycur.f32 = (const float*)d_yhat;
      if (rate) {   // rate term a7 from the materialised raw sigma
        const size_t per = (size_t)hy * wy;
        const int bpi = (int)std::min<size_t>((per * Cy + 256 * 8 - 1) / (256 * 8), 512);
        TRY(m->d_rate_slots.ensure((size_t)B * bpi * 8));
        rate_y_kernel<<<dim3(bpi, B), 256, 0, s>>>((const float*)m->d_hs.p, d_q, q_kind, per, Cy, rc, (double*)m->d_rate_slots.p);
        rate_reduce_kernel<<<B, 256, 0, s>>>((const double*)m->d_rate_slots.p, nullptr, 0, bpi, (double*)m->d_rate.p, 2);
        ctx->launches += 2;
        CU_TRY(cudaGetLastError());
      }
    } else {
      if (!capturing) CU_TRY(cudaEventRecord(m->ev[1], s));
    }
  } else {
    if (!capturing) CU_TRY(cudaEventRecord(m->ev[1], s));
    if (q_kind == 0) {
      // y_hat = q, already float32; a caller-owned device out_yhat still has to receive it
      if (out_yhat && on_device(out_yhat) && d_yhat != d_q) CU_TRY(cudaMemcpyAsync(d_yhat, d_q, n_lat * 4, cudaMemcpyDeviceToDevice, s));
      d_yhat = const_cast<void*>(d_q);
    } else {
      convert_q_kernel<<<(unsigned)((n_lat / 4 + 255) / 256), 256, 0, s>>>(d_q, q_kind, (float*)d_yhat, n_lat / 4);
      ctx->launches++;
      CU_TRY(cudaGetLastError());
    }
    ycur.f32 = (const float*)d_yhat;
  }
  if (!capturing) CU_TRY(cudaEventRecord(m->ev[2], s));
  {
    FinalOut fin; fin.u8 = (uint8_t*)d_u8; fin.crop = (float*)d_f32; fin.H = H; fin.W = W;
    TRY(run_transform(m, m->syn, false, ycur, B, hy, wy, &fin, nullptr, s));
  }
  return SNTC_OK;
  };
  // ---- small batches are launch-bound (6 launches of 10-40 us each, ~60 us of host work per decode): replay the decode as ONE
  // CUDA graph.  Eligible: every tensor device-resident, no metrics / rate / profiling (nothing returns to the host), batch <=
  // SNTC_GRAPH_MAX_BATCH (4).  A (shapes, pointers, stream) key is run eagerly the first time (sizes the workspaces, uploads the
  // band tables), captured the second time, replayed from then on; any change of a pointer -- the caller's or a workspace's --
  // is a different key.  sntc_model_enable_graphs(m, 0) or SNTC_GRAPH=0 turns it off.
  bool graph_ok = m->graphs_on && tc_env_int("SNTC_GRAPH", 1) != 0 && B <= tc_env_int("SNTC_GRAPH_MAX_BATCH", 4) && !original_u8 && !rate && !m->prof_on &&
                  !tc_env_int("SNTC_TC_TRACE", 0) && on_device(q_y) && on_device(out_u8) && (!z_hat || on_device(z_hat)) && (!out_idx || on_device(out_idx)) &&
                  (!out_yhat || on_device(out_yhat)) && (!out_f32 || on_device(out_f32));
  bool ran = false;
  if (graph_ok) {
    GraphKey key{};
    key.v[0] = B; key.v[1] = H; key.v[2] = W; key.v[3] = hy; key.v[4] = wy; key.v[5] = q_kind;
    const void* ptrs[] = {d_q, d_z, d_u8, d_idx, d_f32, out_yhat ? tdata(out_yhat) : nullptr, (void*)s, m->ws_a.p, m->ws_b.p, m->d_yhat.p, m->d_hs.p, m->d_flag.p,
                          m->tc.plane[0].p, m->tc.plane[1].p, m->tc.plane[2].p, m->tc.plane[3].p, m->tc.yh[0].p, m->tc.yh[1].p};
    static_assert(sizeof(ptrs) / sizeof(ptrs[0]) <= sizeof(key.p) / sizeof(key.p[0]), "GraphKey too small");
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) key.p[i] = ptrs[i];
    GraphEntry* ge = nullptr;
    for (auto& e : m->graphs) if (memcmp(&e.key, &key, sizeof(key)) == 0) { ge = &e; break; }
    if (ge && ge->exec) {
      CU_TRY(cudaGraphLaunch(ge->exec, s));
      ctx->launches += ge->launches;
      for (int i = 0; i < SNTC_LAUNCH_KINDS; ++i) ctx->kinds[i] += ge->kinds[i];
      m->ev_valid = false;
      ran = true;
    } else if (ge && !ge->failed) {
      const uint64_t l0 = ctx->launches;
      uint64_t k0[SNTC_LAUNCH_KINDS];
      for (int i = 0; i < SNTC_LAUNCH_KINDS; ++i) k0[i] = ctx->kinds[i];
      cudaGraph_t graph = nullptr;
      cudaError_t ce = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
      int r = ce == cudaSuccess ? enqueue(true) : SNTC_E_CUDA;
      if (ce == cudaSuccess) ce = cudaStreamEndCapture(s, &graph);
      if (r == SNTC_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&ge->exec, graph, 0) == cudaSuccess) {
        ge->launches = ctx->launches - l0;
        for (int i = 0; i < SNTC_LAUNCH_KINDS; ++i) ge->kinds[i] = ctx->kinds[i] - k0[i];
        cudaGraphDestroy(graph);
        CU_TRY(cudaGraphLaunch(ge->exec, s));
        m->ev_valid = false;
        ran = true;
      } else {   // capture refused (a workspace grew, an unsupported call ...): this key stays on the eager path
        if (graph) cudaGraphDestroy(graph);
        ge->exec = nullptr; ge->failed = true;
        ctx->launches = l0;
        for (int i = 0; i < SNTC_LAUNCH_KINDS; ++i) ctx->kinds[i] = k0[i];
        cudaGetLastError();
      }
    }
    if (!ran) {
      TRY(enqueue(false));
      ran = true;
      if (!ge) {   // remember the key as it looks AFTER the eager run (workspaces sized)
        const void* ptrs2[] = {d_q, d_z, d_u8, d_idx, d_f32, out_yhat ? tdata(out_yhat) : nullptr, (void*)s, m->ws_a.p, m->ws_b.p, m->d_yhat.p, m->d_hs.p, m->d_flag.p,
                               m->tc.plane[0].p, m->tc.plane[1].p, m->tc.plane[2].p, m->tc.plane[3].p, m->tc.yh[0].p, m->tc.yh[1].p};
        for (size_t i = 0; i < sizeof(ptrs2) / sizeof(ptrs2[0]); ++i) key.p[i] = ptrs2[i];
        if (m->graphs.size() >= 16) { if (m->graphs.front().exec) cudaGraphExecDestroy(m->graphs.front().exec); m->graphs.erase(m->graphs.begin()); }
        GraphEntry ne; ne.key = key;
        m->graphs.push_back(ne);
      }
    }
  }
  if (!ran) TRY(enqueue(false));
  if (original_u8) {
    TRY(m->d_ssd.ensure((size_t)B * 8));
    if (m->h_ssd_cap < B) {
      if (m->h_ssd) cudaFreeHost(m->h_ssd);
      CU_TRY(cudaHostAlloc((void**)&m->h_ssd, (size_t)B * 8, cudaHostAllocDefault));
      m->h_ssd_cap = B;
    }
    CU_TRY(cudaMemsetAsync(m->d_ssd.p, 0, (size_t)B * 8, s));
    size_t per = (size_t)H * W * Co;
    dim3 grid((unsigned)std::min<size_t>((per + 256 * 16 - 1) / (256 * 16), 1024), B);
    ssd_kernel<<<grid, 256, 0, s>>>((const uint8_t*)d_orig, (const uint8_t*)d_u8, per, (unsigned long long*)m->d_ssd.p);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(m->h_ssd, m->d_ssd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    need_sync = true;
  }
  if (rate) {
    CU_TRY(cudaMemcpyAsync(m->h_rate, m->d_rate.p, (size_t)B * 16, cudaMemcpyDeviceToHost, s));
    need_sync = true;
  }
  CU_TRY(cudaEventRecord(m->ev[3], s));   // stage times are valid for eager decodes only (ev_valid)
  TRY(unstage_out(out_u8, n_img, d_u8, s, &need_sync));
  if (out_idx) TRY(unstage_out(out_idx, n_lat, d_idx, s, &need_sync));
  if (out_f32) TRY(unstage_out(out_f32, n_img * 4, d_f32, s, &need_sync));
  if (out_yhat && !on_device(out_yhat)) TRY(unstage_out(out_yhat, n_lat * 4, d_yhat, s, &need_sync));
  if (need_sync) CU_TRY(cudaStreamSynchronize(s));
  if (rate)
    for (int b = 0; b < B; ++b) { rate[b].bits_y = m->h_rate[2 * b]; rate[b].bits_z = m->h_rate[2 * b + 1]; }
  if (metrics) {
    double npx = (double)H * W * Co;
    for (int b = 0; b < B; ++b) {
      metrics[b].ssd = m->h_ssd[b];
      metrics[b].mse = (double)m->h_ssd[b] / npx;
      // psnr = -10 * (ln mse - 2 ln 255) / ln 10        image_utils.py:37
      metrics[b].psnr = -10.0 * (std::log(metrics[b].mse) - 2.0 * std::log(255.0)) / std::log(10.0);
    }
  }
  return SNTC_OK;
}

extern "C" int sntc_decode(sntc_model* m, const sntc_tensor* z_hat, const sntc_tensor* q_y, int H, int W,
                           sntc_tensor* out_u8, sntc_tensor* out_idx, sntc_tensor* out_yhat, sntc_tensor* out_f32,
                           const sntc_tensor* original_u8, sntc_image_metrics* metrics, void* stream) {
  return decode_impl(m, z_hat, q_y, H, W, out_u8, out_idx, out_yhat, out_f32, original_u8, metrics, nullptr, stream);
}

extern "C" int sntc_decode_rd(sntc_model* m, const sntc_tensor* z_hat, const sntc_tensor* q_y, int H, int W,
                              sntc_tensor* out_u8, sntc_tensor* out_idx, sntc_tensor* out_yhat, sntc_tensor* out_f32,
                              const sntc_tensor* original_u8, sntc_image_metrics* metrics, sntc_image_rate* rate, void* stream) {
  return decode_impl(m, z_hat, q_y, H, W, out_u8, out_idx, out_yhat, out_f32, original_u8, metrics, rate, stream);
}


// ------------------------------------------------------------------------------------------------
// Two-phase decode: q_y can only be range-decoded once the scale-table rows are known, so a real decoder runs
// hyper-synthesis first (-> idx to the host coder, mu stays on the device), then dequantises and synthesises.
extern "C" int sntc_decode_hyper(sntc_model* m, const sntc_tensor* z_hat, sntc_tensor* out_idx, void* stream) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_decode_hyper: model is NULL");
  if (!m->finalized) return fail(SNTC_E_STATE, "sntc_decode_hyper: model not finalized");
  if (!m->has_hyper || !m->has_syn) return fail(SNTC_E_STATE, "sntc_decode_hyper: needs a mean-scale hyperprior model");
  TRY(check_tensor(z_hat, "z_hat", SNTC_DL_FLOAT, 32, 4));
  const int B = (int)z_hat->shape[0], hz = (int)z_hat->shape[1], wz = (int)z_hat->shape[2];
  if (z_hat->shape[3] != m->hyper.in_channels) return fail(SNTC_E_INVALID, "z_hat: wrong channel count");
  const int hy = hz * m->hyper.upsample, wy = wz * m->hyper.upsample, Cy = m->syn.in_channels;
  TRY(check_tensor(out_idx, "out_idx", SNTC_DL_UINT, 8, 4));
  TRY(expect_shape(out_idx, "out_idx", B, hy, wy, Cy));
  m->ph_B = 0;
  if (B == 0) return SNTC_OK;
  sntc_ctx* ctx = m->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = pick_stream(ctx, stream);
  const size_t n_lat = (size_t)B * hy * wy * Cy;
  const void* d_z; void* d_idx; bool need_sync = false;
  TRY(stage_in(m, z_hat, tensor_elems(z_hat) * 4, m->st_z, s, &d_z));
  TRY(stage_out(m, out_idx, n_lat, m->st_idx, &d_idx));
  TRY(m->d_mu.ensure(n_lat * 4));
  bool fused = false;
  if (is_tc(m->desc.precision) && op_on_tc(m, m->hyper, true, m->hyper.ops.size() - 1)) {
    HyperFuse hf;
    hf.q = nullptr; hf.q_kind = 3; hf.Cy = Cy; hf.max_index = (float)(m->desc.num_scales - 1);
    hf.trunc = m->desc.index_rounding == SNTC_INDEX_TRUNC;
    hf.y_hat = (float*)m->d_mu.p; hf.idx = (uint8_t*)d_idx; hf.no_planes = true;
    Cur c0; c0.f32 = (const float*)d_z;
    TRY(run_transform(m, m->hyper, true, c0, B, hz, wz, nullptr, &hf, s));
    fused = hf.done;
  }
  if (!fused) {
    TRY(m->d_hs.ensure(n_lat * 2 * 4));
    FinalOut fin; fin.full = (float*)m->d_hs.p;
    Cur c0; c0.f32 = (const float*)d_z;
    TRY(run_transform(m, m->hyper, true, c0, B, hz, wz, &fin, nullptr, s));
    DequantParams P{};
    P.hs = (const float*)m->d_hs.p; P.q = nullptr; P.q_kind = 3; P.npix = (size_t)B * hy * wy; P.C = Cy;
    P.max_index = (float)(m->desc.num_scales - 1); P.trunc = m->desc.index_rounding == SNTC_INDEX_TRUNC;
    P.y_hat = (float*)m->d_mu.p; P.idx = (uint8_t*)d_idx;
    size_t n = P.npix * (Cy / 4);
    dequant_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(P);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
  }
  TRY(unstage_out(out_idx, n_lat, d_idx, s, &need_sync));
  if (need_sync) CU_TRY(cudaStreamSynchronize(s));
  m->ph_B = B; m->ph_hy = hy; m->ph_wy = wy;
  return SNTC_OK;
}

extern "C" int sntc_decode_latents(sntc_model* m, const sntc_tensor* q_y, int H, int W, sntc_tensor* out_u8, sntc_tensor* out_yhat,
                                   sntc_tensor* out_f32, const sntc_tensor* original_u8, sntc_image_metrics* metrics, void* stream) {
  if (!m) return fail(SNTC_E_INVALID, "sntc_decode_latents: model is NULL");
  if (!m->finalized) return fail(SNTC_E_STATE, "sntc_decode_latents: model not finalized");
  if (m->ph_B == 0) return fail(SNTC_E_STATE, "sntc_decode_latents: no sntc_decode_hyper call is pending");
  if (!q_y || !q_y->shape || q_y->ndim != 4) return fail(SNTC_E_INVALID, "q_y: expected rank 4");
  int q_kind;
  if (q_y->dtype_code == SNTC_DL_FLOAT && q_y->dtype_bits == 32) q_kind = 0;
  else if (q_y->dtype_code == SNTC_DL_INT && q_y->dtype_bits == 16) q_kind = 1;
  else if (q_y->dtype_code == SNTC_DL_INT && q_y->dtype_bits == 8) q_kind = 2;
  else return fail(SNTC_E_INVALID, "q_y: dtype must be float32, int16 or int8");
  TRY(check_tensor(q_y, "q_y", q_y->dtype_code, q_y->dtype_bits, 4));
  const int q_bytes = q_kind == 0 ? 4 : (q_kind == 1 ? 2 : 1);
  const int B = m->ph_B, hy = m->ph_hy, wy = m->ph_wy, Cy = m->syn.in_channels, Co = m->syn.out_channels;
  TRY(expect_shape(q_y, "q_y", B, hy, wy, Cy));
  const int Hp = hy * m->syn.upsample, Wp = wy * m->syn.upsample;
  if (H <= 0 || W <= 0 || H > Hp || W > Wp) return fail(SNTC_E_INVALID, "sntc_decode_latents: image size does not fit the latent grid");
  TRY(check_tensor(out_u8, "out_u8", SNTC_DL_UINT, 8, 4));
  TRY(expect_shape(out_u8, "out_u8", B, H, W, Co));
  if (out_yhat) { TRY(check_tensor(out_yhat, "out_yhat", SNTC_DL_FLOAT, 32, 4)); TRY(expect_shape(out_yhat, "out_yhat", B, hy, wy, Cy)); }
  if (out_f32) { TRY(check_tensor(out_f32, "out_f32", SNTC_DL_FLOAT, 32, 4)); TRY(expect_shape(out_f32, "out_f32", B, H, W, Co)); }
  if ((original_u8 != nullptr) != (metrics != nullptr)) return fail(SNTC_E_INVALID, "sntc_decode_latents: original_u8 and metrics go together");
  if (original_u8) { TRY(check_tensor(original_u8, "original_u8", SNTC_DL_UINT, 8, 4)); TRY(expect_shape(original_u8, "original_u8", B, H, W, Co)); }
  sntc_ctx* ctx = m->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = pick_stream(ctx, stream);
  bool need_sync = false;
  const size_t n_lat = (size_t)B * hy * wy * Cy, n_img = (size_t)B * H * W * Co;
  const void* d_q; void* d_u8; void* d_f32 = nullptr; const void* d_orig = nullptr;
  TRY(stage_in(m, q_y, n_lat * q_bytes, m->st_q, s, &d_q));
  TRY(stage_out(m, out_u8, n_img, m->st_u8, &d_u8));
  if (out_f32) TRY(stage_out(m, out_f32, n_img * 4, m->st_f32, &d_f32));
  if (original_u8) TRY(stage_in(m, original_u8, n_img, m->st_orig, s, &d_orig));
  const bool syn_tc = op_on_tc(m, m->syn, false, 0);
  float* d_yhat = nullptr;
  if (out_yhat && on_device(out_yhat)) d_yhat = (float*)tdata(out_yhat);
  else if (out_yhat || !syn_tc) { TRY(m->d_yhat.ensure(n_lat * 4)); d_yhat = (float*)m->d_yhat.p; }
  __half* hi = nullptr; __half* lo = nullptr;
  if (syn_tc) {
    if (!m->tc.yh[0].ensure(n_lat * 2) || !m->tc.yh[1].ensure(n_lat * 2)) return fail(SNTC_E_CUDA, "cudaMalloc failed for the y_hat planes");
    hi = (__half*)m->tc.yh[0].p; lo = (__half*)m->tc.yh[1].p;
  }
  {
    ProfScope ps(m, s, "dequant_planes", 0);
    const size_t n8 = (n_lat + 7) / 8;   // n_lat % 4 == 0 (model_create); the last thread may own a 4-element tail
    if (hi && n_lat % 8 != 0) return fail(SNTC_E_UNSUPPORTED, "sntc_decode_latents: the fp16 planes need a multiple of 8 latent elements");
    dequant_planes_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, s>>>((const float*)m->d_mu.p, d_q, q_kind, n_lat, d_yhat, hi, lo);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
  }
  Cur ycur; ycur.f32 = d_yhat; ycur.hi = hi; ycur.lo = lo;
  {
    FinalOut fin; fin.u8 = (uint8_t*)d_u8; fin.crop = (float*)d_f32; fin.H = H; fin.W = W;
    TRY(run_transform(m, m->syn, false, ycur, B, hy, wy, &fin, nullptr, s));
  }
  if (original_u8) {
    TRY(m->d_ssd.ensure((size_t)B * 8));
    if (m->h_ssd_cap < B) {
      if (m->h_ssd) cudaFreeHost(m->h_ssd);
      CU_TRY(cudaHostAlloc((void**)&m->h_ssd, (size_t)B * 8, cudaHostAllocDefault));
      m->h_ssd_cap = B;
    }
    CU_TRY(cudaMemsetAsync(m->d_ssd.p, 0, (size_t)B * 8, s));
    size_t per = (size_t)H * W * Co;
    dim3 grid((unsigned)std::min<size_t>((per + 256 * 16 - 1) / (256 * 16), 1024), B);
    ssd_kernel<<<grid, 256, 0, s>>>((const uint8_t*)d_orig, (const uint8_t*)d_u8, per, (unsigned long long*)m->d_ssd.p);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(m->h_ssd, m->d_ssd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    need_sync = true;
  }
  TRY(unstage_out(out_u8, n_img, d_u8, s, &need_sync));
  if (out_f32) TRY(unstage_out(out_f32, n_img * 4, d_f32, s, &need_sync));
  if (out_yhat && !on_device(out_yhat)) TRY(unstage_out(out_yhat, n_lat * 4, d_yhat, s, &need_sync));
  if (need_sync) CU_TRY(cudaStreamSynchronize(s));
  if (metrics) {
    double npx = (double)H * W * Co;
    for (int b = 0; b < B; ++b) {
      metrics[b].ssd = m->h_ssd[b];
      metrics[b].mse = (double)m->h_ssd[b] / npx;
      metrics[b].psnr = -10.0 * (std::log(metrics[b].mse) - 2.0 * std::log(255.0)) / std::log(10.0);
    }
  }
  return SNTC_OK;
}

// ------------------------------------------------------------------------------------------------
// host entropy coder (I/O stage; see sntc_coder.hpp)
struct sntc_coder { Coder c; };

extern "C" int sntc_coder_create(int num_scales, double scale_min, double scale_max, double tail_mass, int precision, sntc_coder** out) {
  if (!out) return fail(SNTC_E_INVALID, "sntc_coder_create: out is NULL");
  *out = nullptr;
  if (num_scales < 1 || num_scales > 256 || !(scale_min > 0) || !(scale_max >= scale_min) || !(tail_mass > 0 && tail_mass < 1) || precision < 8 || precision > 16)
    return fail(SNTC_E_INVALID, "sntc_coder_create: bad parameter");
  auto k = std::make_unique<sntc_coder>();
  k->c.precision = precision; k->c.tail_mass = tail_mass;
  // SCALE_FN(i) = exp(log(SCALE_MIN) + SCALE_FACTOR * i)   mshyper/models.py:27-32
  const double factor = num_scales > 1 ? (std::log(scale_max) - std::log(scale_min)) / (num_scales - 1.0) : 0.0;
  for (int i = 0; i < num_scales; ++i) {
    k->c.scale_rows.push_back(build_normal_table(std::exp(std::log(scale_min) + factor * i), tail_mass, precision));
    if (k->c.scale_rows.back().nsym + 1 > (1 << precision)) return fail(SNTC_E_INVALID, "sntc_coder_create: precision too small for the widest scale");
  }
  *out = k.release();
  return SNTC_OK;
}

extern "C" int sntc_coder_destroy(sntc_coder* k) { delete k; return SNTC_OK; }

extern "C" int sntc_coder_set_prior(sntc_coder* k, int Cz, const float* raw43) {
  if (!k || !raw43 || Cz <= 0) return fail(SNTC_E_INVALID, "sntc_coder_set_prior: bad argument");
  auto sp = [](double x) { return x > 30 ? x : std::log1p(std::exp(x)); };
  k->c.prior_rows.clear();
  for (int c = 0; c < Cz; ++c) {
    const float* p = raw43 + (size_t)c * 43;   // matrix_0[3] bias_0[3] factor_0[3] matrix_1[9] bias_1[3] factor_1[3] matrix_2[9] bias_2[3] factor_2[3] matrix_3[3] bias_3[1]
    DeepFactorizedChannel ch;
    for (int i = 0; i < 3; ++i) { ch.m0[i] = sp(p[i]); ch.b0[i] = p[3 + i]; ch.f0[i] = std::tanh((double)p[6 + i]); }
    for (int i = 0; i < 9; ++i) { ch.m1[i] = sp(p[9 + i]); ch.m2[i] = sp(p[24 + i]); }
    for (int i = 0; i < 3; ++i) { ch.b1[i] = p[18 + i]; ch.f1[i] = std::tanh((double)p[21 + i]); ch.b2[i] = p[33 + i]; ch.f2[i] = std::tanh((double)p[36 + i]); ch.m3[i] = sp(p[39 + i]); }
    ch.b3 = p[42];
    k->c.prior_rows.push_back(build_prior_table(ch, k->c.tail_mass, k->c.precision));
    if (k->c.prior_rows.back().nsym + 1 > (1 << k->c.precision)) return fail(SNTC_E_INVALID, "sntc_coder_set_prior: precision too small for the prior's support");
  }
  return SNTC_OK;
}

static const std::vector<CdfTable>* coder_tables(sntc_coder* k, int kind) {
  if (!k) return nullptr;
  return kind == 0 ? &k->c.scale_rows : (kind == 1 ? &k->c.prior_rows : nullptr);
}

extern "C" int sntc_coder_table(sntc_coder* k, int kind, int row, int32_t* offset, int32_t* nsym, const uint32_t** cdf) {
  const std::vector<CdfTable>* t = coder_tables(k, kind);
  if (!t || row < 0 || row >= (int)t->size()) return fail(SNTC_E_INVALID, "sntc_coder_table: no such table");
  if (offset) *offset = (*t)[row].offset;
  if (nsym) *nsym = (*t)[row].nsym;
  if (cdf) *cdf = (*t)[row].cdf.data();
  return SNTC_OK;
}

extern "C" int sntc_coder_encode(sntc_coder* k, int kind, const int32_t* symbols, const uint8_t* rows, size_t n, uint8_t** bytes, size_t* nbytes) {
  const std::vector<CdfTable>* t = coder_tables(k, kind);
  if (!t || t->empty() || !symbols || !bytes || !nbytes || (kind == 0 && !rows)) return fail(SNTC_E_INVALID, "sntc_coder_encode: bad argument");
  RangeEncoder enc;
  const size_t nrow = t->size();
  for (size_t i = 0; i < n; ++i) {
    const size_t r = kind == 0 ? rows[i] : i % nrow;
    if (r >= nrow) return fail(SNTC_E_INVALID, "sntc_coder_encode: table row out of range");
    encode_symbol(enc, (*t)[r], symbols[i], k->c.precision);
  }
  std::vector<uint8_t> out = enc.finish();
  *bytes = (uint8_t*)malloc(out.size() ? out.size() : 1);
  if (!*bytes) return fail(SNTC_E_INVALID, "sntc_coder_encode: out of memory");
  memcpy(*bytes, out.data(), out.size());
  *nbytes = out.size();
  return SNTC_OK;
}

extern "C" int sntc_coder_decode(sntc_coder* k, int kind, const uint8_t* bytes, size_t nbytes, const uint8_t* rows, size_t n, int32_t* symbols) {
  const std::vector<CdfTable>* t = coder_tables(k, kind);
  if (!t || t->empty() || !symbols || (!bytes && nbytes) || (kind == 0 && !rows)) return fail(SNTC_E_INVALID, "sntc_coder_decode: bad argument");
  RangeDecoder dec(bytes, nbytes);
  const size_t nrow = t->size();
  for (size_t i = 0; i < n; ++i) {
    const size_t r = kind == 0 ? rows[i] : i % nrow;
    if (r >= nrow) return fail(SNTC_E_INVALID, "sntc_coder_decode: table row out of range");
    symbols[i] = decode_symbol(dec, (*t)[r], k->c.precision);
  }
  if (dec.overrun()) return fail(SNTC_E_INVALID, "sntc_coder_decode: bitstream ended before all symbols were decoded (truncated or corrupt)");
  return SNTC_OK;
}

extern "C" void sntc_coder_free(void* p) { free(p); }

// ------------------------------------------------------------------------------------------------
// MS-SSIM of two uint8 image batches (validation metric of the evaluate loop)
extern "C" int sntc_image_msssim(sntc_ctx* ctx, const sntc_tensor* a_u8, const sntc_tensor* b_u8, double* msssim, void* stream) {
  if (!ctx) return fail(SNTC_E_INVALID, "sntc_image_msssim: ctx is NULL");
  if (!msssim) return fail(SNTC_E_INVALID, "sntc_image_msssim: output pointer is NULL");
  TRY(check_tensor(a_u8, "a_u8", SNTC_DL_UINT, 8, 4));
  TRY(check_tensor(b_u8, "b_u8", SNTC_DL_UINT, 8, 4));
  for (int i = 0; i < 4; ++i)
    if (a_u8->shape[i] != b_u8->shape[i]) return fail(SNTC_E_INVALID, "sntc_image_msssim: the two image batches differ in shape");
  const int B = (int)a_u8->shape[0], H = (int)a_u8->shape[1], W = (int)a_u8->shape[2], C = (int)a_u8->shape[3];
  if (B == 0) return SNTC_OK;
  if (C <= 0 || C > 65535 || B > 65535) return fail(SNTC_E_INVALID, "sntc_image_msssim: batch / channel count out of range");
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = pick_stream(ctx, stream);
  const size_t n = (size_t)B * H * W * C, n_al = (n + 255) / 256 * 256;
  const bool stage_a = !on_device(a_u8), stage_b = !on_device(b_u8);
  TRY(ctx->ms_ws.ensure(msssim_ws_bytes(B, H, W, C) + 2 * n_al));
  uint8_t* base = (uint8_t*)ctx->ms_ws.p;
  const uint8_t* da = (const uint8_t*)tdata(a_u8); const uint8_t* db = (const uint8_t*)tdata(b_u8);
  if (stage_a) { CU_TRY(cudaMemcpyAsync(base, da, n, cudaMemcpyHostToDevice, s)); da = base; }
  if (stage_b) { CU_TRY(cudaMemcpyAsync(base + n_al, db, n, cudaMemcpyHostToDevice, s)); db = base + n_al; }
  std::vector<double> stats((size_t)B * MS_SCALES * C * 2, 0.0);
  std::string err;
  const int rc = msssim_run(da, db, B, H, W, C, base + 2 * n_al, stats.data(), s, &ctx->launches, &err);
  if (rc == 1) return fail(SNTC_E_INVALID, err);
  if (rc != 0) return fail(SNTC_E_CUDA, err);
  msssim_combine(stats.data(), B, C, msssim_single_scale(H, W), msssim);
  return SNTC_OK;
}

// ------------------------------------------------------------------------------------------------
// The one collective of the path: all-reduce of a few host doubles over NCCL (see sntc.h).
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    if (!err.empty()) return false;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { const char* e = dlerror(); err = std::string("dlopen(libnccl.so.2) failed: ") + (e ? e : "?"); return false; }
    auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p && err.empty()) err = std::string("libnccl lacks ") + n; return p; };
    GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
    AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!err.empty()) { dlclose(lib); lib = nullptr; return false; }
    return true;
  }
};
static NcclApi g_nccl;

struct sntc_comm {
  sntc_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  double* d_buf = nullptr;   // device scratch
  double* h_buf = nullptr;   // pinned
};
static constexpr int SNTC_COMM_MAX_N = 4096;

#define NCCL_TRY(expr)                                                                                              \
  do {                                                                                                              \
    ncclResult_t r__ = (expr);                                                                                      \
    if (r__ != ncclSuccess) return fail(SNTC_E_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(r__));        \
  } while (0)

extern "C" int sntc_comm_unique_id(void* id128) {
  if (!id128) return fail(SNTC_E_INVALID, "sntc_comm_unique_id: id is NULL");
  if (!g_nccl.load()) return fail(SNTC_E_UNSUPPORTED, "sntc_comm_unique_id: " + g_nccl.err);
  static_assert(sizeof(ncclUniqueId) == SNTC_COMM_ID_BYTES, "NCCL unique id size");
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return SNTC_OK;
}

extern "C" int sntc_comm_create(sntc_ctx* ctx, const void* id128, int rank, int world, sntc_comm** out) {
  if (!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world) return fail(SNTC_E_INVALID, "sntc_comm_create: bad argument");
  *out = nullptr;
  if (!g_nccl.load()) return fail(SNTC_E_UNSUPPORTED, "sntc_comm_create: " + g_nccl.err);
  CU_TRY(cudaSetDevice(ctx->device));
  auto c = std::make_unique<sntc_comm>();
  c->ctx = ctx; c->rank = rank; c->world = world;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NCCL_TRY(g_nccl.CommInitRank(&c->comm, world, id, rank));
  CU_TRY(cudaMalloc((void**)&c->d_buf, SNTC_COMM_MAX_N * sizeof(double)));
  CU_TRY(cudaHostAlloc((void**)&c->h_buf, SNTC_COMM_MAX_N * sizeof(double), cudaHostAllocDefault));
  *out = c.release();
  return SNTC_OK;
}

extern "C" int sntc_comm_destroy(sntc_comm* c) {
  if (!c) return SNTC_OK;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  if (c->d_buf) cudaFree(c->d_buf);
  if (c->h_buf) cudaFreeHost(c->h_buf);
  delete c;
  return SNTC_OK;
}

extern "C" int sntc_comm_allreduce_f64(sntc_comm* c, double* values, int n, int op) {
  if (!c || !values || n < 0 || n > SNTC_COMM_MAX_N) return fail(SNTC_E_INVALID, "sntc_comm_allreduce_f64: bad argument (n <= 4096)");
  if (op != SNTC_REDUCE_SUM && op != SNTC_REDUCE_MAX) return fail(SNTC_E_INVALID, "sntc_comm_allreduce_f64: unknown op");
  if (n == 0) return SNTC_OK;
  CU_TRY(cudaSetDevice(c->ctx->device));
  cudaStream_t s = c->ctx->stream;
  memcpy(c->h_buf, values, (size_t)n * sizeof(double));
  CU_TRY(cudaMemcpyAsync(c->d_buf, c->h_buf, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  NCCL_TRY(g_nccl.AllReduce(c->d_buf, c->d_buf, (size_t)n, ncclFloat64, op == SNTC_REDUCE_SUM ? ncclSum : ncclMax, c->comm, s));
  CU_TRY(cudaMemcpyAsync(c->h_buf, c->d_buf, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  memcpy(values, c->h_buf, (size_t)n * sizeof(double));
  return SNTC_OK;
}

extern "C" int sntc_allreduce_metrics(sntc_comm* c, double sums[5]) { return sntc_comm_allreduce_f64(c, sums, 5, SNTC_REDUCE_SUM); }


// ------------------------------------------------------------------------------------------------
// LPIPS of two image batches (evaluate-loop metric; see sntc_kernels_lpips.cuh)
struct sntc_lpips {
  sntc_ctx* ctx = nullptr;
  int precision = SNTC_PRECISION_TC_F16X3;
  std::vector<ConvLayer> convs;          // 13 x ConvT(3, 1, p = 1) with the flipped VGG16 kernels, relu fused
  std::vector<TcConv> tc;                // tensor-core packing of conv_1 .. conv_12
  std::vector<int> block_last;           // index of the last conv of each VGG block (the tapped layers)
  HostWeights hw;
  float* d_lin[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  std::vector<void*> owned;
  bool finalized = false;
  DevBuf f_a, f_b, st_a, st_b, slots, d_out;
  TcDevBuf plane[4];
  double* h_out = nullptr; int h_out_cap = 0;
};

static const int LPIPS_BLOCKS[5][3] = {{64, 64, 0}, {128, 128, 0}, {256, 256, 256}, {512, 512, 512}, {512, 512, 512}};

extern "C" int sntc_lpips_create(sntc_ctx* ctx, int precision, sntc_lpips** out) {
  if (!ctx || !out) return fail(SNTC_E_INVALID, "sntc_lpips_create: bad argument");
  *out = nullptr;
  if (precision != SNTC_PRECISION_FP32 && precision != SNTC_PRECISION_TC_F16X3) return fail(SNTC_E_INVALID, "sntc_lpips_create: precision must be FP32 or TC_F16X3");
  auto lp = std::make_unique<sntc_lpips>();
  lp->ctx = ctx; lp->precision = precision;
  int cin = 3, i = 0;
  for (int b = 0; b < 5; ++b) {
    for (int j = 0; j < 3 && LPIPS_BLOCKS[b][j]; ++j, ++i) {
      const int cout = LPIPS_BLOCKS[b][j];
      lp->convs.push_back(make_conv("lpips", "conv_" + std::to_string(i), 3, 1, cin, cout, LAYOUT_TFC_IO, true, SNTC_ACT_RELU));
      cin = cout;
    }
    lp->block_last.push_back(i - 1);
  }
  *out = lp.release();
  return SNTC_OK;
}

extern "C" int sntc_lpips_destroy(sntc_lpips* lp) {
  if (!lp) return SNTC_OK;
  cudaSetDevice(lp->ctx->device);
  cudaStreamSynchronize(lp->ctx->stream);
  for (void* p : lp->owned) cudaFree(p);
  for (DevBuf* b : {&lp->f_a, &lp->f_b, &lp->st_a, &lp->st_b, &lp->slots, &lp->d_out}) b->release();
  for (auto& b : lp->plane) b.release();
  if (lp->h_out) cudaFreeHost(lp->h_out);
  delete lp;
  return SNTC_OK;
}

// Variables: lpips.conv_i.kernel [3,3,Cin,Cout] / lpips.conv_i.bias [Cout] (Keras Conv2D layouts, i = 0..12), lpips.lin_l.kernel [C_l] (l = 0..4)
extern "C" int sntc_lpips_load_weights(sntc_lpips* lp, const char* name, const float* host, const int64_t* shape, int ndim) {
  if (!lp || !name || !host || !shape) return fail(SNTC_E_INVALID, "sntc_lpips_load_weights: bad argument");
  if (lp->finalized) return fail(SNTC_E_STATE, "sntc_lpips_load_weights: already finalized");
  std::vector<int64_t> want;
  const std::string n(name);
  for (size_t i = 0; i < lp->convs.size(); ++i) {
    const ConvLayer& c = lp->convs[i];
    if (n == c.sources[0].kernel) want = {3, 3, c.cin, c.cout};
    if (n == c.sources[0].bias) want = {c.cout};
  }
  for (int l = 0; l < 5; ++l)
    if (n == "lpips.lin_" + std::to_string(l) + ".kernel") want = {lp->convs[lp->block_last[l]].cout};
  if (want.empty()) return fail(SNTC_E_INVALID, "sntc_lpips_load_weights: unknown variable " + n);
  if ((int)want.size() != ndim) return fail(SNTC_E_INVALID, "sntc_lpips_load_weights: rank mismatch for " + n);
  size_t cnt = 1;
  for (int d = 0; d < ndim; ++d) { if (shape[d] != want[d]) return fail(SNTC_E_INVALID, "sntc_lpips_load_weights: shape mismatch for " + n); cnt *= (size_t)shape[d]; }
  for (size_t i = 0; i < cnt; ++i) if (!std::isfinite(host[i])) return fail(SNTC_E_INVALID, "sntc_lpips_load_weights: non-finite value in " + n);
  std::vector<float> v(host, host + cnt);
  if (ndim == 4) {   // correlation kernel K[a] -> transposed-conv kernel W[a'] = K[2 - a'] (same [kh,kw,Cin,Cout] layout)
    const size_t plane = (size_t)shape[2] * shape[3];
    for (int ay = 0; ay < 3; ++ay) for (int ax = 0; ax < 3; ++ax)
      memcpy(&v[((size_t)ay * 3 + ax) * plane], host + ((size_t)(2 - ay) * 3 + (2 - ax)) * plane, plane * sizeof(float));
  }
  lp->hw[n] = {std::vector<int64_t>(shape, shape + ndim), std::move(v)};
  return SNTC_OK;
}

extern "C" int sntc_lpips_finalize(sntc_lpips* lp) {
  if (!lp) return fail(SNTC_E_INVALID, "sntc_lpips_finalize: NULL");
  if (lp->finalized) return SNTC_OK;
  CU_TRY(cudaSetDevice(lp->ctx->device));
  for (auto& c : lp->convs)
    for (const std::string& nm : {c.sources[0].kernel, c.sources[0].bias})
      if (!lp->hw.count(nm)) return fail(SNTC_E_STATE, "sntc_lpips_finalize: missing variable " + nm);
  auto up = [&](const void* h, size_t bytes, void** d) -> int {
    CU_TRY(cudaMalloc(d, bytes ? bytes : 4));
    lp->owned.push_back(*d);
    CU_TRY(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
    return SNTC_OK;
  };
  lp->tc.assign(lp->convs.size(), TcConv{});
  for (size_t i = 0; i < lp->convs.size(); ++i) {
    ConvLayer& c = lp->convs[i];
    std::vector<float> b = pack_bias(c, lp->hw);
    TRY(up(b.data(), b.size() * 4, (void**)&c.d_bias));
    c.h_bias = b;
    const bool on_tc = lp->precision == SNTC_PRECISION_TC_F16X3 && tc_conv_supported(c);
    if (on_tc) {
      std::string err;
      if (!lp->ctx->tc.encode) return fail(SNTC_E_CUDA, "sntc_lpips_finalize: cuTensorMapEncodeTiled unavailable");
      if (!tc_pack_conv(lp->ctx->tc, c, lp->hw, lp->tc[i], 0, lp->owned, &err)) return fail(SNTC_E_CUDA, "sntc_lpips_finalize: " + err);
    } else {
      std::vector<float> w = pack_band_weights(c, lp->hw);
      TRY(up(w.data(), w.size() * 4, (void**)&c.d_w));
    }
  }
  for (int l = 0; l < 5; ++l) {
    const std::string nm = "lpips.lin_" + std::to_string(l) + ".kernel";
    if (!lp->hw.count(nm)) return fail(SNTC_E_STATE, "sntc_lpips_finalize: missing variable " + nm);
    const auto& v = lp->hw.at(nm).second;
    TRY(up(v.data(), v.size() * 4, (void**)&lp->d_lin[l]));
  }
  if (lp->precision == SNTC_PRECISION_TC_F16X3) {
    cudaError_t e = cudaFuncSetAttribute(band_gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(band_gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(SNTC_E_CUDA, std::string("sntc_lpips_finalize: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  }
  lp->hw.clear();
  lp->finalized = true;
  return SNTC_OK;
}

extern "C" int sntc_image_lpips(sntc_lpips* lp, const sntc_tensor* a, const sntc_tensor* b, double* lpips, double* per_layer, void* stream) {
  if (!lp || !lpips) return fail(SNTC_E_INVALID, "sntc_image_lpips: bad argument");
  if (!lp->finalized) return fail(SNTC_E_STATE, "sntc_image_lpips: not finalized (weights missing?)");
  if (!a || !b) return fail(SNTC_E_INVALID, "sntc_image_lpips: tensor is NULL");
  const bool u8 = a->dtype_code == SNTC_DL_UINT && a->dtype_bits == 8;
  if (!u8 && !(a->dtype_code == SNTC_DL_FLOAT && a->dtype_bits == 32)) return fail(SNTC_E_INVALID, "sntc_image_lpips: images must be uint8 or float32 in [0, 255]");
  TRY(check_tensor(a, "a", a->dtype_code, a->dtype_bits, 4));
  TRY(check_tensor(b, "b", a->dtype_code, a->dtype_bits, 4));
  for (int i = 0; i < 4; ++i) if (a->shape[i] != b->shape[i]) return fail(SNTC_E_INVALID, "sntc_image_lpips: the two image batches differ in shape");
  const int B = (int)a->shape[0], H = (int)a->shape[1], W = (int)a->shape[2];
  if (a->shape[3] != 3) return fail(SNTC_E_INVALID, "sntc_image_lpips: images must have 3 channels");
  if (B == 0) return SNTC_OK;
  if (H < 16 || W < 16) return fail(SNTC_E_INVALID, "sntc_image_lpips: images must be at least 16 x 16 (five VGG16 blocks)");
  sntc_ctx* ctx = lp->ctx;
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = pick_stream(ctx, stream);
  const size_t px = (size_t)H * W, esz = u8 ? 1 : 4;
  // pairs per pass: the widest activation (block 1: 64 channels at full resolution) of both images stays below ~1 GiB
  const int nb_max = (int)std::max<size_t>(1, ((size_t)1 << 30) / (2 * px * 64 * 4));
  const int nbm = std::min(B, nb_max);
  const size_t n_max = 2 * (size_t)nbm;
  TRY(lp->f_a.ensure(n_max * px * 64 * 4));
  TRY(lp->f_b.ensure(n_max * px * 64 * 4));
  const bool tc = lp->precision == SNTC_PRECISION_TC_F16X3;
  if (tc) for (auto& pb : lp->plane) if (!pb.ensure(n_max * px * 64 * 2)) return fail(SNTC_E_CUDA, "sntc_image_lpips: cudaMalloc failed for the activation planes");
  const int spi = 256;
  TRY(lp->slots.ensure((size_t)nbm * spi * 8));
  TRY(lp->d_out.ensure((size_t)B * 6 * 8));
  CU_TRY(cudaMemsetAsync(lp->d_out.p, 0, (size_t)B * 6 * 8, s));
  const char* ha = (const char*)tdata(a); const char* hb = (const char*)tdata(b);
  LpipsPre pre{{0.458f, 0.448f, 0.450f}, {-0.030f, -0.088f, -0.188f}};   // lpips_tensorflow.py:18-19
  for (int b0 = 0; b0 < B; b0 += nbm) {
    const int nb = std::min(nbm, B - b0), n = 2 * nb;
    // stage / locate the two image chunks, preprocess them into one [n, H, W, 4] tensor (a first, then b)
    const void* da = ha + (size_t)b0 * px * 3 * esz; const void* db = hb + (size_t)b0 * px * 3 * esz;
    if (!on_device(a)) { TRY(lp->st_a.ensure((size_t)nbm * px * 3 * esz)); CU_TRY(cudaMemcpyAsync(lp->st_a.p, da, (size_t)nb * px * 3 * esz, cudaMemcpyHostToDevice, s)); da = lp->st_a.p; }
    if (!on_device(b)) { TRY(lp->st_b.ensure((size_t)nbm * px * 3 * esz)); CU_TRY(cudaMemcpyAsync(lp->st_b.p, db, (size_t)nb * px * 3 * esz, cudaMemcpyHostToDevice, s)); db = lp->st_b.p; }
    float* x0 = (float*)lp->f_a.p;
    const size_t np1 = (size_t)nb * px;
    lpips_preprocess_kernel<<<(unsigned)((np1 + 255) / 256), 256, 0, s>>>(da, u8 ? 1 : 0, np1, pre, (float4*)x0);
    lpips_preprocess_kernel<<<(unsigned)((np1 + 255) / 256), 256, 0, s>>>(db, u8 ? 1 : 0, np1, pre, (float4*)x0 + np1);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    int h = H, w = W, pflip = 0;
    float* cur_f32 = x0;                   // current activation as fp32 (nullptr when only planes exist)
    __half* cur_hi = nullptr; __half* cur_lo = nullptr;
    auto other_f32 = [&](float* p) { return p == (float*)lp->f_a.p ? (float*)lp->f_b.p : (float*)lp->f_a.p; };
    auto next_planes = [&](__half** hi, __half** lo) { *hi = (__half*)lp->plane[pflip * 2].p; *lo = (__half*)lp->plane[pflip * 2 + 1].p; pflip ^= 1; };
    size_t ci = 0;
    for (int blk = 0; blk < 5; ++blk) {
      for (; (int)ci <= lp->block_last[blk]; ++ci) {
        ConvLayer& c = lp->convs[ci];
        const bool tap = (int)ci == lp->block_last[blk];
        const bool this_tc = tc && lp->tc[ci].ok;
        const bool next_tc = tc && !tap && ci + 1 < lp->convs.size() && lp->tc[ci + 1].ok;
        if (this_tc) {
          if (!cur_hi) {   // fp32 -> fp16 hi/lo planes
            __half *hi, *lo;
            next_planes(&hi, &lo);
            const size_t n8 = (size_t)n * h * w * c.cin / 8;
            split_planes_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, s>>>(cur_f32, hi, lo, n8, nullptr);
            ctx->launches++;
            CU_TRY(cudaGetLastError());
            cur_hi = hi; cur_lo = lo;
          }
          TcConvOut o;
          float* dst = nullptr;
          __half *nhi = nullptr, *nlo = nullptr;
          if (tap || !next_tc) { dst = cur_f32 ? other_f32(cur_f32) : (float*)lp->f_a.p; o.f32 = dst; }
          if (next_tc) { next_planes(&nhi, &nlo); o.hi = nhi; o.lo = nlo; }
          std::string err;
          if (tc_run_conv(ctx->tc, c, lp->tc[ci], cur_hi, cur_lo, n, h, w, o, s, &ctx->launches, &err) != TC_OK) return fail(SNTC_E_CUDA, "sntc_image_lpips: " + err);
          ctx->kinds[SNTC_LAUNCH_BAND_TC]++;
          cur_f32 = dst; cur_hi = nhi; cur_lo = nlo;
        } else {
          if (!cur_f32) return fail(SNTC_E_STATE, "sntc_image_lpips: fp32 layer without an fp32 input");
          float* dst = other_f32(cur_f32);
          TRY(run_conv_f32(ctx, c, cur_f32, n, h, w, dst, nullptr, s));
          cur_f32 = dst; cur_hi = nullptr; cur_lo = nullptr;
        }
      }
      // head of this block: cur_f32 = features [n, h, w, C]
      const ConvLayer& lc = lp->convs[lp->block_last[blk]];
      const int npix = h * w;
      lpips_head_kernel<<<dim3(spi, nb), 256, 0, s>>>(cur_f32, nb, npix, lc.cout, lp->d_lin[blk], (double*)lp->slots.p, spi);
      lpips_layer_finalize_kernel<<<nb, 32, 0, s>>>((const double*)lp->slots.p, spi, 1.0 / (double)npix, (double*)lp->d_out.p + (size_t)B * (1 + blk) + b0);
      ctx->launches += 2;
      CU_TRY(cudaGetLastError());
      if (blk < 4) {   // MaxPooling2D(2, 2) into the next block's input
        const int ho = h / 2, wo = w / 2;
        const size_t tot = (size_t)n * ho * wo * (lc.cout / 8);
        const bool nxt_tc = tc && lp->tc[ci].ok;
        __half *hi = nullptr, *lo = nullptr;
        float* dst = nullptr;
        if (nxt_tc) next_planes(&hi, &lo); else dst = other_f32(cur_f32);
        lpips_maxpool_planes_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(cur_f32, n, h, w, lc.cout, hi, lo, dst);
        ctx->launches++;
        CU_TRY(cudaGetLastError());
        cur_f32 = dst; cur_hi = hi; cur_lo = lo;
        h = ho; w = wo;
      }
    }
  }
  if (lp->h_out_cap < B) {
    if (lp->h_out) cudaFreeHost(lp->h_out);
    CU_TRY(cudaHostAlloc((void**)&lp->h_out, (size_t)B * 6 * 8, cudaHostAllocDefault));
    lp->h_out_cap = B;
  }
  CU_TRY(cudaMemcpyAsync(lp->h_out, lp->d_out.p, (size_t)B * 6 * 8, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  for (int i = 0; i < B; ++i) {
    double t = 0.0;
    for (int l = 0; l < 5; ++l) { const double v = lp->h_out[(size_t)B * (1 + l) + i]; t += v; if (per_layer) per_layer[(size_t)i * 5 + l] = v; }
    lpips[i] = t;
  }
  return SNTC_OK;
}

extern "C" int sntc_last_stage_times_ms(sntc_model* m, float out[4]) {
  if (!m || !out) return fail(SNTC_E_INVALID, "sntc_last_stage_times_ms: bad argument");
  if (!m->ev_valid) return fail(SNTC_E_STATE, "sntc_last_stage_times_ms: no eagerly launched decode recorded (graph replays carry no stage events: sntc_model_enable_graphs(m, 0))");
  CU_TRY(cudaSetDevice(m->ctx->device));
  CU_TRY(cudaEventSynchronize(m->ev[3]));
  CU_TRY(cudaEventElapsedTime(&out[0], m->ev[0], m->ev[1]));
  CU_TRY(cudaEventElapsedTime(&out[1], m->ev[1], m->ev[2]));
  CU_TRY(cudaEventElapsedTime(&out[2], m->ev[2], m->ev[3]));
  CU_TRY(cudaEventElapsedTime(&out[3], m->ev[0], m->ev[3]));
  return SNTC_OK;
}
