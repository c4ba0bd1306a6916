/*
 * sntc.h -- C ABI of libsntc.so: the B200-native decode hot path of mandt-lab/shallow-ntc.
 *
 * The reference (pure Python / TensorFlow 2.10) has no FFI.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference repository root), i.e. what a
 * ctypes / cffi binding added to the reference would call instead of the TF op sequence.
 *
 * Conventions
 *   - every function returns an int status: SNTC_OK (0) or a negative SNTC_E_* code;
 *     sntc_last_error() returns a human-readable message for the calling thread.
 *   - tensors are described by `sntc_tensor`, which is layout-identical to DLPack's `DLTensor`
 *     (dlpack.h v0.8): the `dl_tensor` member of a DLManagedTensor capsule can be passed by
 *     pointer, zero-copy.  Tensors must be dense NHWC (strides NULL or contiguous).
 *   - device_type kDLCUDA (2): used in place on the context's device, asynchronously on `stream`.
 *     device_type kDLCPU (1) or kDLCUDAHost (3): staged through the context's device buffers
 *     (host<->device copies are part of the call; the call returns after the outputs landed).
 *   - the caller owns every input/output buffer; the context owns weights and workspace.
 *   - there is no CPU fallback: every entry point fails with SNTC_E_CUDA if no sm_100 GPU is usable.
 */
#ifndef SNTC_H_
#define SNTC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNTC_VERSION 102

/* status codes */
#define SNTC_OK 0
#define SNTC_E_INVALID (-1)   /* bad argument / shape / dtype */
#define SNTC_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define SNTC_E_STATE (-3)     /* call out of order (e.g. decode before finalize, missing weights) */
#define SNTC_E_UNSUPPORTED (-4)

/* DLPack device types / dtype codes (subset) */
#define SNTC_DL_CPU 1
#define SNTC_DL_CUDA 2
#define SNTC_DL_CUDA_HOST 3
#define SNTC_DL_INT 0
#define SNTC_DL_UINT 1
#define SNTC_DL_FLOAT 2

typedef struct sntc_tensor { /* == DLTensor */
  void* data;
  int32_t device_type;
  int32_t device_id;
  int32_t ndim;
  uint8_t dtype_code;
  uint8_t dtype_bits;
  uint16_t dtype_lanes;
  int64_t* shape;
  int64_t* strides; /* NULL = dense row-major */
  uint64_t byte_offset;
} sntc_tensor;

/* Transform classes of the reference's registry, common/transforms.py:383-393 (class_builder). */
enum sntc_transform_kind {
  SNTC_T_NONE = 0,                 /* no hyperprior: factorized/models.py */
  SNTC_T_HYPER_SYNTHESIS = 1,      /* HyperSynthesis            common/transforms.py:222-232 */
  SNTC_T_JPEG_LIKE_HYPER = 2,      /* JPEGLikeHyperSynthesis    common/transforms.py:364-377 */
  SNTC_T_HYPER_SMALL = 3,          /* HyperSynthesisSmall       common/transforms.py:250-262 */
  SNTC_T_JPEG_LIKE_SYNTHESIS = 10, /* JPEGLikeSynthesis         common/transforms.py:265-295 */
  SNTC_T_TWO_LAYER = 11,           /* TwoLayerSynthesis         common/transforms.py:298-317 */
  SNTC_T_TWO_LAYER_RES = 12,       /* TwoLayerResSynthesis      common/transforms.py:320-361 (res_type="conv") */
  SNTC_T_MBT2018 = 13,             /* MBT2018Synthesis          common/transforms.py:158-175 */
  SNTC_T_BLS2017 = 14,             /* BLS2017Synthesis          common/transforms.py:115-134 */
  SNTC_T_CNN = 15,                 /* CNNSynthesis              common/transforms.py:195-206 */
  SNTC_T_TWO_LAYER_RES_D2S = 16    /* TwoLayerResSynthesis(res_type="d2s")  common/transforms.py:339-348: the residual branch is
                                      depth_to_space(2), Conv2D 1x1 (-> 192) + leaky_relu, depth_to_space(2), Conv2D 1x1
                                      (-> 4 * C1) + leaky_relu, depth_to_space(2); variables res.conv_0 / res.conv_1
                                      (kernel [1,1,Cin,Cout], bias).  strides[0] must be 8 and in_channels % 16 == 0. */
};

/* get_activation_op, common/transforms.py:66-78 */
enum sntc_activation {
  SNTC_ACT_NONE = 0,
  SNTC_ACT_RELU = 1,
  SNTC_ACT_LEAKY_RELU = 2, /* tf.nn.leaky_relu, alpha 0.2 */
  SNTC_ACT_IGDN1 = 3,      /* GDN1(inverse=True)  */
  SNTC_ACT_GDN1 = 4,       /* GDN1()              */
  /* MBT2018Synthesis only (common/transforms.py:170 builds tfc.GDN(inverse=True) with the library defaults):
   * tensorflow-compression 2.x defaults are alpha_parameter=1, epsilon_parameter=1, i.e. the SAME function as
   * GDN1(inverse=True): y = x * (beta + |x| @ gamma) -- that is what activation = 0 selects for that class.
   * SNTC_ACT_IGDN_CLASSIC asks for the original (alpha=2, epsilon=.5) form y = x * sqrt(beta + x^2 @ gamma)
   * instead, for checkpoints trained with tfc.GDN(alpha_parameter=2, epsilon_parameter=.5). */
  SNTC_ACT_IGDN_CLASSIC = 5
};

/* Constructor kwargs of one transform class (same meaning as the Python kwargs). */
typedef struct sntc_transform_desc {
  int32_t kind;            /* sntc_transform_kind */
  int32_t in_channels;     /* channels of the latent fed to the transform (Cy, or Cz for hyper) */
  int32_t channels[2];     /* two-layer: channels=(C1, Cout); mbt/cnn: (channels_base, output_channels);
                              bls: (num_filters, 3); jpeg-like: (output_channels, 0);
                              hyper: (bottleneck_size, 0) */
  int32_t kernel_sizes[2]; /* two-layer: kernel_sizes; jpeg-like (hyper): (kernel_size, 0) */
  int32_t strides[2];      /* two-layer: strides; jpeg-like: (strides, 0) */
  int32_t activation;      /* sntc_activation (activation_type) */
  int32_t n_layers;        /* MBT2018Synthesis n_layers (default 4) */
  int32_t use_bias;        /* JPEGLikeSynthesis use_bias */
  int32_t use_offset;      /* JPEGLikeSynthesis use_offset */
} sntc_transform_desc;

/* precision of the contraction kernels */
#define SNTC_PRECISION_FP32 0      /* CUDA-core FFMA, fp32 throughout */
#define SNTC_PRECISION_TC_F16X3 1  /* tcgen05 split-fp16 3-pass, fp32 accumulate in TMEM (fp32-class accuracy) */
/* Opt-in, NOT the parity mode: hyper-synthesis as in TC_F16X3 (mu, idx, y_hat and the rate are bit-identical to it); the
 * synthesis layers drop the a_hi * w_lo cross term (2 MMA passes, tail included).  Reconstruction error grows from ~1e-5 to
 * ~1e-4 (inside the 1e-3 tolerance) and a few percent of the uint8 samples move by one LSB. */
#define SNTC_PRECISION_TC_F16X3_SYN2 2

/* prior of the hyper-latent z (mshyper/models.py:135): tfc.NoisyDeepFactorized(batch_shape=(Cz,)), num_filters (3,3,3).
 * With SNTC_PRIOR_DEEP_FACTORIZED the model expects the raw tfc variables prior.matrix_{0..3} [Cz,f_out,f_in],
 * prior.bias_{0..3} [Cz,f_out,1], prior.factor_{0..2} [Cz,f_out,1] (softplus / tanh are applied when packing). */
#define SNTC_PRIOR_NONE 0
#define SNTC_PRIOR_DEEP_FACTORIZED 1

/* Rule that turns the clamped float index i_c = clamp(exp(raw_sigma), 0, S-1) into the scale-table row a range coder
 * uses.  tensorflow-compression 2.10 does this in ContinuousIndexedEntropyModel._flatten_indexes with
 * tf.cast(indexes, tf.int32), i.e. TRUNCATION (= floor, i_c >= 0): SNTC_INDEX_TRUNC is what the host side
 * (shallow_ntc_b200.models.Model) selects by default.  SNTC_INDEX_RINT (round half to even) is kept for coders that
 * round.  The eval-path rate term (bits_y) uses the continuous i_c under either rule. */
#define SNTC_INDEX_RINT 0
#define SNTC_INDEX_TRUNC 1

/* Everything Model.__init__ / _init_transforms (mshyper/models.py:46-149, factorized/models.py:51-68)
 * fixes about the decode path. */
typedef struct sntc_model_desc {
  int32_t struct_size;            /* sizeof(sntc_model_desc), for ABI evolution */
  sntc_transform_desc hyper;      /* kind SNTC_T_NONE for the factorized model */
  sntc_transform_desc synthesis;
  int32_t num_scales;             /* NUM_SCALES, mshyper/models.py:28 (64) */
  int32_t index_rounding;         /* SNTC_INDEX_* */
  int32_t precision;              /* SNTC_PRECISION_* */
  int32_t prior;                  /* SNTC_PRIOR_*: the hyper-latent prior self._prior, mshyper/models.py:135 (only needed for bits_z) */
} sntc_model_desc;

/* per-image decode metrics: mse_psnr(), common/image_utils.py:26-38 */
typedef struct sntc_image_metrics {
  double mse;
  double psnr;
  uint64_t ssd; /* exact integer sum of squared uint8 differences */
} sntc_image_metrics;

/* per-image rate terms of frame_loss_given_latent_rvs(training=False), in bits (mshyper/models.py:246-259, 278-279):
 *   bits_y = latent_bits       = -sum log2 NoisyNormal(0, SCALE_FN(clamp(exp(raw_sigma), 0, S-1))).prob(q_y)
 *   bits_z = hyper_latent_bits = -sum log2 NoisyDeepFactorized.prob(z_hat)          (0 when the model has no prior)
 * bpp of an image = (bits_y + bits_z) / (H * W)                                      (:300-310) */
typedef struct sntc_image_rate {
  double bits_y;
  double bits_z;
} sntc_image_rate;

typedef struct sntc_ctx sntc_ctx;
typedef struct sntc_model sntc_model;

int sntc_version(void);
const char* sntc_last_error(void);

/* ---- context: one per GPU (one process per GPU in the multi-GPU driver) ---- */
int sntc_create(int device, sntc_ctx** out);
int sntc_destroy(sntc_ctx* ctx);
int sntc_sync(sntc_ctx* ctx);                       /* cudaStreamSynchronize of the context stream + sticky error check */
void* sntc_stream(sntc_ctx* ctx);                   /* the context's cudaStream_t (used when `stream` args are NULL) */
int sntc_device_name(sntc_ctx* ctx, char* buf, size_t n);
/* "domain:bus:device.function" of the context's GPU (cudaDeviceGetPCIBusId): lets the host side pin its feeding thread and
 * its page-locked staging buffers to the GPU's NUMA node (multi-GPU boxes: shallow_ntc_b200.parallel.bind_host_to_gpu). */
int sntc_device_pci_bus_id(sntc_ctx* ctx, char* buf, size_t n);

/* ---- model ----
 * Replaces Model._init_transforms (mshyper/models.py:111-131): transform_builder.build(cls, **kwargs)
 * for "synthesis" and "hyper_synthesis". */
int sntc_model_create(sntc_ctx* ctx, const sntc_model_desc* desc, sntc_model** out);
int sntc_model_destroy(sntc_model* m);
/* Number of variables the model expects and their names/shapes (Keras variable layouts:
 * Conv2DTranspose kernel [kh,kw,Cout,Cin]; SignalConv2D kernel [kh,kw,Cin,Cout]; GDN gamma [C,C], beta [C]). */
int sntc_model_num_variables(sntc_model* m);
int sntc_model_variable(sntc_model* m, int i, const char** name, int64_t shape[4], int* ndim);
/* Replaces tf.train.Checkpoint.restore of the transform variables (common/eval_lib.py:43-45):
 * float32 host array with the effective (de-reparameterised) values; copied, caller keeps ownership. */
int sntc_model_load_weights(sntc_model* m, const char* name, const float* host, const int64_t* shape, int ndim);
/* Packs and uploads all weights; fails with SNTC_E_STATE naming the first missing variable. */
int sntc_model_finalize(sntc_model* m);
/* Small-batch decodes (batch <= 4) whose tensors are all device-resident are replayed as ONE CUDA graph from their third
 * identical call on (first call eager, second captured): the decode is launch-bound there.  Results are identical; the
 * per-stage event times (sntc_last_stage_times_ms) exist for eager decodes only.  on = 0 keeps every decode eager (what
 * Model(profile=True) selects); default on.  Environment: SNTC_GRAPH=0, SNTC_GRAPH_MAX_BATCH. */
int sntc_model_enable_graphs(sntc_model* m, int on);

/* ---- transform-level calls (the Keras-layer __call__ the model code makes) ----
 * self._hyper_synthesis(z_hat)            mshyper/models.py:273  -> out f32 [B, hy, wy, 2*Cy]
 * self._synthesis(y_hat, training=False)  mshyper/models.py:297, factorized/models.py:120-125
 *                                                              -> out f32 [B, Hp, Wp, 3]           */
int sntc_hyper_synthesis(sntc_model* m, const sntc_tensor* z_hat, sntc_tensor* out, void* stream);
int sntc_synthesis(sntc_model* m, const sntc_tensor* y_hat, sntc_tensor* out, void* stream);

/* ---- decoder backward (the gradient tf.GradientTape takes through the two transforms) ----
 * itinf_train_step  mshyper/models.py:401-408:  gradients = tape.gradient(loss, latent_rvs.trainable_variables), where the loss
 * reaches the latents through self._hyper_synthesis(z) (:273) and self._synthesis(y, training=True) (:297).  These two calls are
 * the vector-Jacobian products a tf.custom_gradient around the transform-level calls above returns:
 *   grad_in [B,h,w,Cin] = J_f(x)^T grad_out,   grad_out [B, h*up, w*up, Cout] = d loss / d f(x) on the FULL (padded) output grid
 *   (zeros where unpad_images cropped).  `out` (nullable) receives f(x) of the same pass.
 * The input-gradient of every transposed conv runs as a forward stride-1 band GEMM on a space-to-depth of the gradient; for models
 * created with a tensor-core precision both the forward layers and these backward layers run on tcgen05 (fp32 results kept for the
 * adjoints); GDN1 / relu / leaky_relu adjoints are pointwise kernels.
 * sntc_model_enable_vjp(m, 1) must precede sntc_model_finalize (the backward layers are packed from the host weights);
 * TwoLayerResSynthesis(res_type="d2s") has no backward (SNTC_E_UNSUPPORTED). */
int sntc_model_enable_vjp(sntc_model* m, int on);
int sntc_synthesis_vjp(sntc_model* m, const sntc_tensor* y_hat, const sntc_tensor* grad_out, sntc_tensor* grad_in, sntc_tensor* out, void* stream);
int sntc_hyper_synthesis_vjp(sntc_model* m, const sntc_tensor* z_hat, const sntc_tensor* grad_out, sntc_tensor* grad_in, sntc_tensor* out, void* stream);

/* ---- fused decode ----
 * Replaces mshyper/models.py:269-317 with training=False (factorized/models.py:101-141 when the
 * model has no hyperprior; then z_hat and out_idx must be NULL):
 *   hs = hyper_synthesis(z_hat); mu, sigma = split(hs); i_c = clamp(exp(sigma), 0, S-1); idx = round(i_c)
 *   y_hat = q_y + mu; x = synthesis(y_hat); x = x[:, :H, :W]; u8 = sat_u8(round((x + .5) * 255))
 *   mse, psnr = mse_psnr(original_u8, u8)
 * z_hat  f32 [B, hz, wz, Cz]   integer-valued hyper-latent symbols
 * q_y    f32 | i16 | i8 [B, hy, wy, Cy]   integer latent symbols round(y - mu) from the range decoder
 * out_u8 u8 [B, H, W, 3]; out_idx u8 [B, hy, wy, Cy] (nullable); out_yhat f32 [B, hy, wy, Cy] (nullable);
 * out_f32 f32 [B, H, W, 3] cropped float reconstruction (nullable, debug);
 * original_u8 u8 [B, H, W, 3] (nullable) and metrics[B] (nullable) -- both or neither. */
int sntc_decode(sntc_model* m, const sntc_tensor* z_hat, const sntc_tensor* q_y, int H, int W,
                sntc_tensor* out_u8, sntc_tensor* out_idx, sntc_tensor* out_yhat, sntc_tensor* out_f32,
                const sntc_tensor* original_u8, sntc_image_metrics* metrics, void* stream);

/* sntc_decode plus the rate term: rate[B] receives bits_y / bits_z per image (nullable = sntc_decode).
 * bits_y is accumulated in the epilogue of the last hyper-synthesis GEMM (tensor-core path) from the same
 * raw sigma that produces idx; mu / sigma still never reach HBM.  Mean-scale hyperprior models only. */
int sntc_decode_rd(sntc_model* m, const sntc_tensor* z_hat, const sntc_tensor* q_y, int H, int W,
                   sntc_tensor* out_u8, sntc_tensor* out_idx, sntc_tensor* out_yhat, sntc_tensor* out_f32,
                   const sntc_tensor* original_u8, sntc_image_metrics* metrics, sntc_image_rate* rate, void* stream);

/* MS-SSIM per image of two uint8 batches [B,H,W,C] (host or device), on the device: the validation metric of
 * frame_loss_given_latent_rvs -- tf.image.ssim_multiscale(image, reconstruction, max_val=255.) on the uint8 images, or
 * tf.image.ssim when both sides are < 160 px (mshyper/models.py:321-332, factorized/models.py:145-156).
 * msssim[B] (host memory) is valid on return (the call synchronises `stream`); msssim_db = -10 log10(1 - msssim) (:330).
 * SNTC_E_INVALID when the smallest of the 5 scales is below the 11x11 window (TensorFlow asserts there too). */
int sntc_image_msssim(sntc_ctx* ctx, const sntc_tensor* a_u8, const sntc_tensor* b_u8, double* msssim, void* stream);

/* LPIPS per image of two image batches [B,H,W,3] (uint8, or float32 in [0, 255]; host or device), on the device: the `lpips`
 * column of the evaluate records -- learned_perceptual_metric_model([image_batch, reconstruction]) of the vendored lpips_tf2
 * (lpips_tf2/lpips_tensorflow.py:14-72, called at mshyper/models.py:334-340, factorized/models.py:158-164): Keras VGG16 features of five
 * blocks, unit-normalised over channels, squared difference, learned 1x1 weights, spatial mean, sum over layers.
 * Variables (sntc_lpips_load_weights, float32, Keras layouts): lpips.conv_i.kernel [3,3,Cin,Cout], lpips.conv_i.bias [Cout]
 * (i = 0..12: the 13 VGG16 convolutions in order), lpips.lin_l.kernel [C_l] (l = 0..4: the five Conv2D(1, 1, use_bias=False)).
 * precision: SNTC_PRECISION_TC_F16X3 runs conv_1 .. conv_12 on the tcgen05 band GEMM, SNTC_PRECISION_FP32 everything on FFMA.
 * lpips[B] (host) is valid on return; per_layer[B][5] (nullable) receives the five layer terms. */
typedef struct sntc_lpips sntc_lpips;
int sntc_lpips_create(sntc_ctx* ctx, int precision, sntc_lpips** out);
int sntc_lpips_destroy(sntc_lpips* lp);
int sntc_lpips_load_weights(sntc_lpips* lp, const char* name, const float* host, const int64_t* shape, int ndim);
int sntc_lpips_finalize(sntc_lpips* lp);
int sntc_image_lpips(sntc_lpips* lp, const sntc_tensor* a, const sntc_tensor* b, double* lpips, double* per_layer, void* stream);

/* ---- two-phase decode ----
 * A real decoder cannot have q_y before it knows the scale-table rows: the range decoder needs idx to pick the CDF
 * of every symbol (tfc LocationScaleIndexedEntropyModel.decompress(strings, indexes, loc)).  Phase 1 runs
 * hyper-synthesis and returns idx (mu stays on the device, owned by the model); phase 2 takes the decoded symbols,
 * forms y_hat = q_y + mu and runs synthesis + pixel epilogue.  Results are identical to sntc_decode. */
int sntc_decode_hyper(sntc_model* m, const sntc_tensor* z_hat, sntc_tensor* out_idx, void* stream);
int sntc_decode_latents(sntc_model* m, const sntc_tensor* q_y, int H, int W, sntc_tensor* out_u8, sntc_tensor* out_yhat,
                        sntc_tensor* out_f32, const sntc_tensor* original_u8, sntc_image_metrics* metrics, void* stream);

/* ---- host entropy coder (CPU; the I/O stage either side of the GPU path) ----
 * The reference builds its entropy models with compression=False and never produces a bitstream
 * (mshyper/models.py:246-251), so this defines the missing half: quantised CDF tables in the manner of
 * tensorflow-compression 2.10 (tail_mass, range_coder_precision, overflow symbol + Elias-gamma escape) and a
 * byte-oriented range coder.  kind 0 = scale-table rows of NoisyNormal(0, SCALE_FN(i)), one row index (idx) per
 * symbol; kind 1 = per-channel tables of the hyper-latent prior (row = symbol position % Cz), after
 * sntc_coder_set_prior with the raw DeepFactorized variables packed [Cz][43] in the order matrix_0[3] bias_0[3]
 * factor_0[3] matrix_1[9] bias_1[3] factor_1[3] matrix_2[9] bias_2[3] factor_2[3] matrix_3[3] bias_3[1]. */
typedef struct sntc_coder sntc_coder;
int sntc_coder_create(int num_scales, double scale_min, double scale_max, double tail_mass, int precision, sntc_coder** out);
int sntc_coder_destroy(sntc_coder* k);
int sntc_coder_set_prior(sntc_coder* k, int Cz, const float* raw43);
int sntc_coder_table(sntc_coder* k, int kind, int row, int32_t* offset, int32_t* nsym, const uint32_t** cdf);
int sntc_coder_encode(sntc_coder* k, int kind, const int32_t* symbols, const uint8_t* rows, size_t n, uint8_t** bytes, size_t* nbytes);
int sntc_coder_decode(sntc_coder* k, int kind, const uint8_t* bytes, size_t nbytes, const uint8_t* rows, size_t n, int32_t* symbols);
void sntc_coder_free(void* p); /* releases *bytes of sntc_coder_encode */

/* ---- profile hook: profile_utils.with_timing (common/profile_utils.py:62-76) ----
 * Device time in ms of the stages of the last sntc_decode on this model (CUDA events on the
 * launching stream): [0] hyper_synthesis_time, [1] dequant/index, [2] synthesis_time, [3] total. */
int sntc_last_stage_times_ms(sntc_model* m, float out[4]);
/* Per-layer device timing (CUDA events on the launching stream around each layer of the plan).
 * enable(1) clears the records; while enabled every transform / decode call appends records.
 * Records are aggregated by label ("hyper_synthesis.layer_2", "synthesis.base_conv", "dequant_index", ...):
 * total ms, number of timed intervals, and the algorithmic MACs of ONE interval (reference counting:
 * B*h*w*k*k*Cin*Cout, as in results/flops_per_pixel.csv; 0 for pointwise stages). */
int sntc_profile_enable(sntc_model* m, int on);
int sntc_profile_count(sntc_model* m);
int sntc_profile_get(sntc_model* m, int i, const char** label, float* total_ms, int* intervals, double* macs_per_interval);

/* Number of kernels this library has launched on the context since creation (bench "gpu_launches"). */
uint64_t sntc_launch_count(sntc_ctx* ctx);
/* The same, by kernel family, so that a caller (and the parity tests) can tell WHICH path served a model created with
 * SNTC_PRECISION_TC_F16X3 -- layers that are not TMA-addressable (input channels < 64 or not a multiple of 8, e.g.
 * JPEGLikeSynthesis(use_offset=True)) run on the FFMA band GEMM and show up under SNTC_LAUNCH_BAND_F32:
 *   out[SNTC_LAUNCH_TOTAL] all kernels | [SNTC_LAUNCH_BAND_TC] tcgen05 band GEMM | [SNTC_LAUNCH_BAND_F32] FFMA band GEMM
 *   | [SNTC_LAUNCH_TAIL_MMA] warp-MMA two-layer tail | [SNTC_LAUNCH_TAIL_TC] tcgen05 two-layer tail
 *   | [SNTC_LAUNCH_FINAL_F32] CUDA-core final conv kernels (FFMA tail / rgb cell kernel). */
#define SNTC_LAUNCH_TOTAL 0
#define SNTC_LAUNCH_BAND_TC 1
#define SNTC_LAUNCH_BAND_F32 2
#define SNTC_LAUNCH_TAIL_MMA 3
#define SNTC_LAUNCH_TAIL_TC 4
#define SNTC_LAUNCH_FINAL_F32 5
#define SNTC_LAUNCH_KINDS 6
int sntc_launch_counts(sntc_ctx* ctx, uint64_t out[SNTC_LAUNCH_KINDS]);

/* ---- device / pinned memory helpers for hosts without a CUDA array library ---- */
int sntc_malloc(sntc_ctx* ctx, size_t bytes, void** out);
int sntc_free(sntc_ctx* ctx, void* p);
int sntc_host_alloc(sntc_ctx* ctx, size_t bytes, void** out); /* pinned */
/* pinned, with flags: SNTC_HOST_WRITE_COMBINED = cudaHostAllocWriteCombined -- for staging buffers the host only WRITES (the
 * symbols on their way up): not snooped, so the device reads them faster over PCIe; host reads from such memory are very slow. */
#define SNTC_HOST_WRITE_COMBINED 1
int sntc_host_alloc_flags(sntc_ctx* ctx, size_t bytes, unsigned flags, void** out);
int sntc_host_free(sntc_ctx* ctx, void* p);
int sntc_memcpy_h2d(sntc_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int sntc_memcpy_d2h(sntc_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int sntc_memset(sntc_ctx* ctx, void* dst, int value, size_t bytes, void* stream);

/* ---- extra streams for copy/compute overlap (the streaming decode pipeline) ---- */
int sntc_stream_create(sntc_ctx* ctx, void** out);
int sntc_stream_destroy(sntc_ctx* ctx, void* stream);
int sntc_stream_wait_event(sntc_ctx* ctx, void* stream, void* event); /* stream == NULL: the context stream */
int sntc_stream_sync(sntc_ctx* ctx, void* stream);

/* ---- multi-GPU: the ONE collective of the path ----
 * Images are independent units: ranks never exchange data on the decode path.  What the reference does after the loop is an
 * arithmetic mean of per-image metrics (mshyper/models.py:300-317, common/train_lib.py:64-68); with one process per GPU that is
 * one all-reduce(sum) of [sum psnr, sum mse, sum bits_y, sum bits_z, n_images] over NCCL (NVLink 5 / NVSwitch; 40 bytes: latency
 * only).  libsntc resolves NCCL with dlopen("libnccl.so.2") on first use -- no link-time dependency, no PyTorch.
 *   rank 0: sntc_comm_unique_id(id) -> hand the 128 bytes to every rank (file, socket, MPI, the launcher's store ...)
 *   all   : sntc_comm_create(ctx, id, rank, world, &comm)            (ncclCommInitRank on the context's device)
 *   all   : sntc_allreduce_metrics(comm, sums)  /  sntc_comm_allreduce_f64(comm, v, n, op)  (in place, host doubles; returns
 *           after the result is back in host memory; op SNTC_REDUCE_SUM / SNTC_REDUCE_MAX; n <= 4096) */
#define SNTC_COMM_ID_BYTES 128
#define SNTC_REDUCE_SUM 0
#define SNTC_REDUCE_MAX 1
typedef struct sntc_comm sntc_comm;
int sntc_comm_unique_id(void* id128);
int sntc_comm_create(sntc_ctx* ctx, const void* id128, int rank, int world, sntc_comm** out);
int sntc_comm_destroy(sntc_comm* comm);
int sntc_comm_allreduce_f64(sntc_comm* comm, double* values, int n, int op);
int sntc_allreduce_metrics(sntc_comm* comm, double sums[5]);

/* ---- CUDA-event timers on the launching stream (bench.py) ---- */
int sntc_event_create(sntc_ctx* ctx, void** out);
int sntc_event_destroy(sntc_ctx* ctx, void* ev);
int sntc_event_record(sntc_ctx* ctx, void* ev, void* stream);
int sntc_event_elapsed_ms(sntc_ctx* ctx, void* start, void* stop, float* ms); /* synchronizes on stop */

#ifdef __cplusplus
}
#endif
#endif /* SNTC_H_ */
