"""Decode-side mirror of ``mshyper/models.py`` / ``factorized/models.py``.

``Model(transform_config=...)`` takes the reference's config dict (``mshyper/configs/*.py``),
builds the synthesis and hyper-synthesis transforms through the same registry mechanism
(``mshyper/models.py:111-131``) and exposes the generative half of
``frame_loss_given_latent_rvs(training=False)`` (``mshyper/models.py:269-317``) as ``decompress`` --
a method the reference does not have (it never runs a range coder); it starts from the
decoded integer symbols and runs entirely in libsntc on the GPU.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from ._lib import lib, check, ModelDesc, TransformDesc, ImageMetrics, ImageRate
from .tensors import Context, as_tensor, empty_like_kind, DeviceArray
from .transforms import class_builder as transform_builder

# Scale-table row rule.  tfc 2.10 ContinuousIndexedEntropyModel._flatten_indexes is tf.cast(indexes, tf.int32):
# truncation.  'rint' stays selectable for coders that round.  The default is a deliberate choice, recorded in DESIGN.md section 5.
DEFAULT_INDEX_ROUNDING = "trunc"

# Fixed configs for the ScaleIndexedEntropyModel (mshyper/models.py:27-34).
NUM_SCALES = 64
SCALE_MIN = 0.11
SCALE_MAX = 256.
_PRECISION = {"fp32": _lib.PRECISION_FP32, "tc": _lib.PRECISION_TC_F16X3, "tc_f16x3": _lib.PRECISION_TC_F16X3,
              "tc_syn2": _lib.PRECISION_TC_F16X3_SYN2}   # opt-in: 2-pass synthesis layers (see include/sntc.h)
_ROUNDING = {"rint": _lib.INDEX_RINT, "trunc": _lib.INDEX_TRUNC}


class _NativeModel:
  def __init__(self, ctx, handle):
    self.ctx, self.handle = ctx, handle

  def variables(self):
    out = {}
    for i in range(lib.sntc_model_num_variables(self.handle)):
      name = C.c_char_p()
      shape = (C.c_int64 * 4)()
      nd = C.c_int()
      check(lib.sntc_model_variable(self.handle, i, C.byref(name), shape, C.byref(nd)))
      out[name.value.decode()] = tuple(shape[d] for d in range(nd.value))
    return out

  def __del__(self):
    try:
      if self.handle and self.ctx.handle:
        lib.sntc_model_destroy(self.handle)
    except Exception:
      pass


def _create_model(ctx: Context, hyper: TransformDesc, syn: TransformDesc, weights: dict, precision="fp32",
                  index_rounding=DEFAULT_INDEX_ROUNDING, num_scales=NUM_SCALES, prior=False, vjp=False) -> _NativeModel:
  desc = ModelDesc(struct_size=C.sizeof(ModelDesc), hyper=hyper, synthesis=syn, num_scales=num_scales,
                   index_rounding=_ROUNDING[index_rounding], precision=_PRECISION[precision],
                   prior=_lib.PRIOR_DEEP_FACTORIZED if prior else _lib.PRIOR_NONE)
  h = C.c_void_p()
  check(lib.sntc_model_create(ctx.handle, C.byref(desc), C.byref(h)))
  m = _NativeModel(ctx, h)
  for name, shape in m.variables().items():
    if name not in weights:
      raise KeyError(f"missing variable {name} {shape}")
    w = np.ascontiguousarray(weights[name], dtype=np.float32)
    shp = (C.c_int64 * w.ndim)(*w.shape)
    check(lib.sntc_model_load_weights(h, name.encode(), w.ctypes.data_as(C.POINTER(C.c_float)), shp, w.ndim))
  if vjp:   # decoder backward: the backward layers are packed from the host weights at finalize
    check(lib.sntc_model_enable_vjp(h, 1))
  check(lib.sntc_model_finalize(h))
  return m


def bottleneck_size_of(analysis_cfg: dict) -> int:
  """What ``Model._init_transforms`` learns by pushing a dummy image through the analysis transform
  (``mshyper/models.py:117-119``), read off the config instead (the encoder is out of scope)."""
  cls = analysis_cfg.get("cls")
  if cls == "ElicAnalysis":
    return int(analysis_cfg["channels"][-1])
  if cls in ("CNNAnalysis", "MBT2018Analysis"):
    oc = analysis_cfg.get("output_channels")
    return int(oc if oc is not None else analysis_cfg["channels_base"])
  if cls == "BLS2017Analysis":
    return int(analysis_cfg["num_filters"])
  raise NotImplementedError(f"cannot infer the bottleneck size of analysis class {cls!r}; pass bottleneck_size=")


class Model:
  """Mean-scale hyperprior decode model (``mshyper/models.py``); ``hyperprior=False`` gives the
  factorized-prior model (``factorized/models.py``: no z, no scale indexes, DOWNSAMPLE_FACTOR 16)."""

  def __init__(self, transform_config, bottleneck_size=None, hyperprior=True, profile=False, device=0,
               precision="fp32", index_rounding=DEFAULT_INDEX_ROUNDING, ctx: Context | None = None, prior=False, vjp=False,
               **_ignored_training_kwargs):
    self._transform_config = transform_config
    self._vjp = bool(vjp)      # also build the decoder backward (synthesis_vjp / hyper_synthesis_vjp)
    self._with_prior = bool(prior) and hyperprior    # self._prior = tfc.NoisyDeepFactorized(...)   mshyper/models.py:135
    self._profile = profile
    self.precision = precision
    self.index_rounding = index_rounding
    self.hyperprior = hyperprior
    if bottleneck_size is None:
      bottleneck_size = bottleneck_size_of(dict(transform_config["analysis"]))
    self._bottleneck_size = int(bottleneck_size)
    self._init_transforms(transform_config)
    self._ctx = ctx
    self._device = device
    self._native = None
    self._weights = None

  def _init_transforms(self, transform_config):
    synthesis_cfg = dict(transform_config["synthesis"])
    self._synthesis = transform_builder.build(synthesis_cfg.pop("cls"), **synthesis_cfg)   # mshyper/models.py:114-115
    self._synthesis.in_channels = self._bottleneck_size
    if self.hyperprior:
      if "hyper_synthesis" in transform_config:
        hyper_synthesis_cfg = dict(transform_config["hyper_synthesis"])
      else:
        hyper_synthesis_cfg = dict(cls="HyperSynthesis", bottleneck_size=self._bottleneck_size)            # :126-129
      self._hyper_synthesis = transform_builder.build(hyper_synthesis_cfg.pop("cls"), **hyper_synthesis_cfg)
      # the hyper-latent has bottleneck_size channels for every shipped hyper-analysis (HyperAnalysis :209-219)
      self._hyper_synthesis.in_channels = int(transform_config.get("hyper_bottleneck_size", self._bottleneck_size))
      self.downsample_factor = self._synthesis.upsample * self._hyper_synthesis.upsample                   # :137-140
    else:
      self._hyper_synthesis = None
      self.downsample_factor = self._synthesis.upsample                                                   # factorized/models.py:30

  # ---------------------------------------------------------------------------------------------
  @property
  def latent_channels(self):
    return self._bottleneck_size

  @property
  def hyper_channels(self):
    return self._hyper_synthesis.in_channels if self._hyper_synthesis is not None else 0

  def variable_shapes(self) -> dict:
    v = {}
    if self._hyper_synthesis is not None:
      v.update(self._hyper_synthesis.variable_shapes(self._hyper_synthesis.in_channels))
    v.update(self._synthesis.variable_shapes(self._bottleneck_size))
    if self._with_prior:   # raw tfc.DeepFactorized variables, num_filters=(3,3,3)
      f, Cz = (1, 3, 3, 3, 1), self.hyper_channels
      for i in range(4):
        v[f"prior.matrix_{i}"] = (Cz, f[i + 1], f[i])
        v[f"prior.bias_{i}"] = (Cz, f[i + 1], 1)
        if i < 3:
          v[f"prior.factor_{i}"] = (Cz, f[i + 1], 1)
    return v

  def latent_shapes(self, batch, H, W):
    """Shapes of (z_hat, q_y) for images of H x W: pad_images to a multiple of downsample_factor
    (mshyper/models.py:218, image_utils.py:41-66)."""
    d = self.downsample_factor
    Hp, Wp = -(-H // d) * d, -(-W // d) * d
    us = self._synthesis.upsample
    y = (batch, Hp // us, Wp // us, self._bottleneck_size)
    z = (batch, Hp // d, Wp // d, self.hyper_channels) if self._hyper_synthesis is not None else None
    return z, y

  def load_weights(self, weights: dict):
    self._weights = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
    self._native = None

  def _ensure_native(self):
    if self._native is not None:
      return
    if self._weights is None:
      raise RuntimeError("load_weights() must be called before decompress()")
    if self._ctx is None:
      self._ctx = Context(self._device)
    hyper = (self._hyper_synthesis.desc(self._hyper_synthesis.in_channels) if self._hyper_synthesis is not None
             else TransformDesc(kind=_lib.T_NONE))
    syn = self._synthesis.desc(self._bottleneck_size)
    self._native = _create_model(self._ctx, hyper, syn, self._weights, self.precision, self.index_rounding, prior=self._with_prior,
                                 vjp=self._vjp)
    if self._profile:   # stage times come from CUDA events around eagerly launched kernels: no graph replay
      check(lib.sntc_model_enable_graphs(self._native.handle, 0))

  @property
  def ctx(self) -> Context:
    if self._ctx is None:
      self._ctx = Context(self._device)
    return self._ctx

  # ---------------------------------------------------------------------------------------------
  def decompress(self, z_hat, q_y, image_hw, *, return_idx=True, return_yhat=False, return_float=False,
                 original=None, return_bits=False, out=None, stream=None, sync=True):
    """Decode a batch from its integer symbols.

    z_hat [B, Hp/64, Wp/64, Cz] float32 (None for the factorized model); q_y [B, Hp/16, Wp/16, Cy]
    float32 / int16 / int8 = round(y - mu); image_hw = (H, W) of the un-padded images.
    Inputs may be numpy (host; copies are part of the call), DeviceArray / anything with
    ``__cuda_array_interface__`` or ``__dlpack__`` (zero-copy).  Returns a dict with ``image`` uint8
    [B,H,W,3] and, as requested, ``idx`` uint8 (scale-table rows), ``y_hat``, ``float`` (cropped float
    reconstruction in [-0.5, 0.5]), ``mse`` / ``psnr`` per image (when ``original`` uint8 is given), and
    ``hyper_synthesis_time`` / ``synthesis_time`` seconds when ``profile=True`` (profile_utils.with_timing).
    ``return_bits=True`` adds the rate term (mshyper/models.py:278-279, 300-310): ``bits_y``, ``bits_z`` per image
    (``bits_z`` = 0 unless the model was built with ``prior=True``) and ``bpp`` = (bits_y + bits_z) / (H * W).
    ``out`` may carry pre-allocated buffers under the same keys."""
    self._ensure_native()
    ctx = self._ctx
    H, W = int(image_hw[0]), int(image_hw[1])
    q = as_tensor(q_y, ctx.device)
    B, hy, wy, Cy = q.shape
    z = as_tensor(z_hat, ctx.device) if self._hyper_synthesis is not None else None
    if self._hyper_synthesis is None and z_hat is not None:
      raise ValueError("the factorized model takes no z_hat")
    out = dict(out or {})
    Co = self._synthesis.out_channels

    def buf(key, shape, dtype, want):
      if not want:
        return None
      if key not in out:
        out[key] = empty_like_kind(ctx, q_y, shape, dtype)
      return as_tensor(out[key], ctx.device)

    t_img = buf("image", (B, H, W, Co), np.uint8, True)
    t_idx = buf("idx", (B, hy, wy, Cy), np.uint8, return_idx and self._hyper_synthesis is not None)
    t_yh = buf("y_hat", (B, hy, wy, Cy), np.float32, return_yhat)
    t_f = buf("float", (B, H, W, Co), np.float32, return_float)
    t_orig = as_tensor(original, ctx.device)
    metrics = (ImageMetrics * B)() if original is not None else None
    rate = (ImageRate * B)() if return_bits else None
    check(lib.sntc_decode_rd(self._native.handle, z.byref() if z else None, q.byref(), H, W, t_img.byref(),
                             t_idx.byref() if t_idx else None, t_yh.byref() if t_yh else None, t_f.byref() if t_f else None,
                             t_orig.byref() if t_orig else None, metrics, rate, stream))
    if sync:
      ctx.sync()
    if rate is not None:
      out["bits_y"] = np.array([r.bits_y for r in rate])
      out["bits_z"] = np.array([r.bits_z for r in rate])
      out["bpp"] = (out["bits_y"] + out["bits_z"]) / float(H * W)
    if metrics is not None:
      out["mse"] = np.array([m.mse for m in metrics])
      out["psnr"] = np.array([m.psnr for m in metrics])
      out["ssd"] = np.array([m.ssd for m in metrics], dtype=np.uint64)
    if self._profile:
      t = (C.c_float * 4)()
      check(lib.sntc_last_stage_times_ms(self._native.handle, t))
      out["hyper_synthesis_time"] = (t[0] + t[1]) * 1e-3
      out["synthesis_time"] = t[2] * 1e-3
    return out

  def decode_hyper(self, z_hat, out=None):
    """Phase 1 of the two-phase decode (sntc_decode_hyper): hyper-synthesis -> scale-table rows idx uint8
    [B,hy,wy,Cy] for the range decoder; mu stays on the device for ``decode_latents``."""
    self._ensure_native()
    z = as_tensor(z_hat, self._ctx.device)
    B, hz, wz, _ = z.shape
    up = self._hyper_synthesis.upsample
    if out is None:
      out = empty_like_kind(self._ctx, z_hat, (B, hz * up, wz * up, self._bottleneck_size), np.uint8)
    o = as_tensor(out, self._ctx.device)
    check(lib.sntc_decode_hyper(self._native.handle, z.byref(), o.byref(), None))
    self._ctx.sync()
    return out

  def decode_latents(self, q_y, image_hw, *, return_yhat=False, return_float=False, original=None, out=None):
    """Phase 2 (sntc_decode_latents): y_hat = q_y + mu (mu of the last ``decode_hyper``), synthesis, crop, uint8."""
    self._ensure_native()
    ctx = self._ctx
    H, W = int(image_hw[0]), int(image_hw[1])
    q = as_tensor(q_y, ctx.device)
    B, hy, wy, Cy = q.shape
    Co = self._synthesis.out_channels
    out = dict(out or {})

    def buf(key, shape, dtype, want):
      if not want:
        return None
      if key not in out:
        out[key] = empty_like_kind(ctx, q_y, shape, dtype)
      return as_tensor(out[key], ctx.device)

    t_img = buf("image", (B, H, W, Co), np.uint8, True)
    t_yh = buf("y_hat", (B, hy, wy, Cy), np.float32, return_yhat)
    t_f = buf("float", (B, H, W, Co), np.float32, return_float)
    t_orig = as_tensor(original, ctx.device)
    metrics = (ImageMetrics * B)() if original is not None else None
    check(lib.sntc_decode_latents(self._native.handle, q.byref(), H, W, t_img.byref(), t_yh.byref() if t_yh else None,
                                  t_f.byref() if t_f else None, t_orig.byref() if t_orig else None, metrics, None))
    ctx.sync()
    if metrics is not None:
      out["mse"] = np.array([m.mse for m in metrics])
      out["psnr"] = np.array([m.psnr for m in metrics])
      out["ssd"] = np.array([m.ssd for m in metrics], dtype=np.uint64)
    return out

  # ---------------------------------------------------------------------------------------------
  # intra-frame sharding (BASELINE configs[4]): latent-row bands with recomputed halos, see tiling.py
  def band_plan(self, image_hw, n_bands):
    from .tiling import plan_bands
    hyp = self._hyper_synthesis
    return plan_bands(int(image_hw[0]), self._synthesis.conv_chain(), self._synthesis.upsample,
                      hyp.conv_chain() if hyp is not None else None, hyp.upsample if hyp is not None else 1, int(n_bands))

  def band_inputs(self, z_hat, q_y, band):
    """The (contiguous) sub-tensors of host symbol arrays a band decodes: what a rank would be handed by the range decoder."""
    z = None if z_hat is None else np.ascontiguousarray(z_hat[:, band.z_rows[0]:band.z_rows[1]])
    return z, np.ascontiguousarray(q_y[:, band.y_rows[0]:band.y_rows[1]])

  def decompress_band(self, z_band, q_band, image_hw, band, **kw):
    """Decode ONE band from its sub-tensors (``band_inputs``).  Returns the usual dict restricted to the band: ``image``
    [B, r1-r0, W, 3] (rows ``band.rows`` of the frame), ``idx`` / ``y_hat`` [B, c1-c0, wy, Cy] (latent rows ``band.y_core``).
    Host or device inputs; the crops are views / device-to-host slices of the sub-decode's outputs."""
    W = int(image_hw[1])
    full = self.decompress(z_band, q_band, (band.sub_h, W), **kw)
    out = dict(full)
    k0, k1 = band.keep
    c0, c1 = band.y_core[0] - band.y_rows[0], band.y_core[1] - band.y_rows[0]
    for key, (a, b) in (("image", (k0, k1)), ("float", (k0, k1)), ("idx", (c0, c1)), ("y_hat", (c0, c1))):
      if key in out:
        v = out[key] if isinstance(out[key], np.ndarray) else out[key].to_host()
        out[key] = v[:, a:b]
    return out

  def decompress_tiled(self, z_hat, q_y, image_hw, n_bands, **kw):
    """Whole frames decoded band by band on THIS GPU and stitched: bit-identical to ``decompress`` (the test of the halo
    derivation).  On N GPUs rank r runs ``decompress_band`` on band r only; no collective is needed to produce the frame
    (each rank owns its rows)."""
    bands = self.band_plan(image_hw, n_bands)
    parts = [self.decompress_band(*self.band_inputs(z_hat, q_y, b), image_hw, b, **kw) for b in bands]
    out = {}
    for key in ("image", "float", "idx", "y_hat"):
      if key in parts[0]:
        out[key] = np.concatenate([p[key] for p in parts], axis=1)
    return out

  def evaluate(self, z_hat, q_y, originals, image_hw=None, **kw):
    """Model.evaluate (mshyper/models.py:415-433) from decoded symbols: yields one metrics dict per image
    (see eval_lib.evaluate_symbols)."""
    from .eval_lib import evaluate_symbols
    return evaluate_symbols(self, z_hat, q_y, originals, image_hw, **kw)

  def profile_layers(self, on: bool):
    """Per-layer CUDA-event timing inside libsntc (sntc_profile_enable)."""
    self._ensure_native()
    check(lib.sntc_profile_enable(self._native.handle, int(bool(on))))

  def layer_profile(self) -> dict:
    """label -> dict(ms total, n intervals, macs per interval)."""
    self._ensure_native()
    out = {}
    for i in range(lib.sntc_profile_count(self._native.handle)):
      name, ms, n, macs = C.c_char_p(), C.c_float(), C.c_int(), C.c_double()
      check(lib.sntc_profile_get(self._native.handle, i, C.byref(name), C.byref(ms), C.byref(n), C.byref(macs)))
      out[name.value.decode()] = dict(ms=float(ms.value), n=int(n.value), macs=float(macs.value))
    return out

  def hyper_synthesis(self, z_hat, out=None):
    """self._hyper_synthesis(z_hat) (mshyper/models.py:273) -> [B, hy, wy, 2*Cy] = mu || raw sigma."""
    self._ensure_native()
    z = as_tensor(z_hat, self._ctx.device)
    B, h, w, _ = z.shape
    up = self._hyper_synthesis.upsample
    if out is None:
      out = empty_like_kind(self._ctx, z_hat, (B, h * up, w * up, self._hyper_synthesis.out_channels), np.float32)
    o = as_tensor(out, self._ctx.device)
    check(lib.sntc_hyper_synthesis(self._native.handle, z.byref(), o.byref(), None))
    self._ctx.sync()
    return out

  def synthesis(self, y_hat, out=None):
    """self._synthesis(y_hat, training=False) (mshyper/models.py:297) -> [B, Hp, Wp, 3] float32."""
    self._ensure_native()
    y = as_tensor(y_hat, self._ctx.device)
    B, h, w, _ = y.shape
    up = self._synthesis.upsample
    if out is None:
      out = empty_like_kind(self._ctx, y_hat, (B, h * up, w * up, self._synthesis.out_channels), np.float32)
    o = as_tensor(out, self._ctx.device)
    check(lib.sntc_synthesis(self._native.handle, y.byref(), o.byref(), None))
    self._ctx.sync()
    return out


  # ---------------------------------------------------------------------------------------------
  # decoder backward: what tape.gradient propagates through the two transforms in itinf_train_step
  # (mshyper/models.py:401-408).  Model(..., vjp=True).
  def _vjp_call(self, fn, transform, x, grad_out, return_out, grad_in=None, sync=True):
    if not self._vjp:
      raise RuntimeError("construct the model with vjp=True to use the decoder backward")
    self._ensure_native()
    xt = as_tensor(x, self._ctx.device)
    B, h, w, _ = xt.shape
    up = transform.upsample
    gt = as_tensor(grad_out, self._ctx.device)
    gin = grad_in if grad_in is not None else empty_like_kind(self._ctx, x, tuple(xt.shape), np.float32)
    out = empty_like_kind(self._ctx, x, (B, h * up, w * up, transform.out_channels), np.float32) if return_out else None
    check(fn(self._native.handle, xt.byref(), gt.byref(), as_tensor(gin, self._ctx.device).byref(),
             as_tensor(out, self._ctx.device).byref() if return_out else None, None))
    if sync:
      self._ctx.sync()
    return (gin, out) if return_out else gin

  def synthesis_vjp(self, y_hat, grad_out, return_out=False, **kw):
    """J^T grad_out of self._synthesis at y_hat: grad_out is d loss / d synthesis(y_hat) on the full padded grid [B, Hp, Wp, 3]
    (zeros where unpad_images cropped, mshyper/models.py:297-298); returns d loss / d y_hat [B, hy, wy, Cy] (and the forward output)."""
    return self._vjp_call(lib.sntc_synthesis_vjp, self._synthesis, y_hat, grad_out, return_out, **kw)

  def hyper_synthesis_vjp(self, z_hat, grad_out, return_out=False, **kw):
    """J^T grad_out of self._hyper_synthesis at z_hat: grad_out [B, hy, wy, 2*Cy] = d loss / d (mu || raw sigma) (mshyper/models.py:273-275)."""
    return self._vjp_call(lib.sntc_hyper_synthesis_vjp, self._hyper_synthesis, z_hat, grad_out, return_out, **kw)


class FactorizedModel(Model):
  """factorized/models.py: the same decode without a hyperprior."""

  def __init__(self, transform_config, **kw):
    kw.pop("hyperprior", None)
    super().__init__(transform_config, hyperprior=False, **kw)

  def decompress(self, *args, **kw):
    """decompress(q_y, image_hw, ...) -- also accepts the mshyper form decompress(None, q_y, image_hw, ...)."""
    if len(args) == 3:
      if args[0] is not None:
        raise ValueError("the factorized model takes no z_hat")
      args = args[1:]
    q_y, image_hw = args
    return super().decompress(None, q_y, image_hw, **kw)


# The reference's five benchmark configurations (transform_config dicts copied from the config files).
CONFIGS = {
  "jpegl": dict(  # mshyper/configs/jpegl.py:36-39
    analysis=dict(cls="ElicAnalysis", channels=(192, 192, 192, 320)),
    synthesis=dict(cls="JPEGLikeSynthesis", kernel_size=18, strides=16)),
  "two_layer_syn": dict(  # mshyper/configs/two_layer_syn.py:36-40
    analysis=dict(cls="ElicAnalysis", channels=(192, 192, 192, 320)),
    synthesis=dict(cls="TwoLayerResSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5),
                   activation_type="igdn", res_type="conv")),
  "two_layer_syn2": dict(  # mshyper/configs/two_layer_syn2.py:47-50 (C1 in {12, 24, 48}, :87-89)
    analysis=dict(cls="CNNAnalysis", channels_base=256, output_channels=320),
    synthesis=dict(cls="TwoLayerSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5), activation_type="igdn")),
  "mbt2018": dict(  # mshyper/configs/mbt2018.py:34-39
    analysis=dict(cls="MBT2018Analysis", channels_base=192, output_channels=320),
    synthesis=dict(cls="MBT2018Synthesis", channels_base=192, output_channels=3)),
  "bls2017": dict(  # factorized/configs/bls2017.py:35-38
    analysis=dict(cls="BLS2017Analysis", num_filters=256),
    synthesis=dict(cls="BLS2017Synthesis", num_filters=256)),
}


def build_config(name: str, **kw) -> Model:
  cfg = {k: dict(v) for k, v in CONFIGS[name.split(":")[0]].items()}
  if ":" in name:  # e.g. two_layer_syn2:24
    c1 = int(name.split(":")[1])
    cfg["synthesis"]["channels"] = (c1, 3)
  if name.startswith("bls2017"):
    return FactorizedModel(cfg, **kw)
  return Model(cfg, **kw)
