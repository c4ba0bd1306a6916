"""Decoder backward (f4: the gradient through the transforms in itinf_train_step, mshyper/models.py:401-408).
CPU part: the two statements of the oracle (numpy adjoints by definition, torch float64 autograd) against each other and against
the forward oracle.  GPU part (-m gpu): libsntc's sntc_*_vjp against the autograd oracle, fp32 and tensor-core backward."""
import numpy as np
import pytest

from shallow_ntc_b200 import transforms as T, synthetic
from oracle import ntc_oracle as O
from oracle import torch_vjp as V

CASES = [   # (class, ctor kwargs, input channels, latent h x w)
  ("HyperSynthesis", dict(bottleneck_size=16), 16, (3, 4)),
  ("HyperSynthesis", dict(bottleneck_size=16, activation_type="leaky_relu"), 16, (2, 3)),
  ("JPEGLikeHyperSynthesis", dict(bottleneck_size=8), 12, (3, 2)),
  ("HyperSynthesisSmall", dict(bottleneck_size=8), 8, (4, 3)),
  ("JPEGLikeSynthesis", dict(kernel_size=18, strides=16), 16, (2, 3)),
  ("JPEGLikeSynthesis", dict(kernel_size=18, strides=16, use_offset=True, use_bias=False), 16, (2, 2)),
  ("TwoLayerSynthesis", dict(channels=(12, 3)), 16, (2, 3)),
  ("TwoLayerSynthesis", dict(channels=(8, 3), activation_type="relu"), 16, (2, 2)),
  ("TwoLayerSynthesis", dict(channels=(8, 3), activation_type="gdn"), 16, (2, 2)),
  ("TwoLayerResSynthesis", dict(channels=(12, 3)), 16, (3, 2)),
  ("MBT2018Synthesis", dict(channels_base=8), 8, (2, 3)),
  ("BLS2017Synthesis", dict(num_filters=8), 8, (3, 2)),
  ("CNNSynthesis", dict(channels_base=8, activation_type="igdn"), 8, (2, 2)),
  ("CNNSynthesis", dict(channels_base=8), 8, (2, 2)),
]


def small_case(cls, kw, cin, hw, seed=5, B=2):
  t = T.class_builder.build(cls, **kw)
  wts = synthetic.make_weights(t.variable_shapes(cin), "stress", synthesis_cls=cls)
  rng = np.random.default_rng(seed)
  x = rng.standard_normal((B, hw[0], hw[1], cin)) * 1.5
  g = rng.standard_normal((B, hw[0] * t.upsample, hw[1] * t.upsample, t.out_channels))
  okw = {k: v for k, v in kw.items() if k not in ("bottleneck_size", "channels_base", "num_filters", "channels", "kernel_size")}
  return t, wts, x, g, okw


@pytest.mark.parametrize("cls,kw,cin,hw", CASES)
def test_autograd_statement_forward_matches_the_oracle(cls, kw, cin, hw):
  t, wts, x, g, okw = small_case(cls, kw, cin, hw)
  fn = O.hyper_synthesis_by_name if t.role == "hyper_synthesis" else O.synthesis
  want = fn(cls, wts, x, okw)
  out, gin = V.transform_vjp(cls, wts, x, g, okw)
  assert out.shape == want.shape and np.abs(out - want).max() < 1e-10 * max(1.0, np.abs(want).max())
  assert gin.shape == x.shape
  # <g, J v> == <J^T g, v>
  v = np.random.default_rng(9).standard_normal(x.shape)
  jv = V.transform_jvp(cls, wts, x, v, okw)
  lhs, rhs = float((g * jv).sum()), float((gin * v).sum())
  assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))


def test_numpy_adjoints_match_autograd():
  """conv_transpose_input_grad / gdn1_vjp / activation_vjp by definition == autograd, through two whole transforms."""
  t, wts, x, g, okw = small_case("HyperSynthesis", dict(bottleneck_size=16), 16, (3, 4))
  out, gin = O.hyper_synthesis_vjp(wts, x, g)
  o2, g2 = V.transform_vjp("HyperSynthesis", wts, x, g, okw)
  assert np.abs(out - o2).max() < 1e-10 and np.abs(gin - g2).max() < 1e-10 * max(1.0, np.abs(g2).max())
  for cls, res in (("TwoLayerResSynthesis", True), ("TwoLayerSynthesis", False)):
    for act in ("igdn", "gdn", "relu"):
      t, wts, x, g, okw = small_case(cls, dict(channels=(12, 3), activation_type=act), 16, (3, 2))
      out, gin = O.two_layer_res_synthesis_vjp(wts, x, g, activation_type=act, res=res)
      o2, g2 = V.transform_vjp(cls, wts, x, g, okw)
      assert np.abs(out - o2).max() < 1e-10 and np.abs(gin - g2).max() < 1e-10 * max(1.0, np.abs(g2).max()), (cls, act)


def test_conv_transpose_input_grad_geometries():
  """every (k, s, p) on the path, ragged sizes: the adjoint identity <g, conv(x)> == <conv^T(g), x>."""
  rng = np.random.default_rng(3)
  for k, s, keras in ((13, 8, True), (5, 2, True), (3, 1, True), (18, 16, True), (6, 4, True), (5, 2, False), (9, 4, False), (3, 1, False)):
    p = O.keras_same_pad(k, s) if keras else O.tfc_same_pad(k)
    w = rng.standard_normal((k, k, 3, 4))
    x = rng.standard_normal((1, 3, 2, 4))
    g = rng.standard_normal((1, 3 * s, 2 * s, 3))
    y = O.conv_transpose_scatter(x, w, None, s, p)
    gx = O.conv_transpose_input_grad(g, w, s, p)
    assert abs(float((g * y).sum()) - float((gx * x).sum())) < 1e-9


# --------------------------------------------------------------------------------------------------
# -m gpu: libsntc's decoder backward against the autograd oracle

GPU_CASES = [   # (class, ctor kwargs, input channels, latent h x w, batch)
  ("HyperSynthesis", dict(bottleneck_size=64), 64, (5, 7), 2),
  ("HyperSynthesis", dict(bottleneck_size=64, activation_type="leaky_relu"), 64, (4, 3), 1),
  ("JPEGLikeHyperSynthesis", dict(bottleneck_size=32), 64, (3, 5), 2),
  ("HyperSynthesisSmall", dict(bottleneck_size=64), 64, (6, 5), 1),
  ("JPEGLikeSynthesis", dict(kernel_size=18, strides=16), 64, (3, 4), 2),
  ("JPEGLikeSynthesis", dict(kernel_size=18, strides=16, use_offset=True, use_bias=False), 64, (2, 3), 1),
  ("TwoLayerSynthesis", dict(channels=(12, 3)), 64, (3, 4), 2),
  ("TwoLayerSynthesis", dict(channels=(24, 3), activation_type="relu"), 64, (2, 3), 1),
  ("TwoLayerSynthesis", dict(channels=(12, 3), activation_type="gdn"), 64, (2, 3), 1),
  ("TwoLayerResSynthesis", dict(channels=(12, 3)), 64, (4, 3), 2),
  ("TwoLayerResSynthesis", dict(channels=(12, 3), activation_type="leaky_relu"), 64, (2, 2), 1),
  ("MBT2018Synthesis", dict(channels_base=96), 64, (3, 4), 1),
  ("BLS2017Synthesis", dict(num_filters=128), 128, (3, 2), 1),
  ("CNNSynthesis", dict(channels_base=64, activation_type="igdn"), 64, (2, 3), 1),
  ("CNNSynthesis", dict(channels_base=64), 64, (2, 2), 2),
]


def _rel(a, b):
  return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("cls,kw,cin,hw,B", GPU_CASES)
def test_transform_vjp_matches_autograd(gpu_ctx, cls, kw, cin, hw, B, precision):
  """Transform.vjp (sntc_synthesis_vjp / sntc_hyper_synthesis_vjp through the C ABI) == J^T g of the float64 autograd oracle:
  every registry class, fp32 band GEMMs and the tcgen05 band GEMMs of the tensor-core precision."""
  t, wts, x, g, okw = small_case(cls, kw, cin, hw, B=B)
  t._ctx = gpu_ctx
  t.precision = precision
  t.load_weights(wts)
  x32, g32 = x.astype(np.float32), g.astype(np.float32)
  before = gpu_ctx.launch_counts
  gin = t.vjp(x32, g32)
  after = gpu_ctx.launch_counts
  out, want = V.transform_vjp(cls, wts, x32.astype(np.float64), g32.astype(np.float64), okw)
  assert gin.shape == x.shape and gin.dtype == np.float32
  err = _rel(gin, want)
  tol = 2e-5 if precision == "fp32" else 1e-4     # relative to max |grad|; the split-fp16 product carries ~2^-22 per MAC
  assert err < tol, (cls, precision, err)
  if precision == "tc":   # the wide backward layers (s*s*Cout >= 64 input channels) must have run on the tensor cores
    assert after["band_tc"] > before["band_tc"], (before, after)
  else:
    assert after["band_tc"] == before["band_tc"]
  # the forward value of the same pass
  fwd = t(x32)
  assert _rel(fwd, out) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_model_decoder_vjp_headline_shapes(gpu_ctx, precision):
  """Model(vjp=True).synthesis_vjp / hyper_synthesis_vjp on the headline architecture (two_layer_syn, 320 channels) for one
  128 x 192 image: the two heavy pieces of tape.gradient in itinf_train_step (mshyper/models.py:401-408)."""
  from shallow_ntc_b200 import build_config
  model = build_config("two_layer_syn", precision=precision, ctx=gpu_ctx, vjp=True)
  cfg = model._transform_config["synthesis"]
  wts = synthetic.make_weights(model.variable_shapes(), "stress", synthesis_cls=cfg["cls"])
  model.load_weights(wts)
  H, W = 128, 192
  zs, ys = model.latent_shapes(1, H, W)
  rng = np.random.default_rng(2)
  y = (rng.standard_normal(ys) * 2).astype(np.float32)
  z = np.rint(rng.standard_normal(zs) * 1.5).astype(np.float32)
  gx = rng.standard_normal((1, H, W, 3)).astype(np.float32)
  gh = rng.standard_normal((1, ys[1], ys[2], 2 * ys[3])).astype(np.float32)
  gy, out = model.synthesis_vjp(y, gx, return_out=True)
  kw = {k: v for k, v in cfg.items() if k not in ("cls", "channels", "kernel_sizes")}
  o_ref, gy_ref = V.transform_vjp(cfg["cls"], wts, y.astype(np.float64), gx.astype(np.float64), kw)
  tol = 2e-5 if precision == "fp32" else 1e-4
  assert _rel(gy, gy_ref) < tol and _rel(out, o_ref) < tol, (_rel(gy, gy_ref), _rel(out, o_ref))
  gz = model.hyper_synthesis_vjp(z, gh)
  _, gz_ref = V.transform_vjp("HyperSynthesis", wts, z.astype(np.float64), gh.astype(np.float64), {})
  assert _rel(gz, gz_ref) < tol, _rel(gz, gz_ref)
  # a model built without vjp=True refuses loudly
  plain = build_config("two_layer_syn", precision=precision, ctx=gpu_ctx)
  plain.load_weights(wts)
  with pytest.raises(RuntimeError):
    plain.synthesis_vjp(y, gx)


@pytest.mark.gpu
def test_vjp_error_behaviour(gpu_ctx):
  """Loud failures, never a silent fallback: no backward plan without sntc_model_enable_vjp, none for res_type='d2s', shapes
  are validated, an empty batch is a no-op; device-resident tensors give the same gradient as host tensors."""
  from shallow_ntc_b200 import Model, SntcError, _lib
  from shallow_ntc_b200._lib import lib
  from shallow_ntc_b200.tensors import as_tensor
  elic = dict(cls="ElicAnalysis", channels=(192, 192, 192, 64))
  syn = dict(cls="TwoLayerResSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5), activation_type="igdn")
  rng = np.random.default_rng(4)

  def build(syn_cfg, **kw):
    m = Model(dict(analysis=elic, synthesis=syn_cfg), precision="fp32", ctx=gpu_ctx, **kw)
    m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls="TwoLayerResSynthesis"))
    return m
  y = rng.standard_normal((1, 2, 3, 64)).astype(np.float32)
  g = rng.standard_normal((1, 32, 48, 3)).astype(np.float32)
  # C level: a finalized model without the plan
  plain = build(syn)
  plain._ensure_native()
  gin = np.empty_like(y)
  rc = lib.sntc_synthesis_vjp(plain._native.handle, as_tensor(y).byref(), as_tensor(g).byref(), as_tensor(gin).byref(), None, None)
  assert rc == -3 and b"sntc_model_enable_vjp" in lib.sntc_last_error()        # SNTC_E_STATE
  assert lib.sntc_model_enable_vjp(plain._native.handle, 1) == -3                  # too late: the host weights are gone
  # d2s has no backward
  d2s = build(dict(syn, res_type="d2s"), vjp=True)
  with pytest.raises(SntcError, match="d2s"):
    d2s.synthesis_vjp(y, g)
  m = build(syn, vjp=True)
  with pytest.raises(SntcError, match="grad_out"):
    m.synthesis_vjp(y, np.ascontiguousarray(g[:, :16]))
  with pytest.raises(SntcError, match="channel"):
    m.synthesis_vjp(np.ascontiguousarray(y[..., :32]), g)
  empty = m.synthesis_vjp(y[:0], g[:0])
  assert empty.shape == (0, 2, 3, 64)
  host = m.synthesis_vjp(y, g)
  dev = m.synthesis_vjp(gpu_ctx.to_device(y), gpu_ctx.to_device(g))
  assert np.array_equal(host, dev.to_host())
