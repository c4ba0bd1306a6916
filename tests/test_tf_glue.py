"""not-gpu: the reference-side binding (shallow_ntc_b200/tf_glue.py) against stand-ins that are duck-typed like the Keras /
tensorflow-compression objects of the reference (TensorFlow itself cannot be installed here)."""
import functools
import types

import numpy as np
import pytest

from shallow_ntc_b200 import tf_glue, build_config, CONFIGS, Model
from shallow_ntc_b200 import transforms as b200_transforms


class _Var:
  """tf.Variable stand-in: only .numpy() is used."""
  def __init__(self, a):
    self._a = np.asarray(a, dtype=np.float32)

  def numpy(self):
    return self._a


class _Gdn:
  def __init__(self, rng, c):
    self.beta, self.gamma = _Var(1 + rng.random(c)), _Var(rng.random((c, c)))


class Conv2DTranspose:
  def __init__(self, rng, k, cin, cout, use_bias=True, activation=None):
    self.kernel, self.use_bias, self.activation = _Var(rng.standard_normal((k, k, cout, cin))), use_bias, activation
    self.bias = _Var(rng.standard_normal(cout)) if use_bias else None


class SignalConv2D:
  def __init__(self, rng, k, cin, cout, activation=None):
    self.kernel, self.bias, self.use_bias, self.activation = _Var(rng.standard_normal((k, k, cin, cout))), _Var(rng.standard_normal(cout)), True, activation


def _fake(name, **attrs):
  """An instance of a class NAMED like the reference's (export recognises transforms by class name)."""
  obj = type(name, (), {})()
  for k, v in attrs.items():
    setattr(obj, k, v)
  return obj


def fake_reference_model(name, rng):
  """An object shaped like mshyper.models.Model / factorized.models.Model after _init_transforms, for config `name`."""
  cfg = CONFIGS[name.split(":")[0]]
  syn = dict(cfg["synthesis"])
  C = 256 if name == "bls2017" else 320
  cls = syn["cls"]
  if cls == "JPEGLikeSynthesis":
    s = _fake(cls, conv=Conv2DTranspose(rng, 18, C, 3))
  elif cls == "TwoLayerResSynthesis":
    act = _Gdn(rng, 12)
    s = _fake(cls, base_conv=Conv2DTranspose(rng, 13, C, 12, activation=act), res=Conv2DTranspose(rng, 13, C, 12), activation=act,
              out_conv=Conv2DTranspose(rng, 5, 12, 3))
  elif cls == "TwoLayerSynthesis":
    s = _fake(cls, conv1=Conv2DTranspose(rng, 13, C, 12, activation=_Gdn(rng, 12)), conv2=Conv2DTranspose(rng, 5, 12, 3))
  elif cls == "MBT2018Synthesis":
    chans = [C, 192, 192, 192, 3]
    s = _fake(cls, layers=[SignalConv2D(rng, 5, chans[i], chans[i + 1], activation=_Gdn(rng, 192) if i < 3 else None) for i in range(4)])
  elif cls == "BLS2017Synthesis":
    s = _fake(cls, layers=[SignalConv2D(rng, 5, C, C, _Gdn(rng, C)), SignalConv2D(rng, 5, C, C, _Gdn(rng, C)), SignalConv2D(rng, 9, C, 3)])
  m = types.SimpleNamespace(_synthesis=s, _bottleneck_size=C, _hyper_synthesis=None, _prior=None)
  if name != "bls2017":
    relu = lambda x: x
    m._hyper_synthesis = _fake("HyperSynthesis", layers=[Conv2DTranspose(rng, 5, C, C, activation=relu), Conv2DTranspose(rng, 5, C, 480, activation=relu),
                                                         Conv2DTranspose(rng, 3, 480, 640)])
    f = (1, 3, 3, 3, 1)
    base = types.SimpleNamespace(_matrices=[_Var(rng.standard_normal((C, f[i + 1], f[i]))) for i in range(4)],
                                 _biases=[_Var(rng.standard_normal((C, f[i + 1], 1))) for i in range(4)],
                                 _factors=[_Var(rng.standard_normal((C, f[i + 1], 1))) for i in range(3)])
    m._prior = types.SimpleNamespace(base=base)
  return m


@pytest.mark.parametrize("name", ["jpegl", "two_layer_syn", "two_layer_syn2", "mbt2018", "bls2017"])
def test_export_weights_names_and_shapes_are_what_the_model_expects(name):
  rng = np.random.default_rng(0)
  ref = fake_reference_model(name, rng)
  w = tf_glue.export_weights(ref)
  want = build_config(name, prior=name != "bls2017").variable_shapes()
  assert {k: v.shape for k, v in w.items()} == {k: tuple(v) for k, v in want.items()}
  assert all(v.dtype == np.float32 and v.flags.c_contiguous for v in w.values())
  # values are the layer's own, not re-ordered
  if name == "two_layer_syn":
    assert np.array_equal(w["synthesis.res.kernel"], ref._synthesis.res.kernel.numpy())
    assert np.array_equal(w["synthesis.activation.gamma"], ref._synthesis.activation.gamma.numpy())
    assert np.array_equal(w["prior.factor_2"], ref._prior.base._factors[2].numpy())


def test_cnn_synthesis_exports_its_single_shared_activation():
  rng = np.random.default_rng(1)
  act = _Gdn(rng, 192)
  chans = [320, 192, 192, 192, 3]
  layer = _fake("CNNSynthesis", layers=[Conv2DTranspose(rng, 5, chans[i], chans[i + 1], activation=act if i < 3 else None) for i in range(4)])
  w = tf_glue.export_transform_weights(layer, "synthesis")
  want = b200_transforms.CNNSynthesis(192, activation_type="igdn").variable_shapes(320)
  assert {k: v.shape for k, v in w.items()} == {k: tuple(v) for k, v in want.items()}


def test_profile_wrappers_are_unwrapped():
  """Model(profile=True) replaces the transforms by with_timing(tf.function(layer)) (mshyper/models.py:142-146)."""
  rng = np.random.default_rng(2)
  layer = _fake("JPEGLikeSynthesis", conv=Conv2DTranspose(rng, 18, 320, 3))
  tf_function = types.SimpleNamespace(python_function=layer)            # what tf.function(layer) exposes

  @functools.wraps(tf_function)
  def timed(*a, **k):
    return None
  timed.__wrapped__ = tf_function
  w = tf_glue.export_transform_weights(timed, "synthesis")
  assert set(w) == {"synthesis.conv.kernel", "synthesis.conv.bias"}


def test_d2s_residual_is_exported_and_unknown_classes_are_refused():
  """res_type="d2s" (common/transforms.py:339-348): Sequential [Lambda, Conv2D, Lambda, Conv2D, Lambda] -> res.conv_0 / res.conv_1."""
  rng = np.random.default_rng(3)
  conv2d = lambda cin, cout: types.SimpleNamespace(kernel=_Var(rng.standard_normal((1, 1, cin, cout))), bias=_Var(rng.standard_normal(cout)), use_bias=True)
  lam = lambda: types.SimpleNamespace()
  seq = _fake("Sequential", layers=[lam(), conv2d(80, 192), lam(), conv2d(48, 48), lam()])
  d2s = _fake("TwoLayerResSynthesis", base_conv=Conv2DTranspose(rng, 13, 320, 12), res=seq, activation=None,
              out_conv=Conv2DTranspose(rng, 5, 12, 3))
  w = tf_glue.export_transform_weights(d2s, "synthesis")
  assert w["synthesis.res.conv_0.kernel"].shape == (1, 1, 80, 192) and w["synthesis.res.conv_1.bias"].shape == (48,)
  from shallow_ntc_b200 import transforms as T
  want = T.TwoLayerResSynthesis(activation_type=None, res_type="d2s").variable_shapes(320)
  assert {k: v.shape for k, v in w.items()} == {k: tuple(v) for k, v in want.items()}
  broken = _fake("TwoLayerResSynthesis", base_conv=Conv2DTranspose(rng, 13, 320, 12), res=_fake("Sequential", layers=[]), activation=None,
                 out_conv=Conv2DTranspose(rng, 5, 12, 3))
  with pytest.raises(NotImplementedError):
    tf_glue.export_transform_weights(broken, "synthesis")
  with pytest.raises(NotImplementedError):
    tf_glue.export_transform_weights(_fake("ElicSynthesis"), "synthesis")


def test_patch_registry_swaps_only_the_decoder_side_classes():
  class Builder(dict):
    def build(self, name, **kw):
      return self[name](**kw)
  ref_classes = {n: type(n, (), {}) for n in tf_glue.DECODER_CLASSES + ("ElicAnalysis", "HyperAnalysis", "CNNAnalysis")}
  mod = types.SimpleNamespace(class_builder=Builder(ref_classes))
  old = tf_glue.patch_registry(mod)
  assert set(old) == set(tf_glue.DECODER_CLASSES)
  assert mod.class_builder["ElicAnalysis"] is ref_classes["ElicAnalysis"]
  t = mod.class_builder.build("TwoLayerResSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5), activation_type="igdn", res_type="conv")
  assert isinstance(t, b200_transforms.TwoLayerResSynthesis) and t.upsample == 16
  h = mod.class_builder.build("HyperSynthesis", bottleneck_size=320)
  assert isinstance(h, b200_transforms.HyperSynthesis)


def test_b200_model_from_config_json_and_restored_model():
  rng = np.random.default_rng(4)
  ref = fake_reference_model("two_layer_syn", rng)
  model_config = dict(transform_config={k: dict(v) for k, v in CONFIGS["two_layer_syn"].items()}, rd_lambda=0.08)
  m = tf_glue.b200_model_from(ref, model_config, precision="tc")
  assert isinstance(m, Model) and m.precision == "tc" and m.index_rounding == "trunc" and m.latent_channels == 320
  assert set(m._weights) == set(m.variable_shapes())
  fm = tf_glue.b200_model_from(fake_reference_model("bls2017", rng), dict(transform_config={k: dict(v) for k, v in CONFIGS["bls2017"].items()}))
  assert not fm.hyperprior and fm.latent_channels == 256 and set(fm._weights) == set(fm.variable_shapes())


def test_dlpack_export_is_zero_copy_and_owns_its_producer():
  """tensors.to_dlpack: the capsule tf.experimental.dlpack.from_dlpack would take; consumed here by torch (tests only)."""
  import gc
  import torch
  from shallow_ntc_b200.tensors import to_dlpack, as_tensor, _DL_LIVE
  n0 = len(_DL_LIVE)
  a = np.arange(2 * 3 * 4 * 3, dtype=np.uint8).reshape(2, 3, 4, 3)
  t = torch.utils.dlpack.from_dlpack(to_dlpack(a))
  assert t.dtype == torch.uint8 and tuple(t.shape) == a.shape and np.array_equal(t.numpy(), a)
  a[1, 2, 3, 2] = 200
  assert int(t[1, 2, 3, 2]) == 200                       # same memory
  assert len(_DL_LIVE) == n0 + 1
  del t
  gc.collect()
  assert len(_DL_LIVE) == n0                              # consumer's deleter released the producer
  cap = to_dlpack(np.zeros((1, 2, 2, 4), np.float32))    # never consumed: the capsule destructor releases it
  r = as_tensor(cap)
  assert r.shape == (1, 2, 2, 4)
  del r, cap
  gc.collect()
  assert len(_DL_LIVE) == n0


def test_differentiable_wraps_forward_and_vjp_in_a_custom_gradient():
  """tf_glue.differentiable: forward = layer(x), gradient = layer.vjp(x, dy) (itinf_train_step, mshyper/models.py:401-408),
  with a stand-in for the two TensorFlow entry points it touches."""
  seen = {}

  class FakeTf:
    @staticmethod
    def custom_gradient(fn):
      def run(x):
        y, g = fn(x)
        seen["grad"] = g
        return y
      return run

    @staticmethod
    def convert_to_tensor(a):
      return np.asarray(a)

  class Layer:
    def __call__(self, x, training=None):
      return 2.0 * x

    def vjp(self, x, dy):
      seen["vjp_x"] = x
      return 2.0 * dy

  f = tf_glue.differentiable(Layer(), tf=FakeTf)
  x = np.arange(6, dtype=np.float32).reshape(1, 1, 2, 3)
  y = f(x, training=True)
  assert np.array_equal(y, 2 * x)
  dy = np.ones_like(x)
  assert np.array_equal(seen["grad"](dy), 2 * dy) and np.array_equal(seen["vjp_x"], x)
  assert isinstance(f.__wrapped__, Layer)
