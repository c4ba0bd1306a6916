"""Drop-in registry of the decoder-side transform classes of ``common/transforms.py``.

Same class names, same constructor kwargs, same ``class_builder.build(name, **kwargs)`` mechanism
(``common/transforms.py:380-393``, ``common/utils.py:58-71``), same call convention
(``layer(x)`` / ``layer(x, training=False)`` -> NHWC float32), but the arithmetic runs in libsntc
(hand-written sm_100a CUDA) instead of TensorFlow.  Encoder-side classes are not provided.

Variable layouts are the reference's: Keras ``Conv2DTranspose.kernel`` [kh,kw,Cout,Cin];
tfc ``SignalConv2D.kernel`` [kh,kw,Cin,Cout]; GDN ``beta`` [C], ``gamma`` [C,C] (in,out) -- always the
*effective* (de-reparameterised) values, i.e. ``layer.kernel`` / ``layer.beta`` / ``layer.gamma``.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from ._lib import lib, check, TransformDesc, ModelDesc

_ACT = {None: _lib.ACT_NONE, "relu": _lib.ACT_RELU, "leaky_relu": _lib.ACT_LEAKY_RELU, "lrelu": _lib.ACT_LEAKY_RELU,
        "igdn": _lib.ACT_IGDN1, "igdn1": _lib.ACT_IGDN1, "gdn": _lib.ACT_GDN1, "gdn1": _lib.ACT_GDN1}


def activation_code(activation):
  """get_activation_op (common/transforms.py:66-78) restricted to what the decode path supports."""
  key = activation.lower() if isinstance(activation, str) else activation
  if key not in _ACT:
    raise NotImplementedError(f"activation {activation!r} is not supported on the B200 decode path")
  return _ACT[key]


class ClassBuilder(dict):
  """common/utils.py:58-71."""

  def build(self, class_name, **kwargs):
    cls = self[class_name]
    return cls(**kwargs)


def _keras_convt(prefix, name, k, cin, cout, bias=True):
  v = {f"{prefix}.{name}.kernel": (k, k, cout, cin)}
  if bias:
    v[f"{prefix}.{name}.bias"] = (cout,)
  return v


def _signal_conv(prefix, name, k, cin, cout):
  return {f"{prefix}.{name}.kernel": (k, k, cin, cout), f"{prefix}.{name}.bias": (cout,)}


def _gdn(prefix, name, c):
  return {f"{prefix}.{name}.beta": (c,), f"{prefix}.{name}.gamma": (c, c)}


def _keras_pad(k, s):
  return max(k - s, 0) // 2          # Conv2DTranspose(padding="SAME"): pad_before of the SAME correlation it is the gradient of


def _tfc_pad(k):
  return (k - 1) // 2                # SignalConv2D(padding="same_zeros")


class Transform:
  """Base of the shim classes.  A transform is lazily bound to a hyper-only / synthesis-only libsntc
  model the first time it is called on its own; inside ``models.Model`` the fused decode is used."""
  role = "synthesis"
  upsample = 1

  def __init__(self):
    self.in_channels = None
    self._weights = {}
    self._model = None
    self._ctx = None
    self.precision = "fp32"

  # -- description ------------------------------------------------------------------------------
  def desc(self, in_channels: int) -> TransformDesc:
    raise NotImplementedError

  def variable_shapes(self, in_channels: int) -> dict:
    raise NotImplementedError

  @property
  def out_channels(self):
    raise NotImplementedError

  def conv_chain(self):
    """[(kernel, stride, pad), ...] of the transposed convolutions in forward order: out[o] += in[n] * W[a], o = n*s + a - pad.
    Pointwise stages do not appear.  Drives the halo computation of the intra-frame band split (tiling.py)."""
    raise NotImplementedError

  def count_params(self, in_channels: int | None = None) -> int:
    cin = in_channels if in_channels is not None else self.in_channels
    return int(sum(int(np.prod(s)) for s in self.variable_shapes(cin).values()))

  # -- weights ----------------------------------------------------------------------------------
  def load_weights(self, weights: dict):
    """weights: name -> float32 array, names as in ``variable_shapes`` (prefix ``synthesis.`` /
    ``hyper_synthesis.``).  Mirrors restoring the Keras variables from a checkpoint."""
    self._weights = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items() if k.startswith(self.role + ".")}
    self._model = None

  # -- standalone call --------------------------------------------------------------------------
  def _bind(self, cin, device=0, vjp=False):
    from .tensors import Context
    from .models import _create_model
    if self.in_channels is None:
      self.in_channels = int(cin)
    if self._ctx is None:
      self._ctx = Context(device)
    hyper = self.desc(self.in_channels) if self.role == "hyper_synthesis" else TransformDesc(kind=_lib.T_NONE)
    syn = self.desc(self.in_channels) if self.role == "synthesis" else TransformDesc(kind=_lib.T_NONE)
    self._model = _create_model(self._ctx, hyper, syn, self._weights, precision=self.precision, vjp=vjp)
    self._has_vjp = vjp

  def __call__(self, x, training=None):
    from .tensors import as_tensor, empty_like_kind
    xr = as_tensor(x)
    shp = xr.shape
    if self._model is None:
      self._bind(shp[-1])
    out = empty_like_kind(self._ctx, x if not hasattr(x, "__dlpack__") or isinstance(x, np.ndarray) else None,
                          (shp[0], shp[1] * self.upsample, shp[2] * self.upsample, self.out_channels), np.float32)
    outr = as_tensor(out)
    fn = lib.sntc_hyper_synthesis if self.role == "hyper_synthesis" else lib.sntc_synthesis
    check(fn(self._model.handle, xr.byref(), outr.byref(), None))
    self._ctx.sync()
    return out

  def vjp(self, x, grad_out):
    """J(x)^T grad_out: what ``tape.gradient`` propagates through ``layer(x)`` (the iterative-inference step of the reference,
    ``mshyper/models.py:401-408``).  ``grad_out`` has the shape of ``layer(x)``; returns an array of the shape of ``x``.
    The body of a ``tf.custom_gradient`` around ``__call__``."""
    from .tensors import as_tensor, empty_like_kind
    xr = as_tensor(x)
    if self._model is None or not getattr(self, "_has_vjp", False):   # the backward layers are packed on first use
      self._bind(xr.shape[-1], vjp=True)
    gin = empty_like_kind(self._ctx, x if not hasattr(x, "__dlpack__") or isinstance(x, np.ndarray) else None, tuple(xr.shape), np.float32)
    fn = lib.sntc_hyper_synthesis_vjp if self.role == "hyper_synthesis" else lib.sntc_synthesis_vjp
    check(fn(self._model.handle, xr.byref(), as_tensor(grad_out).byref(), as_tensor(gin).byref(), None, None))
    self._ctx.sync()
    return gin


# ---------------------------------------------------------------------------------------------
class HyperSynthesis(Transform):
  """common/transforms.py:222-232."""
  role = "hyper_synthesis"
  upsample = 4

  def __init__(self, bottleneck_size, activation_type="relu"):
    super().__init__()
    self.bottleneck_size = int(bottleneck_size)
    self.activation_type = activation_type

  def conv_chain(self):
    return [(5, 2, _keras_pad(5, 2)), (5, 2, _keras_pad(5, 2)), (3, 1, _keras_pad(3, 1))]

  @property
  def out_channels(self):
    return self.bottleneck_size * 2

  def desc(self, in_channels):
    d = TransformDesc(kind=_lib.T_HYPER_SYNTHESIS, in_channels=in_channels, activation=activation_code(self.activation_type))
    d.channels[0] = self.bottleneck_size
    return d

  def variable_shapes(self, in_channels):
    c = self.bottleneck_size
    v = {}
    v.update(_keras_convt(self.role, "layer_0", 5, in_channels, c))
    v.update(_keras_convt(self.role, "layer_1", 5, c, int(c * 1.5)))
    v.update(_keras_convt(self.role, "layer_2", 3, int(c * 1.5), c * 2))
    return v


class JPEGLikeHyperSynthesis(Transform):
  """common/transforms.py:364-377."""
  role = "hyper_synthesis"
  upsample = 4

  def __init__(self, bottleneck_size, kernel_size=6):
    super().__init__()
    self.bottleneck_size, self.kernel_size = int(bottleneck_size), int(kernel_size)

  def conv_chain(self):
    return [(self.kernel_size, 4, _keras_pad(self.kernel_size, 4))]

  @property
  def out_channels(self):
    return self.bottleneck_size * 2

  def desc(self, in_channels):
    d = TransformDesc(kind=_lib.T_JPEG_LIKE_HYPER, in_channels=in_channels)
    d.channels[0] = self.bottleneck_size
    d.kernel_sizes[0] = self.kernel_size
    return d

  def variable_shapes(self, in_channels):
    return _keras_convt(self.role, "conv", self.kernel_size, in_channels, self.bottleneck_size * 2)


class HyperSynthesisSmall(Transform):
  """common/transforms.py:250-262."""
  role = "hyper_synthesis"
  upsample = 2

  def __init__(self, bottleneck_size):
    super().__init__()
    self.bottleneck_size = int(bottleneck_size)

  def conv_chain(self):
    return [(5, 2, _tfc_pad(5)), (3, 1, _tfc_pad(3))]

  @property
  def out_channels(self):
    return self.bottleneck_size * 2

  def desc(self, in_channels):
    d = TransformDesc(kind=_lib.T_HYPER_SMALL, in_channels=in_channels)
    d.channels[0] = self.bottleneck_size
    return d

  def variable_shapes(self, in_channels):
    c = self.bottleneck_size
    v = _signal_conv(self.role, "layer_0", 5, in_channels, int(c * 1.5))
    v.update(_signal_conv(self.role, "layer_1", 3, int(c * 1.5), c * 2))
    return v


class JPEGLikeSynthesis(Transform):
  """common/transforms.py:265-295."""

  def __init__(self, output_channels=3, kernel_size=16, strides=16, padding="SAME", use_bias=True, use_offset=False):
    super().__init__()
    if padding != "SAME":
      raise NotImplementedError("only padding='SAME' is used by the reference configs")
    self.output_channels, self.kernel_size, self.strides = int(output_channels), int(kernel_size), int(strides)
    self.use_bias, self.use_offset = bool(use_bias), bool(use_offset)
    self.upsample = self.strides

  def conv_chain(self):
    return [(self.kernel_size, self.strides, _keras_pad(self.kernel_size, self.strides))]

  @property
  def out_channels(self):
    return self.output_channels

  def desc(self, in_channels):
    d = TransformDesc(kind=_lib.T_JPEG_LIKE_SYNTHESIS, in_channels=in_channels, use_bias=int(self.use_bias), use_offset=int(self.use_offset))
    d.channels[0] = self.output_channels
    d.kernel_sizes[0] = self.kernel_size
    d.strides[0] = self.strides
    return d

  def variable_shapes(self, in_channels):
    return _keras_convt(self.role, "conv", self.kernel_size, in_channels + int(self.use_offset), self.output_channels, self.use_bias)


class TwoLayerSynthesis(Transform):
  """common/transforms.py:298-317."""

  def __init__(self, channels=(24, 3), strides=(8, 2), kernel_sizes=(13, 5), activation_type="igdn"):
    super().__init__()
    self.channels, self.strides, self.kernel_sizes = tuple(channels), tuple(strides), tuple(kernel_sizes)
    self.activation_type = activation_type
    self.upsample = self.strides[0] * self.strides[1]

  def conv_chain(self):
    return [(k, s, _keras_pad(k, s)) for k, s in zip(self.kernel_sizes, self.strides)]

  @property
  def out_channels(self):
    return self.channels[1]

  def _kind(self):
    return _lib.T_TWO_LAYER

  def desc(self, in_channels):
    d = TransformDesc(kind=self._kind(), in_channels=in_channels, activation=activation_code(self.activation_type))
    for i in range(2):
      d.channels[i], d.strides[i], d.kernel_sizes[i] = self.channels[i], self.strides[i], self.kernel_sizes[i]
    return d

  def variable_shapes(self, in_channels):
    v = _keras_convt(self.role, "conv1", self.kernel_sizes[0], in_channels, self.channels[0])
    if activation_code(self.activation_type) in (_lib.ACT_IGDN1, _lib.ACT_GDN1):
      v.update(_gdn(self.role, "activation", self.channels[0]))
    v.update(_keras_convt(self.role, "conv2", self.kernel_sizes[1], self.channels[0], self.channels[1]))
    return v


class TwoLayerResSynthesis(TwoLayerSynthesis):
  """common/transforms.py:320-361.  ``res_type="conv"`` (every shipped config): the residual is a second transposed conv.
  ``res_type="d2s"`` (:339-348): the residual is depth_to_space(2) -> Conv2D 1x1 (192) + leaky_relu -> depth_to_space(2) ->
  Conv2D 1x1 (4 * channels[0]) + leaky_relu -> depth_to_space(2); its variables are ``res.conv_0`` / ``res.conv_1`` (kernel
  ``[1, 1, Cin, Cout]``, bias).  ``'leaky_relu'`` is taken as ``tf.nn.leaky_relu`` (alpha 0.2, what the string means in the Keras
  versions that know it; Keras 2.10 does not -- assumption A11 in DESIGN.md section 5)."""

  D2S_WIDTH = 192

  def __init__(self, channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5), activation_type="igdn", res_type="conv"):
    super().__init__(channels, strides, kernel_sizes, activation_type)
    if res_type not in ("conv", "d2s"):
      raise NotImplementedError(f"res_type={res_type!r}")     # the reference raises NotImplementedError too (:349-350)
    if res_type == "d2s" and self.strides[0] != 8:
      raise ValueError("res_type='d2s' upsamples by three depth_to_space(2) steps: strides[0] must be 8")
    self.res_type = res_type

  def _kind(self):
    return _lib.T_TWO_LAYER_RES if self.res_type == "conv" else _lib.T_TWO_LAYER_RES_D2S

  def variable_shapes(self, in_channels):
    v = _keras_convt(self.role, "base_conv", self.kernel_sizes[0], in_channels, self.channels[0])
    if self.res_type == "conv":
      v.update(_keras_convt(self.role, "res", self.kernel_sizes[0], in_channels, self.channels[0]))
    else:
      if in_channels % 16:
        raise ValueError("res_type='d2s' needs a latent channel count that is a multiple of 16")
      w = self.D2S_WIDTH
      v.update({f"{self.role}.res.conv_0.kernel": (1, 1, in_channels // 4, w), f"{self.role}.res.conv_0.bias": (w,),
                f"{self.role}.res.conv_1.kernel": (1, 1, w // 4, 4 * self.channels[0]), f"{self.role}.res.conv_1.bias": (4 * self.channels[0],)})
    if activation_code(self.activation_type) in (_lib.ACT_IGDN1, _lib.ACT_GDN1):
      v.update(_gdn(self.role, "activation", self.channels[0]))
    v.update(_keras_convt(self.role, "out_conv", self.kernel_sizes[1], self.channels[0], self.channels[1]))
    return v


class MBT2018Synthesis(Transform):
  """common/transforms.py:158-175.  Its activations are ``tfc.GDN(inverse=True)`` with the library defaults, which in
  tensorflow-compression 2.x are alpha_parameter = epsilon_parameter = 1: the same function as ``GDN1(inverse=True)``
  (x * (beta + |x| @ gamma)).  ``gdn_form="classic"`` (an extension, not a reference kwarg) selects the original
  x * sqrt(beta + x^2 @ gamma) for weights trained with ``tfc.GDN(alpha_parameter=2, epsilon_parameter=.5)``."""

  def __init__(self, channels_base, n_layers=4, output_channels=3, *, gdn_form="gdn1"):
    super().__init__()
    if gdn_form not in ("gdn1", "classic"):
      raise ValueError(f"gdn_form={gdn_form!r}")
    self.gdn_form = gdn_form
    self.channels_base, self.n_layers = int(channels_base), int(n_layers)
    self.output_channels = int(output_channels) if output_channels is not None else int(channels_base)
    self.upsample = 2 ** self.n_layers

  def conv_chain(self):
    return [(5, 2, _tfc_pad(5))] * self.n_layers

  @property
  def out_channels(self):
    return self.output_channels

  def desc(self, in_channels):
    d = TransformDesc(kind=_lib.T_MBT2018, in_channels=in_channels, n_layers=self.n_layers,
                      activation=_lib.ACT_IGDN_CLASSIC if self.gdn_form == "classic" else _lib.ACT_NONE)
    d.channels[0], d.channels[1] = self.channels_base, self.output_channels
    return d

  def variable_shapes(self, in_channels):
    v, cin = {}, in_channels
    for i in range(self.n_layers):
      last = i + 1 == self.n_layers
      v.update(_signal_conv(self.role, f"layer_{i}", 5, cin, self.output_channels if last else self.channels_base))
      if not last:
        v.update(_gdn(self.role, f"igdn_{i}", self.channels_base))
      cin = self.channels_base
    return v


class BLS2017Synthesis(Transform):
  """common/transforms.py:115-134."""
  upsample = 16

  def __init__(self, num_filters):
    super().__init__()
    self.num_filters = int(num_filters)

  def conv_chain(self):
    return [(5, 2, _tfc_pad(5)), (5, 2, _tfc_pad(5)), (9, 4, _tfc_pad(9))]

  @property
  def out_channels(self):
    return 3

  def desc(self, in_channels):
    d = TransformDesc(kind=_lib.T_BLS2017, in_channels=in_channels)
    d.channels[0], d.channels[1] = self.num_filters, 3
    return d

  def variable_shapes(self, in_channels):
    c = self.num_filters
    v = _signal_conv(self.role, "layer_0", 5, in_channels, c)
    v.update(_gdn(self.role, "igdn_0", c))
    v.update(_signal_conv(self.role, "layer_1", 5, c, c))
    v.update(_gdn(self.role, "igdn_1", c))
    v.update(_signal_conv(self.role, "layer_2", 9, c, 3))
    return v


class CNNSynthesis(Transform):
  """common/transforms.py:195-206."""
  upsample = 16

  def __init__(self, channels_base, output_channels=3, activation_type="leaky_relu"):
    super().__init__()
    self.channels_base, self.output_channels = int(channels_base), int(output_channels)
    self.activation_type = activation_type

  def conv_chain(self):
    return [(5, 2, _keras_pad(5, 2))] * 4

  @property
  def out_channels(self):
    return self.output_channels

  def desc(self, in_channels):
    d = TransformDesc(kind=_lib.T_CNN, in_channels=in_channels, activation=activation_code(self.activation_type))
    d.channels[0], d.channels[1] = self.channels_base, self.output_channels
    return d

  def variable_shapes(self, in_channels):
    v, cin = {}, in_channels
    for i in range(4):
      v.update(_keras_convt(self.role, f"layer_{i}", 5, cin, self.output_channels if i == 3 else self.channels_base))
      cin = self.channels_base
    if activation_code(self.activation_type) in (_lib.ACT_IGDN1, _lib.ACT_GDN1):
      v.update(_gdn(self.role, "activation", self.channels_base))  # one shared activation object (:199-204)
    return v


classes = [
  BLS2017Synthesis, CNNSynthesis, HyperSynthesis, MBT2018Synthesis, HyperSynthesisSmall,
  JPEGLikeSynthesis, TwoLayerSynthesis, TwoLayerResSynthesis, JPEGLikeHyperSynthesis,
]
# Register the transform classes so they can be built from config dicts (common/transforms.py:392-393).
class_builder = ClassBuilder({cls.__name__: cls for cls in classes})
