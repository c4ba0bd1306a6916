"""Reference-side binding: what a maintainer of mandt-lab/shallow-ntc imports to run its decode on libsntc.

TensorFlow is imported lazily and only by the functions that touch ``tf.Tensor`` objects (this package has no TensorFlow
dependency; TF 2.10 is not installable in the build container, so everything here is exercised with stand-in objects in
``tests/test_tf_glue.py`` -- duck-typed exactly like the Keras / tensorflow-compression objects they replace).

  * ``export_weights(tf_model)``     Keras / tfc variables of a restored ``mshyper.models.Model`` / ``factorized.models.Model``
                                     (``common/eval_lib.py:11-53``) -> the name -> float32 array dict ``Model.load_weights`` takes
  * ``patch_registry(transforms)``   swaps the decoder-side classes of ``common.transforms.class_builder`` (``:380-393``)
  * ``b200_model_from(tf_model, model_config)``  one call: config dict + weights -> ``shallow_ntc_b200.Model``
  * ``from_tf`` / ``to_tf``          zero-copy tensor hand-off through DLPack capsules (``tf.experimental.dlpack``)
  * ``evaluate_like_reference``      ``Model.evaluate`` (``mshyper/models.py:415-433``) with the decode half on the GPU path
  * ``differentiable(layer)``        ``tf.custom_gradient`` around a B200 transform (forward ``layer(x)``, backward ``layer.vjp``):
                                     keeps the ``GradientTape`` of ``itinf_train_step`` (``mshyper/models.py:401-408``) working

Variable conventions (always the EFFECTIVE, de-reparameterised values): Keras ``Conv2DTranspose.kernel`` [kh,kw,Cout,Cin],
tfc ``SignalConv2D.kernel`` [kh,kw,Cin,Cout] (the spatial-domain kernel property, not the RDFT variable), ``GDN.beta`` [C],
``GDN.gamma`` [C_in, C_out] (the properties, not the reparameterised variables).
"""
from __future__ import annotations

import importlib

import numpy as np


def _tf():
  return importlib.import_module("tensorflow")


def _np(v):
  """tf.Variable / tf.Tensor / numpy -> float32 numpy."""
  if hasattr(v, "numpy"):
    v = v.numpy()
  return np.ascontiguousarray(np.asarray(v), dtype=np.float32)


def _unwrap(obj):
  """``Model(profile=True)`` wraps every transform as with_timing(tf.function(layer)) (mshyper/models.py:142-146): find the
  Keras object again (functools.wraps keeps ``__wrapped__``; tf.function keeps ``python_function``)."""
  seen = 0
  while seen < 8:
    nxt = getattr(obj, "__wrapped__", None) or getattr(obj, "python_function", None)
    if nxt is None or nxt is obj:
      break
    obj, seen = nxt, seen + 1
  return obj


def _conv(w, prefix, layer):
  w[prefix + ".kernel"] = _np(layer.kernel)
  if getattr(layer, "use_bias", True) and getattr(layer, "bias", None) is not None:
    w[prefix + ".bias"] = _np(layer.bias)


def _gdn(w, prefix, act):
  w[prefix + ".beta"] = _np(act.beta)
  w[prefix + ".gamma"] = _np(act.gamma)


def _is_gdn(act):
  return act is not None and hasattr(act, "beta") and hasattr(act, "gamma")


def export_transform_weights(layer, role: str) -> dict:
  """One decoder-side transform object of ``common/transforms.py`` -> {``<role>.<var>``: array}.  ``role`` is ``"synthesis"`` or
  ``"hyper_synthesis"``; the class is recognised by name, the variables are read through the attribute names the reference uses."""
  layer = _unwrap(layer)
  cls = type(layer).__name__
  w = {}
  if cls in ("JPEGLikeSynthesis", "JPEGLikeHyperSynthesis"):                      # :265-295, :364-377  (self.conv)
    _conv(w, f"{role}.conv", layer.conv)
  elif cls == "TwoLayerSynthesis":                                               # :298-317  (conv1 carries the activation)
    _conv(w, f"{role}.conv1", layer.conv1)
    if _is_gdn(layer.conv1.activation):
      _gdn(w, f"{role}.activation", layer.conv1.activation)
    _conv(w, f"{role}.conv2", layer.conv2)
  elif cls == "TwoLayerResSynthesis":                                            # :320-361
    _conv(w, f"{role}.base_conv", layer.base_conv)
    if type(layer.res).__name__ == "Conv2DTranspose" or hasattr(layer.res, "kernel"):   # res_type="conv"
      _conv(w, f"{role}.res", layer.res)
    else:                                                                          # res_type="d2s" (:339-348): Sequential of
      convs = [l for l in layer.res.layers if hasattr(l, "kernel")]               # Lambda(d2s), Conv2D, Lambda, Conv2D, Lambda
      if len(convs) != 2:
        raise NotImplementedError("TwoLayerResSynthesis: unknown residual branch (expected res_type 'conv' or 'd2s')")
      for i, sub in enumerate(convs):
        _conv(w, f"{role}.res.conv_{i}", sub)
    if _is_gdn(layer.activation):
      _gdn(w, f"{role}.activation", layer.activation)
    _conv(w, f"{role}.out_conv", layer.out_conv)
  elif cls in ("HyperSynthesis", "HyperSynthesisSmall", "CNNSynthesis", "MBT2018Synthesis", "BLS2017Synthesis"):   # Sequential stacks
    shared_done = False
    for i, sub in enumerate(layer.layers):
      _conv(w, f"{role}.layer_{i}", sub)
      act = getattr(sub, "activation", None)
      if _is_gdn(act):
        if cls == "CNNSynthesis":                                                # ONE activation object shared by layers 0-2 (:199-204)
          if not shared_done:
            _gdn(w, f"{role}.activation", act)
            shared_done = True
        else:                                                                    # tfc.GDN(name="igdn_i") / get_act() per layer
          _gdn(w, f"{role}.igdn_{i}", act)
  else:
    raise NotImplementedError(f"{cls} is not a decoder-side transform of the B200 path")
  return w


def export_prior_weights(prior, prefix="prior") -> dict:
  """tfc.NoisyDeepFactorized(batch_shape=(Cz,)) (mshyper/models.py:135) -> the RAW DeepFactorized variables
  ``prior.matrix_i`` [Cz,f_out,f_in], ``prior.bias_i`` [Cz,f_out,1], ``prior.factor_i`` [Cz,f_out,1] (softplus / tanh are applied
  when libsntc packs them).  tfc keeps them as ``prior.base._matrices / _biases / _factors``."""
  base = getattr(prior, "base", prior)
  w = {}
  for i, m in enumerate(base._matrices):
    w[f"{prefix}.matrix_{i}"] = _np(m)
  for i, b in enumerate(base._biases):
    w[f"{prefix}.bias_{i}"] = _np(b)
  for i, f in enumerate(base._factors):
    w[f"{prefix}.factor_{i}"] = _np(f)
  return w


def export_weights(tf_model, with_prior=True) -> dict:
  """A restored reference ``Model`` (``eval_lib.load_latest_ckpt``) -> weights for ``shallow_ntc_b200.Model.load_weights``."""
  w = export_transform_weights(tf_model._synthesis, "synthesis")
  hyp = getattr(tf_model, "_hyper_synthesis", None)
  if hyp is not None:
    w.update(export_transform_weights(hyp, "hyper_synthesis"))
    if with_prior and getattr(tf_model, "_prior", None) is not None:
      w.update(export_prior_weights(tf_model._prior))
  return w


DECODER_CLASSES = ("JPEGLikeSynthesis", "TwoLayerSynthesis", "TwoLayerResSynthesis", "HyperSynthesis", "JPEGLikeHyperSynthesis",
                   "HyperSynthesisSmall", "MBT2018Synthesis", "BLS2017Synthesis", "CNNSynthesis")


def patch_registry(transforms_module, names=DECODER_CLASSES):
  """``common.transforms.class_builder[name] = <B200 class>`` for the decoder-side classes: ``Model._init_transforms``
  (mshyper/models.py:111-131) then builds the B200 transforms from the unchanged config dicts.  Returns the replaced classes."""
  from . import transforms as b200
  old = {}
  for name in names:
    if name in transforms_module.class_builder:
      old[name] = transforms_module.class_builder[name]
      transforms_module.class_builder[name] = getattr(b200, name)
  return old


def b200_model_from(tf_model, model_config: dict, precision="tc", device=0, **kw):
  """config.json["model_config"] of a workdir (common/train_lib.py:325-336) + the restored TF model -> a loaded B200 model."""
  from .models import Model, FactorizedModel
  tcfg = model_config["transform_config"]
  hyper = getattr(tf_model, "_hyper_synthesis", None) is not None
  cls = Model if hyper else FactorizedModel
  m = cls(tcfg, precision=precision, device=device, prior=hyper, bottleneck_size=getattr(tf_model, "_bottleneck_size", None), **kw)
  m.load_weights(export_weights(tf_model, with_prior=hyper))
  return m


def from_tf(t):
  """tf.Tensor -> something ``as_tensor`` takes, zero-copy (DLPack capsule; device tensors stay on the device)."""
  return _tf().experimental.dlpack.to_dlpack(t)


def to_tf(arr):
  """DeviceArray / DeviceView / numpy produced by the decode -> tf.Tensor, zero-copy."""
  from .tensors import to_dlpack
  return _tf().experimental.dlpack.from_dlpack(to_dlpack(arr))


def differentiable(layer, tf=None):
  """``tf.custom_gradient`` around a B200 transform: forward ``layer(x)``, gradient ``layer.vjp(x, dy)`` (libsntc's
  ``sntc_synthesis_vjp`` / ``sntc_hyper_synthesis_vjp``).  With the two transforms of a model wrapped like this,
  ``itinf_train_step`` (``mshyper/models.py:401-408``: ``tape.gradient(loss, latent_rvs.trainable_variables)`` through
  ``frame_loss_given_latent_rvs(..., training=True)``) runs unchanged -- the rate terms, ``sga_round`` and the optimizer stay
  TensorFlow, the decoder forward and backward run on the GPU path.  Eager mode (what ``itinf.py`` uses); inside a
  ``tf.function`` wrap the returned callable in ``tf.py_function``.  ``tf`` is injectable for the stand-in tests."""
  tf = tf or _tf()

  @tf.custom_gradient
  def call(x):
    xn = _np(x)
    y = layer(xn)

    def grad(dy):
      return tf.convert_to_tensor(layer.vjp(xn, _np(dy)))
    return tf.convert_to_tensor(y), grad

  def wrapped(x, training=None):     # the reference calls self._synthesis(y, training=training) and self._hyper_synthesis(z)
    return call(x)
  wrapped.__wrapped__ = layer
  return wrapped


def symbols_of(tf_model, image):
  """The integer symbols the decode starts from, computed with the reference's own encoder side (out of scope here):
  ``infer_latent_rvs`` (mshyper/models.py:212-232) -> z_hat = round(z) (:253-259), mu from the reference hyper-synthesis,
  q = round(y - mu) (:278-283, latent_rvs_lib.py:95-102).  Returns numpy (z_hat | None, q) ready for ``Model.decompress``."""
  tf = _tf()
  rvs = tf_model.infer_latent_rvs(image)            # LatentRVCollection(uq=(UQLatentRV(z), UQLatentRV(y))); (rvs, timing) when profile=True
  if isinstance(rvs, tuple):
    rvs = rvs[0]
  if getattr(tf_model, "_hyper_synthesis", None) is None:
    return None, np.rint(_np(rvs.uq[0].loc))       # factorized/models.py:70-87: uq = (UQLatentRV(y),)
  z, y = rvs.uq[0].loc, rvs.uq[1].loc
  z_hat = tf.round(z)
  hs = tf_model._hyper_synthesis(z_hat)
  if isinstance(hs, tuple):          # profile=True: (result, seconds)
    hs = hs[0]
  mu = hs[..., :hs.shape[-1] // 2]
  return _np(z_hat), np.rint(_np(y) - _np(mu))


def evaluate_like_reference(tf_model, b200_model, images, rd_lambda=None):
  """``Model.evaluate`` (mshyper/models.py:415-433): one record per image, encoder side on TensorFlow, decode on libsntc."""
  from .eval_lib import evaluate_symbols
  for i, img in enumerate(images):
    img = img if len(img.shape) == 4 else img[None]
    z_hat, q = symbols_of(tf_model, img)
    H, W = int(img.shape[1]), int(img.shape[2])
    u8 = np.clip(np.rint((_np(img) + 0.5) * 255.0), 0, 255).astype(np.uint8)          # data_lib.floats_to_pixels
    for rec in evaluate_symbols(b200_model, z_hat, q, u8, (H, W), batch_size=1, rd_lambda=rd_lambda):
      rec["instance_id"] = i
      yield rec
