"""Decoder backward (vector-Jacobian products of the synthesis / hyper-synthesis transforms) by torch autograd in float64.
TEST INFRASTRUCTURE ONLY -- never imported by the product.

What it restates: the gradient that ``tf.GradientTape`` takes through ``self._synthesis`` / ``self._hyper_synthesis`` in the
iterative-inference step of the reference (``mshyper/models.py:401-408`` ``itinf_train_step``: ``tape.gradient(loss,
latent_rvs.trainable_variables)`` through ``frame_loss_given_latent_rvs`` ``:273, 297``).  The forward statements below are
written with ``torch.nn.functional.conv_transpose2d`` (independently of ``oracle/ntc_oracle.py``'s scatter loops; the two are
compared in ``tests/test_vjp.py``) for every registry class of ``common/transforms.py``; the backward is autograd's.

Pinning: same status as the forward oracle ("parity unpinned" at the TensorFlow boundary) -- the forward
functions are checked against T0, and a Jacobian-vector identity <g, J v> == <J^T g, v> is checked numerically.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

NP_DT = np.float64     # tools/vjp_bench.py times the float32 version as the CPU stand-in


def _t(a):
  return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=NP_DT)))


def _nchw(x):
  return _t(x).permute(0, 3, 1, 2).contiguous()


def _nhwc(x):
  return x.detach().permute(0, 2, 3, 1).contiguous().numpy()


def conv_t(x, kernel, bias, s, keras=True, p=None):
  """out[o] += in[n] * W[a], o = n*s + a - p, cropped to n_in*s (assumptions A1 / A2 of the oracle).
  kernel: Keras [kh,kw,Cout,Cin] or tfc [kh,kw,Cin,Cout]."""
  k = kernel.shape[0]
  w = _t(kernel)
  w = w.permute(3, 2, 0, 1) if keras else w.permute(2, 3, 0, 1)      # torch: [Cin, Cout, kh, kw]
  if p is None:
    p = max(k - s, 0) // 2 if keras else (k - 1) // 2
  nh, nw = x.shape[2] * s, x.shape[3] * s
  y = F.conv_transpose2d(x, w.contiguous(), _t(bias) if bias is not None else None, stride=s)
  return y[:, :, p:p + nh, p:p + nw]


def gdn1(x, beta, gamma, inverse=True):
  """common/transforms.py:27-63: norm[j] = beta[j] + sum_i |x[i]| gamma[i, j]."""
  norm = torch.einsum("bihw,ij->bjhw", x.abs(), _t(gamma)) + _t(beta)[None, :, None, None]
  return x * norm if inverse else x / norm


def gdn_classic(x, beta, gamma, inverse=True):
  norm = torch.sqrt(torch.einsum("bihw,ij->bjhw", x * x, _t(gamma)) + _t(beta)[None, :, None, None])
  return x * norm if inverse else x / norm


def activation(x, act, wts=None, prefix=None):
  """common/transforms.py:66-78 get_activation_op."""
  if act is None:
    return x
  a = act.lower()
  if a == "relu":
    return F.relu(x)
  if a in ("leaky_relu", "lrelu"):
    return F.leaky_relu(x, 0.2)
  if a in ("igdn", "igdn1"):
    return gdn1(x, wts[prefix + ".beta"], wts[prefix + ".gamma"], True)
  if a in ("gdn", "gdn1"):
    return gdn1(x, wts[prefix + ".beta"], wts[prefix + ".gamma"], False)
  raise NotImplementedError(act)


def forward(cls, wts, x, kwargs=None):
  """x: NCHW float64 tensor.  Every decoder-side class of common/transforms.py (lines cited in oracle/ntc_oracle.py)."""
  kw = dict(kwargs or {})
  g = lambda n: wts[n]
  if cls == "HyperSynthesis":                                    # :222-232
    act = kw.get("activation_type", "relu")
    p = "hyper_synthesis"
    x = activation(conv_t(x, g(f"{p}.layer_0.kernel"), g(f"{p}.layer_0.bias"), 2), act)
    x = activation(conv_t(x, g(f"{p}.layer_1.kernel"), g(f"{p}.layer_1.bias"), 2), act)
    return conv_t(x, g(f"{p}.layer_2.kernel"), g(f"{p}.layer_2.bias"), 1)
  if cls == "JPEGLikeHyperSynthesis":                            # :364-377
    return conv_t(x, g("hyper_synthesis.conv.kernel"), g("hyper_synthesis.conv.bias"), 4)
  if cls == "HyperSynthesisSmall":                               # :250-262
    p = "hyper_synthesis"
    x = F.relu(conv_t(x, g(f"{p}.layer_0.kernel"), g(f"{p}.layer_0.bias"), 2, keras=False))
    return conv_t(x, g(f"{p}.layer_1.kernel"), g(f"{p}.layer_1.bias"), 1, keras=False)
  p = "synthesis"
  if cls == "JPEGLikeSynthesis":                                 # :265-295
    if kw.get("use_offset", False):
      x = torch.cat([x, torch.ones_like(x[:, :1])], 1)
    return conv_t(x, g(f"{p}.conv.kernel"), g(f"{p}.conv.bias") if kw.get("use_bias", True) else None, kw.get("strides", 16))
  if cls == "TwoLayerSynthesis":                                 # :298-317
    s1, s2 = kw.get("strides", (8, 2))
    t = activation(conv_t(x, g(f"{p}.conv1.kernel"), g(f"{p}.conv1.bias"), s1), kw.get("activation_type", "igdn"), wts, f"{p}.activation")
    return conv_t(t, g(f"{p}.conv2.kernel"), g(f"{p}.conv2.bias"), s2)
  if cls == "TwoLayerResSynthesis":                              # :320-361 (res_type="conv")
    if kw.get("res_type", "conv") != "conv":
      raise NotImplementedError("res_type")
    s1, s2 = kw.get("strides", (8, 2))
    base = activation(conv_t(x, g(f"{p}.base_conv.kernel"), g(f"{p}.base_conv.bias"), s1), kw.get("activation_type", "igdn"), wts, f"{p}.activation")
    res = conv_t(x, g(f"{p}.res.kernel"), g(f"{p}.res.bias"), s1)
    return conv_t(base + res, g(f"{p}.out_conv.kernel"), g(f"{p}.out_conv.bias"), s2)
  if cls == "MBT2018Synthesis":                                  # :158-175
    n = kw.get("n_layers", 4)
    form = gdn_classic if kw.get("gdn_form", "gdn1") == "classic" else gdn1
    for i in range(n):
      x = conv_t(x, g(f"{p}.layer_{i}.kernel"), g(f"{p}.layer_{i}.bias"), 2, keras=False)
      if i + 1 < n:
        x = form(x, g(f"{p}.igdn_{i}.beta"), g(f"{p}.igdn_{i}.gamma"), True)
    return x
  if cls == "BLS2017Synthesis":                                  # :115-134
    for i, s in enumerate((2, 2, 4)):
      x = conv_t(x, g(f"{p}.layer_{i}.kernel"), g(f"{p}.layer_{i}.bias"), s, keras=False)
      if i < 2:
        x = gdn1(x, g(f"{p}.igdn_{i}.beta"), g(f"{p}.igdn_{i}.gamma"), True)
    return x
  if cls == "CNNSynthesis":                                      # :195-206 (one activation object shared by layers 0-2)
    act = kw.get("activation_type", "leaky_relu")
    for i in range(4):
      x = conv_t(x, g(f"{p}.layer_{i}.kernel"), g(f"{p}.layer_{i}.bias"), 2)
      if i < 3:
        x = activation(x, act, wts, f"{p}.activation")
    return x
  raise KeyError(cls)


def transform_vjp(cls, wts, x_nhwc, grad_out_nhwc, kwargs=None):
  """(out, grad_in) in NHWC float64: out = f(x); grad_in = J_f(x)^T grad_out (what tape.gradient returns for <grad_out, f(x)>)."""
  x = _nchw(x_nhwc).requires_grad_(True)
  out = forward(cls, wts, x, kwargs)
  (gin,) = torch.autograd.grad(out, x, _nchw(grad_out_nhwc))
  return _nhwc(out), _nhwc(gin)


def transform_jvp(cls, wts, x_nhwc, v_nhwc, kwargs=None):
  """J_f(x) v by forward-mode autodiff (for the <g, J v> == <J^T g, v> identity)."""
  x, v = _nchw(x_nhwc), _nchw(v_nhwc)
  _, jv = torch.autograd.functional.jvp(lambda t: forward(cls, wts, t, kwargs), x, v)
  return _nhwc(jv)
