"""CPU oracle for the shallow-ntc decode hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  ``shallow_ntc_b200`` never does.

PARITY UNPINNED at the TensorFlow / tensorflow-compression boundary, with two exceptions
listed first below: the reference (mandt-lab/shallow-ntc) has no tests for this path
and its arithmetic lives in third-party packages that are absent from
``/root/reference`` and cannot be installed offline (``requirements.txt:6-9``:
tensorflow==2.10.0, tensorflow_compression==2.10.0,
tensorflow_probability==0.18.0).  What IS pinned, and checked in
``tests/test_oracle.py``:

* OUTPUTS THE REFERENCE ITSELF RECORDED: (i) the PNG images embedded in the output cells of
  ``notebooks/vis_syn_filters.ipynb`` (real TF-2.10 runs of the trained jpegl model; extracted by
  ``tests/golden/make_golden_notebook.py`` into ``tests/golden/notebook_jpegl_responses.npz``): the response of one active latent
  pixel covers rows / columns 0..16 of the 2 x 2-latent image in all 100 recorded channels, on a constant background -- this
  pins the alignment of A1 for Keras ``Conv2DTranspose(18, 16, 'SAME')`` (p = 1; p = 0 or 2 contradict the recording) and that
  the 1 x 1 and 2 x 2 latent grids show the same kernel taps; (ii) LPIPS: the values ``lpips_tf2/test.py:17-19`` records for the
  image pairs it ships (``tests/test_lpips.py``);
* the reference's own structural known-answers (``results/all_params.csv``,
  ``results/flops_per_pixel.csv``, ``notebooks/get_flops.ipynb``): parameter
  counts, FLOPs/pixel and tensor shapes of every transform restated here;
* an independent implementation of each conv convention (torch-CPU
  ``conv_transpose2d`` + crop, and the gradient-of-SAME-conv definition TF uses),
  committed as fixtures under ``tests/golden/`` by ``tests/golden/make_golden.py``.

Third-party semantics restated (assumptions A1-A8, each a single switch):

A1  Keras ``Conv2DTranspose(padding="SAME")``: out[o] += in[n] * W[a] with
    o = n*s + a - p, p = max(k-s,0)//2, output length n_in*s, kernel variable
    layout [kh, kw, Cout, Cin]; then bias, then activation.
A2  tfc ``SignalConv2D(corr=False, strides_up=s, padding="same_zeros")``: same
    scatter form with p = (k-1)//2 (odd k), kernel layout [kh, kw, Cin, Cout].
A3  ``tf.round`` is round-half-to-even.
A4  ``tf.saturate_cast(x, uint8)`` clamps to [0, 255] then converts.
A5  ``GDN1`` (``common/transforms.py:8-63``): x * (beta + |x| @ gamma) when inverse,
    x / (...) otherwise; gamma indexed [in, out].  ``tfc.GDN(inverse=True)`` with the
    library defaults (``MBT2018Synthesis``, ``common/transforms.py:170``): in
    tensorflow-compression 2.x the defaults are alpha_parameter=1, epsilon_parameter=1
    (gdn.py: "GDN1", Johnston et al. 2019) -- the SAME function as GDN1; the reference's
    own ``GDN1`` class passes exactly those two values to ``tfc.GDN.__init__`` and calls
    itself "a copy of tfc.GDN that only implements the GDN1 activation".  The original
    form (alpha=2, epsilon=.5): x * sqrt(beta + x^2 @ gamma) is ``gdn_form='classic'``.
A6  ``LocationScaleIndexedEntropyModel._normalize_indexes`` clamps the float
    index to [0, num_scales-1]; the table row a range coder uses comes from
    ``ContinuousIndexedEntropyModel._flatten_indexes``: ``tf.cast(indexes, tf.int32)``,
    i.e. truncation (``index_rounding='trunc'``, the default here and in the product;
    'rint' = round-half-even stays selectable).  As recalled from tfc 2.10's
    ``entropy_models/continuous_indexed.py``; not verifiable offline.
A10 The rate term applies no likelihood lower bound (tfc 2.x entropy models dropped the
    ``likelihood_bound`` of tfc 1.x; ``mshyper/models.py:246-251`` passes none).
A7  ``ContinuousBatchedEntropyModel.quantize``: round(x - off) + off.
A8  ``tf.nn.leaky_relu`` alpha = 0.2; Keras "relu" = max(x, 0).
A9  ``tf.image.ssim`` / ``tf.image.ssim_multiscale`` (TF 2.10 ``image_ops_impl.py``) on uint8 inputs with
    ``max_val=255.``: images and max_val go through ``convert_image_dtype(., float32)`` (x / 255, max_val 1);
    11x11 Gaussian window (sigma 1.5, normalised by a softmax = outer product of the normalised 1-D window),
    VALID depthwise filtering, k1 = .01, k2 = .03; 5 scales with weights (.0448, .2856, .3001, .2363, .1333),
    2x2 VALID average pooling between scales after SYMMETRIC end-padding of odd sizes; relu on every factor,
    weighted geometric mean over scales, arithmetic mean over channels.

Two tiers share one code path, selected by ``dtype``:
T0  float64, per-tap scatter definition  -> the truth used for parity.
T1  float32, one BLAS GEMM [pixels x Cin] @ [Cin x k*k*Cout] + col2im slice adds
    -> the timed CPU stand-in for the reference's TF-2.10 CPU decode.
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------
# constants: mshyper/models.py:27-34
NUM_SCALES = 64
SCALE_MIN = 0.11
SCALE_MAX = 256.0


def scale_fn(i):
  """mshyper/models.py:31-32  SCALE_FN(i) = exp(log(.11) + i*(log 256 - log .11)/63)."""
  factor = (math.log(SCALE_MAX) - math.log(SCALE_MIN)) / (NUM_SCALES - 1.0)
  return np.exp(math.log(SCALE_MIN) + factor * np.asarray(i, dtype=np.float64))


# --------------------------------------------------------------------------
# transposed convolutions (A1, A2)

def keras_same_pad(k: int, s: int) -> int:
  """pad_before of the SAME correlation whose input-gradient Keras' Conv2DTranspose is."""
  return max(k - s, 0) // 2


def tfc_same_pad(k: int) -> int:
  return (k - 1) // 2


def _valid_range(n_in: int, s: int, a: int, p: int):
  """Input indices n with 0 <= n*s + a - p < n_in*s."""
  lo = 0 if a >= p else -((a - p) // s)  # ceil((p-a)/s)
  hi = min(n_in - 1, (n_in * s - 1 + p - a) // s)
  return lo, hi


def conv_transpose_scatter(x, w_oi, bias, s: int, p: int, dtype=np.float64, gemm_form=False):
  """out[b, ny*s+ay-p, nx*s+ax-p, co] += x[b, ny, nx, ci] * w_oi[ay, ax, co, ci]; + bias.

  ``w_oi`` is [kh, kw, Cout, Cin].  Output is [B, h*s, w*s, Cout].
  ``gemm_form`` (tier T1): a single [B*h*w, Cin] @ [Cin, k*k*Cout] GEMM followed by
  col2im slice adds, instead of one small matmul per tap (tier T0)."""
  x = np.asarray(x, dtype=dtype)
  w_oi = np.asarray(w_oi, dtype=dtype)
  B, h, w, cin = x.shape
  kh, kw, cout, cin2 = w_oi.shape
  assert cin == cin2, (x.shape, w_oi.shape)
  out = np.zeros((B, h * s, w * s, cout), dtype=dtype)
  if gemm_form:
    wm = np.ascontiguousarray(w_oi.transpose(3, 0, 1, 2).reshape(cin, kh * kw * cout))
    cols = (x.reshape(-1, cin) @ wm).reshape(B, h, w, kh, kw, cout)
  for ay in range(kh):
    ylo, yhi = _valid_range(h, s, ay, p)
    if yhi < ylo:
      continue
    for ax in range(kw):
      xlo, xhi = _valid_range(w, s, ax, p)
      if xhi < xlo:
        continue
      if gemm_form:
        contrib = cols[:, ylo:yhi + 1, xlo:xhi + 1, ay, ax, :]
      else:
        contrib = x[:, ylo:yhi + 1, xlo:xhi + 1, :] @ w_oi[ay, ax].T
      oy0 = ylo * s + ay - p
      ox0 = xlo * s + ax - p
      out[:, oy0:oy0 + (yhi - ylo) * s + 1:s, ox0:ox0 + (xhi - xlo) * s + 1:s, :] += contrib
  if bias is not None:
    out += np.asarray(bias, dtype=dtype)
  return out


def keras_conv2d_transpose(x, kernel, bias, strides: int, dtype=np.float64, gemm_form=False):
  """tf.keras.layers.Conv2DTranspose(padding="SAME") (A1); kernel [kh,kw,Cout,Cin].
  Call sites: common/transforms.py:85-90, 284, 307-313, 331-357, 371."""
  k = kernel.shape[0]
  return conv_transpose_scatter(x, kernel, bias, strides, keras_same_pad(k, strides), dtype, gemm_form)


def tfc_signal_conv2d_up(x, kernel, bias, strides_up: int, dtype=np.float64, gemm_form=False):
  """tfc.SignalConv2D(corr=False, strides_up=s, padding="same_zeros") (A2); kernel [kh,kw,Cin,Cout].
  Call sites: common/transforms.py:122-133, 169-172, 254-261."""
  k = kernel.shape[0]
  assert k % 2 == 1, "only odd kernel supports are used on the hot path"
  w_oi = np.asarray(kernel).transpose(0, 1, 3, 2)
  return conv_transpose_scatter(x, w_oi, bias, strides_up, tfc_same_pad(k), dtype, gemm_form)


# --------------------------------------------------------------------------
# activations (A5, A8)

def relu(x):
  return np.maximum(x, 0)


def leaky_relu(x, alpha=0.2):
  return np.where(x >= 0, x, x * np.asarray(alpha, dtype=x.dtype))


def gdn1(x, beta, gamma, inverse: bool):
  """common/transforms.py:27-63 (GDN1.call): norm = beta + |x| @ gamma; x*norm or x/norm."""
  dt = x.dtype
  norm = np.abs(x) @ np.asarray(gamma, dtype=dt) + np.asarray(beta, dtype=dt)
  return x * norm if inverse else x / norm


def gdn_classic(x, beta, gamma, inverse: bool):
  """tfc.GDN(alpha_parameter=2, epsilon_parameter=.5), the original GDN: norm = sqrt(beta + x^2 @ gamma)."""
  dt = x.dtype
  norm = np.sqrt(np.square(x) @ np.asarray(gamma, dtype=dt) + np.asarray(beta, dtype=dt))
  return x * norm if inverse else x / norm


def apply_activation(x, act, weights=None, prefix=None):
  """common/transforms.py:66-78 get_activation_op."""
  if act is None:
    return x
  a = act.lower()
  if a == "relu":
    return relu(x)
  if a in ("leaky_relu", "lrelu"):
    return leaky_relu(x)
  if a in ("igdn", "igdn1"):
    return gdn1(x, weights[prefix + ".beta"], weights[prefix + ".gamma"], inverse=True)
  if a in ("gdn", "gdn1"):
    return gdn1(x, weights[prefix + ".beta"], weights[prefix + ".gamma"], inverse=False)
  raise NotImplementedError(act)


# --------------------------------------------------------------------------
# decoder backward (f4: the gradient tf.GradientTape takes through the transforms in itinf_train_step,
# mshyper/models.py:401-408) -- the adjoints of the statements above, by definition

def conv_transpose_input_grad(g, w_oi, s: int, p: int, dtype=np.float64):
  """Adjoint of conv_transpose_scatter w.r.t. x:  gx[b, ny, nx, ci] = sum_{ay, ax, co} g[b, ny*s+ay-p, nx*s+ax-p, co] * w_oi[ay, ax, co, ci]
  (positions outside g contribute nothing; the bias has no input gradient)."""
  g = np.asarray(g, dtype=dtype)
  w_oi = np.asarray(w_oi, dtype=dtype)
  B, H, W, cout = g.shape
  kh, kw, cout2, cin = w_oi.shape
  assert cout == cout2 and H % s == 0 and W % s == 0
  h, w = H // s, W // s
  gx = np.zeros((B, h, w, cin), dtype=dtype)
  for ay in range(kh):
    ylo, yhi = _valid_range(h, s, ay, p)
    if yhi < ylo:
      continue
    for ax in range(kw):
      xlo, xhi = _valid_range(w, s, ax, p)
      if xhi < xlo:
        continue
      oy0, ox0 = ylo * s + ay - p, xlo * s + ax - p
      gx[:, ylo:yhi + 1, xlo:xhi + 1, :] += g[:, oy0:oy0 + (yhi - ylo) * s + 1:s, ox0:ox0 + (xhi - xlo) * s + 1:s, :] @ w_oi[ay, ax]
  return gx


def gdn1_vjp(x, g, beta, gamma, inverse: bool):
  """Adjoint of gdn1 at x:  t_j = x_j * n_j (inverse) or x_j / n_j,  n_j = beta_j + sum_i |x_i| gamma_ij
     inverse:  gx_k = g_k n_k + sign(x_k) sum_j g_j x_j gamma_kj
     forward:  gx_k = g_k / n_k - sign(x_k) sum_j g_j x_j / n_j^2 gamma_kj        (sign(0) = 0, the subgradient TF uses for abs)"""
  dt = x.dtype
  gamma = np.asarray(gamma, dtype=dt)
  n = np.abs(x) @ gamma + np.asarray(beta, dtype=dt)
  if inverse:
    return g * n + np.sign(x) * ((g * x) @ gamma.T)
  return g / n - np.sign(x) * ((g * x / (n * n)) @ gamma.T)


def activation_vjp(pre, g, act, weights=None, prefix=None):
  """Adjoint of apply_activation at the pre-activation value `pre` (tf.nn.relu / leaky_relu: derivative 0 / alpha at pre <= 0 ... < 0)."""
  if act is None:
    return g
  a = act.lower()
  if a == "relu":
    return g * (pre > 0)
  if a in ("leaky_relu", "lrelu"):
    return g * np.where(pre > 0, 1.0, 0.2).astype(g.dtype)
  if a in ("igdn", "igdn1"):
    return gdn1_vjp(pre, g, weights[prefix + ".beta"], weights[prefix + ".gamma"], True)
  if a in ("gdn", "gdn1"):
    return gdn1_vjp(pre, g, weights[prefix + ".beta"], weights[prefix + ".gamma"], False)
  raise NotImplementedError(act)


def hyper_synthesis_vjp(wts, z_hat, g, activation_type="relu", prefix="hyper_synthesis"):
  """J^T g of hyper_synthesis at z_hat (float64): returns (out, grad_z)."""
  dt = np.float64
  K = lambda i: np.asarray(wts[f"{prefix}.layer_{i}.kernel"], dt)
  pre0 = keras_conv2d_transpose(z_hat, K(0), wts[f"{prefix}.layer_0.bias"], 2, dt)
  a0 = apply_activation(pre0, activation_type)
  pre1 = keras_conv2d_transpose(a0, K(1), wts[f"{prefix}.layer_1.bias"], 2, dt)
  a1 = apply_activation(pre1, activation_type)
  out = keras_conv2d_transpose(a1, K(2), wts[f"{prefix}.layer_2.bias"], 1, dt)
  g = conv_transpose_input_grad(g, K(2), 1, keras_same_pad(3, 1))
  g = conv_transpose_input_grad(activation_vjp(pre1, g, activation_type), K(1), 2, keras_same_pad(5, 2))
  g = conv_transpose_input_grad(activation_vjp(pre0, g, activation_type), K(0), 2, keras_same_pad(5, 2))
  return out, g


def two_layer_res_synthesis_vjp(wts, y_hat, g, strides=(8, 2), activation_type="igdn", prefix="synthesis", res=True):
  """J^T g of two_layer_res_synthesis (res=True) / two_layer_synthesis (res=False) at y_hat (float64): (out, grad_y)."""
  dt = np.float64
  n1, n2 = ("base_conv", "out_conv") if res else ("conv1", "conv2")
  K1, K2 = np.asarray(wts[f"{prefix}.{n1}.kernel"], dt), np.asarray(wts[f"{prefix}.{n2}.kernel"], dt)
  pre = keras_conv2d_transpose(y_hat, K1, wts[f"{prefix}.{n1}.bias"], strides[0], dt)
  t = apply_activation(pre, activation_type, wts, f"{prefix}.activation")
  if res:
    Kr = np.asarray(wts[f"{prefix}.res.kernel"], dt)
    t = t + keras_conv2d_transpose(y_hat, Kr, wts[f"{prefix}.res.bias"], strides[0], dt)
  out = keras_conv2d_transpose(t, K2, wts[f"{prefix}.{n2}.bias"], strides[1], dt)
  gt = conv_transpose_input_grad(g, K2, strides[1], keras_same_pad(K2.shape[0], strides[1]))
  p1 = keras_same_pad(K1.shape[0], strides[0])
  gy = conv_transpose_input_grad(activation_vjp(pre, gt, activation_type, wts, f"{prefix}.activation"), K1, strides[0], p1)
  if res:
    gy = gy + conv_transpose_input_grad(gt, Kr, strides[0], p1)
  return out, gy


# --------------------------------------------------------------------------
# transforms (weights: dict name -> ndarray in the reference's native layouts)

def hyper_synthesis(wts, z_hat, activation_type="relu", dtype=np.float64, gemm_form=False, prefix="hyper_synthesis"):
  """common/transforms.py:222-232 HyperSynthesis: ConvT k5s2 -> ConvT k5s2 -> ConvT k3s1."""
  x = z_hat
  for i, s in enumerate((2, 2, 1)):
    x = keras_conv2d_transpose(x, wts[f"{prefix}.layer_{i}.kernel"], wts[f"{prefix}.layer_{i}.bias"], s, dtype, gemm_form)
    if i < 2:
      x = apply_activation(x, activation_type)
  return x


def jpeg_like_hyper_synthesis(wts, z_hat, dtype=np.float64, gemm_form=False, prefix="hyper_synthesis"):
  """common/transforms.py:364-377: one ConvT(k, 4) to 2*C channels."""
  return keras_conv2d_transpose(z_hat, wts[f"{prefix}.conv.kernel"], wts[f"{prefix}.conv.bias"], 4, dtype, gemm_form)


def hyper_synthesis_small(wts, z_hat, dtype=np.float64, gemm_form=False, prefix="hyper_synthesis"):
  """common/transforms.py:250-262: SignalConv 5x5 up2 + relu -> SignalConv 3x3 up1."""
  x = tfc_signal_conv2d_up(z_hat, wts[f"{prefix}.layer_0.kernel"], wts[f"{prefix}.layer_0.bias"], 2, dtype, gemm_form)
  x = relu(x)
  return tfc_signal_conv2d_up(x, wts[f"{prefix}.layer_1.kernel"], wts[f"{prefix}.layer_1.bias"], 1, dtype, gemm_form)


def jpeg_like_synthesis(wts, y_hat, strides=16, use_bias=True, use_offset=False, dtype=np.float64, gemm_form=False,
                        prefix="synthesis"):
  """common/transforms.py:265-295 JPEGLikeSynthesis."""
  x = np.asarray(y_hat, dtype=dtype)
  if use_offset:
    x = np.concatenate([x, np.ones(x.shape[:3] + (1,), dtype=dtype)], axis=-1)
  b = wts[f"{prefix}.conv.bias"] if use_bias else None
  return keras_conv2d_transpose(x, wts[f"{prefix}.conv.kernel"], b, strides, dtype, gemm_form)


def two_layer_synthesis(wts, y_hat, strides=(8, 2), activation_type="igdn", dtype=np.float64, gemm_form=False,
                        prefix="synthesis"):
  """common/transforms.py:298-317 TwoLayerSynthesis: conv2(act(conv1(z)))."""
  x = keras_conv2d_transpose(y_hat, wts[f"{prefix}.conv1.kernel"], wts[f"{prefix}.conv1.bias"], strides[0], dtype, gemm_form)
  x = apply_activation(x, activation_type, wts, f"{prefix}.activation")
  return keras_conv2d_transpose(x, wts[f"{prefix}.conv2.kernel"], wts[f"{prefix}.conv2.bias"], strides[1], dtype, gemm_form)


def depth_to_space2(x):
  """tf.nn.depth_to_space(x, 2) on NHWC (DCR order): out[b, 2y+dy, 2x+dx, c] = in[b, y, x, (2 dy + dx) * C + c]."""
  B, h, w, c4 = x.shape
  c = c4 // 4
  return x.reshape(B, h, w, 2, 2, c).transpose(0, 1, 3, 2, 4, 5).reshape(B, 2 * h, 2 * w, c)


def d2s_residual(wts, y_hat, dtype=np.float64, prefix="synthesis"):
  """common/transforms.py:339-348, the res_type="d2s" Sequential: depth_to_space(2), Conv2D 1x1 + leaky_relu (twice), then
  depth_to_space(2).  A11: the Keras activation string 'leaky_relu' = tf.nn.leaky_relu, alpha 0.2 (Keras 2.10 itself does not
  know the string; later versions define it with negative_slope 0.2)."""
  x = np.asarray(y_hat, dtype)
  for i in range(2):
    x = depth_to_space2(x)
    x = x @ np.asarray(wts[f"{prefix}.res.conv_{i}.kernel"], dtype)[0, 0] + np.asarray(wts[f"{prefix}.res.conv_{i}.bias"], dtype)
    x = np.where(x > 0, x, dtype(0.2) * x)
  return depth_to_space2(x)


def two_layer_res_synthesis(wts, y_hat, strides=(8, 2), activation_type="igdn", dtype=np.float64, gemm_form=False,
                            prefix="synthesis", res_type="conv"):
  """common/transforms.py:320-361 TwoLayerResSynthesis:
  out_conv(act(base_conv(z)) + res(z)); the activation sits inside base_conv (:331-334)."""
  base = keras_conv2d_transpose(y_hat, wts[f"{prefix}.base_conv.kernel"], wts[f"{prefix}.base_conv.bias"], strides[0], dtype, gemm_form)
  base = apply_activation(base, activation_type, wts, f"{prefix}.activation")
  if res_type == "conv":
    res = keras_conv2d_transpose(y_hat, wts[f"{prefix}.res.kernel"], wts[f"{prefix}.res.bias"], strides[0], dtype, gemm_form)
  elif res_type == "d2s":
    res = d2s_residual(wts, y_hat, dtype, prefix)
  else:
    raise NotImplementedError(res_type)
  return keras_conv2d_transpose(base + res, wts[f"{prefix}.out_conv.kernel"], wts[f"{prefix}.out_conv.bias"], strides[1], dtype, gemm_form)


def mbt2018_synthesis(wts, y_hat, n_layers=4, dtype=np.float64, gemm_form=False, prefix="synthesis", gdn_form="gdn1"):
  """common/transforms.py:158-175: n_layers x SignalConv2D 5x5 up2, tfc.GDN(inverse=True) after all but the last
  (A5: the tfc 2.x defaults make it IGDN1; gdn_form='classic' = the alpha=2 / epsilon=.5 form)."""
  x = y_hat
  act = gdn1 if gdn_form == "gdn1" else gdn_classic
  for i in range(n_layers):
    x = tfc_signal_conv2d_up(x, wts[f"{prefix}.layer_{i}.kernel"], wts[f"{prefix}.layer_{i}.bias"], 2, dtype, gemm_form)
    if i + 1 < n_layers:
      x = act(x, wts[f"{prefix}.igdn_{i}.beta"], wts[f"{prefix}.igdn_{i}.gamma"], inverse=True)
  return x


def bls2017_synthesis(wts, y_hat, dtype=np.float64, gemm_form=False, prefix="synthesis"):
  """common/transforms.py:115-134: 5x5 up2 + IGDN1, 5x5 up2 + IGDN1, 9x9 up4 (each layer its own GDN1)."""
  x = y_hat
  for i, s in enumerate((2, 2, 4)):
    x = tfc_signal_conv2d_up(x, wts[f"{prefix}.layer_{i}.kernel"], wts[f"{prefix}.layer_{i}.bias"], s, dtype, gemm_form)
    if i < 2:
      x = gdn1(x, wts[f"{prefix}.igdn_{i}.beta"], wts[f"{prefix}.igdn_{i}.gamma"], inverse=True)
  return x


def cnn_synthesis(wts, y_hat, activation_type="leaky_relu", dtype=np.float64, gemm_form=False, prefix="synthesis"):
  """common/transforms.py:195-206: 4 x Keras ConvT k5s2; ONE activation object shared by layers 0-2."""
  x = y_hat
  for i in range(4):
    x = keras_conv2d_transpose(x, wts[f"{prefix}.layer_{i}.kernel"], wts[f"{prefix}.layer_{i}.bias"], 2, dtype, gemm_form)
    if i < 3:
      x = apply_activation(x, activation_type, wts, f"{prefix}.activation")
  return x


_SYNTHESIS = {
  "JPEGLikeSynthesis": lambda w, y, kw, dt, g: jpeg_like_synthesis(
    w, y, kw.get("strides", 16), kw.get("use_bias", True), kw.get("use_offset", False), dt, g),
  "TwoLayerSynthesis": lambda w, y, kw, dt, g: two_layer_synthesis(
    w, y, kw.get("strides", (8, 2)), kw.get("activation_type", "igdn"), dt, g),
  "TwoLayerResSynthesis": lambda w, y, kw, dt, g: two_layer_res_synthesis(
    w, y, kw.get("strides", (8, 2)), kw.get("activation_type", "igdn"), dt, g, res_type=kw.get("res_type", "conv")),
  "MBT2018Synthesis": lambda w, y, kw, dt, g: mbt2018_synthesis(w, y, kw.get("n_layers", 4), dt, g, gdn_form=kw.get("gdn_form", "gdn1")),
  "BLS2017Synthesis": lambda w, y, kw, dt, g: bls2017_synthesis(w, y, dt, g),
  "CNNSynthesis": lambda w, y, kw, dt, g: cnn_synthesis(w, y, kw.get("activation_type", "leaky_relu"), dt, g),
}

_HYPER = {
  "HyperSynthesis": lambda w, z, kw, dt, g: hyper_synthesis(w, z, kw.get("activation_type", "relu"), dt, g),
  "JPEGLikeHyperSynthesis": lambda w, z, kw, dt, g: jpeg_like_hyper_synthesis(w, z, dt, g),
  "HyperSynthesisSmall": lambda w, z, kw, dt, g: hyper_synthesis_small(w, z, dt, g),
}


def synthesis(cls, wts, y_hat, kwargs=None, dtype=np.float64, gemm_form=False):
  return _SYNTHESIS[cls](wts, y_hat, kwargs or {}, dtype, gemm_form)


def hyper_synthesis_by_name(cls, wts, z_hat, kwargs=None, dtype=np.float64, gemm_form=False):
  return _HYPER[cls](wts, z_hat, kwargs or {}, dtype, gemm_form)


# --------------------------------------------------------------------------
# entropy-model glue and pixel epilogue

def scale_indexes(raw_sigma, index_rounding="trunc"):
  """mshyper/models.py:274-276 + tfc index handling (A6).

  Returns (i_c float64 clamped continuous index, idx uint8 table row, dist = distance
  of i_c to the nearest rounding boundary -- used for the tie margin of SURVEY F8)."""
  i_f = np.exp(np.asarray(raw_sigma, dtype=np.float64))
  i_c = np.minimum(np.maximum(i_f, 0.0), NUM_SCALES - 1.0)
  top = NUM_SCALES - 1.0
  if index_rounding == "rint":
    idx = np.rint(i_c)
    # boundaries k + .5, k = 0 .. S-2; above the clamp only the last one (S - 1.5) is near
    dist = np.where(i_f >= top, i_f - (top - 0.5), np.abs(np.abs(i_c - np.floor(i_c)) - 0.5))
  elif index_rounding == "trunc":
    idx = np.floor(i_c)
    # boundaries are the integers 1 .. S-1: below 1 only 1 is a boundary (exp > 0), above the clamp only S-1
    frac = i_c - np.floor(i_c)
    dist = np.where(i_f >= top, i_f - top, np.where(i_f < 1.0, 1.0 - i_f, np.minimum(frac, 1.0 - frac)))
  else:
    raise ValueError(index_rounding)
  return i_c, idx.astype(np.uint8), dist


def dequantize(q, mu):
  """mshyper/models.py:278-283; latent_rvs_lib.py:95-102: y_hat = round(y - mu) + mu = q + mu, one fp32 add."""
  return (np.asarray(q, dtype=np.float32) + np.asarray(mu, dtype=np.float32)).astype(np.float32)


def quantize_latent(y, offset=0.0):
  """A7 / A3: round_half_even(y - off) + off (the upstream definition of the integer symbols)."""
  return np.rint(np.asarray(y) - offset) + offset


def unpad_images(x, H, W):
  """common/image_utils.py:69-71."""
  return x[:, :H, :W, :]


def floats_to_pixels(x):
  """common/data_lib.py:28-29,48-52 + image_utils.py:22-23 (training=False), evaluated in float32
  like the reference: saturate_cast_u8(round_half_even((x + 0.5) * 255))."""
  x32 = np.asarray(x, dtype=np.float32)
  v = (x32 + np.float32(0.5)) * np.float32(255.0)
  return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def mse_psnr(a_u8, b_u8, max_val=255.0):
  """common/image_utils.py:26-38: per-image mean squared difference and PSNR."""
  a = np.asarray(a_u8, dtype=np.float64)
  b = np.asarray(b_u8, dtype=np.float64)
  mses = np.mean(np.square(a - b), axis=tuple(range(1, a.ndim)))
  with np.errstate(divide="ignore"):
    psnrs = -10.0 * (np.log(mses) - 2.0 * math.log(max_val)) / math.log(10.0)
  return mses, psnrs


MSSSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)   # tf.image: _MSSSIM_WEIGHTS


def _ssim_window(size=11, sigma=1.5):
  """tf.image _fspecial_gauss: softmax over the 2-D log-weights == outer product of this normalised 1-D window."""
  c = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
  g = np.exp(-0.5 * c * c / (sigma * sigma))
  return g / g.sum()


def _valid_filter(x, g):
  """Separable VALID filtering of [B,H,W,C] over H and W."""
  k = g.size
  H, W = x.shape[1], x.shape[2]
  t = sum(g[i] * x[:, i:H - k + 1 + i] for i in range(k))
  return sum(g[i] * t[:, :, i:W - k + 1 + i] for i in range(k))


def _ssim_per_channel(x, y, max_val=1.0, k1=0.01, k2=0.03):
  """tf.image _ssim_per_channel + _ssim_helper (compensation = 1): (mean of luminance*cs, mean of cs) per (image, channel)."""
  g = _ssim_window()
  c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
  m0, m1 = _valid_filter(x, g), _valid_filter(y, g)
  num0 = m0 * m1 * 2.0
  den0 = m0 * m0 + m1 * m1
  lum = (num0 + c1) / (den0 + c1)
  num1 = _valid_filter(x * y, g) * 2.0
  den1 = _valid_filter(x * x + y * y, g)
  cs = (num1 - num0 + c2) / (den1 - den0 + c2)
  return (lum * cs).mean(axis=(1, 2)), cs.mean(axis=(1, 2))


def _pool2(x):
  """ssim_multiscale's downscaling: SYMMETRIC pad of one row / column at the END of odd dimensions, then 2x2 VALID mean."""
  if x.shape[1] % 2:
    x = np.concatenate([x, x[:, -1:]], axis=1)
  if x.shape[2] % 2:
    x = np.concatenate([x, x[:, :, -1:]], axis=2)
  return 0.25 * (x[:, 0::2, 0::2] + x[:, 0::2, 1::2] + x[:, 1::2, 0::2] + x[:, 1::2, 1::2])


def msssim(a_u8, b_u8):
  """The validation-mode MS-SSIM of the reference (mshyper/models.py:321-332, factorized/models.py:145-156) for uint8
  images [B,H,W,C]: tf.image.ssim when both sides are < 160 px, tf.image.ssim_multiscale(max_val=255.) otherwise.
  Returns (msssim [B], msssim_db [B])."""
  x = np.asarray(a_u8, dtype=np.float64) / 255.0
  y = np.asarray(b_u8, dtype=np.float64) / 255.0
  H, W = x.shape[1], x.shape[2]
  if H < 160 and W < 160:
    val = _ssim_per_channel(x, y)[0].mean(axis=-1)
  else:
    mcs = []
    for k in range(len(MSSSIM_WEIGHTS)):
      if k > 0:
        x, y = _pool2(x), _pool2(y)
      if x.shape[1] < 11 or x.shape[2] < 11:
        raise ValueError("ssim_multiscale: image too small for 5 scales of an 11x11 window")
      ssim_pc, cs = _ssim_per_channel(x, y)
      mcs.append(np.maximum(cs, 0.0))
    mcs.pop()
    fac = np.stack(mcs + [np.maximum(ssim_pc, 0.0)], axis=-1)            # [B, C, 5]
    val = np.prod(fac ** np.asarray(MSSSIM_WEIGHTS), axis=-1).mean(axis=-1)
  with np.errstate(divide="ignore"):
    db = -10.0 * np.log(1.0 - val) / math.log(10.0)
  return val, db


def noisy_normal_bits(q, i_c):
  """a7: bits of the integer symbol q = round(y - mu) under tfc.NoisyNormal(loc=0, scale=SCALE_FN(i_c)) -- what
  LocationScaleIndexedEntropyModel.__call__(y, indexes, loc=mu, training=False) sums into latent_bits
  (mshyper/models.py:246-248, 278-279).  i_c is the CONTINUOUS clamped index (compression=False never rounds it).
  P = Phi((q+.5)/s) - Phi((q-.5)/s), evaluated in the log domain on the upper tail like tfc's UniformNoiseAdapter
  (log survival functions of the two bin edges), so far tails stay finite."""
  from scipy.special import log_ndtr
  sig = scale_fn(np.asarray(i_c, dtype=np.float64))
  aq = np.abs(np.asarray(q, dtype=np.float64))
  la = log_ndtr(-(aq - 0.5) / sig)
  lb = log_ndtr(-(aq + 0.5) / sig)
  with np.errstate(divide="ignore"):
    lp = la + np.log(-np.expm1(lb - la))
  return -lp / math.log(2.0)


gaussian_bits = noisy_normal_bits   # earlier name


def _softplus(x):
  return np.logaddexp(0.0, x)


def _log_sigmoid(x):
  return -np.logaddexp(0.0, -x)


def deep_factorized_logits(x, wts, prefix="prior"):
  """tfc.DeepFactorized._logits_cumulative with num_filters=(3,3,3) (the default the reference uses,
  mshyper/models.py:135).  x [..., C]; variables are the RAW tfc variables: matrix_i [C,f_out,f_in] (softplus applied
  here), bias_i [C,f_out,1], factor_i [C,f_out,1] (tanh applied here)."""
  x = np.asarray(x, dtype=np.float64)
  h = x[..., None]                                   # [..., C, f=1]
  for i in range(4):
    m = _softplus(np.asarray(wts[f"{prefix}.matrix_{i}"], dtype=np.float64))      # [C, fo, fi]
    h = np.einsum("coi,...ci->...co", m, h) + np.asarray(wts[f"{prefix}.bias_{i}"], dtype=np.float64)[:, :, 0]
    if i < 3:
      h = h + np.tanh(np.asarray(wts[f"{prefix}.factor_{i}"], dtype=np.float64)[:, :, 0]) * np.tanh(h)
  return h[..., 0]


def deep_factorized_bits(z_hat, wts, prefix="prior"):
  """hyper_latent_bits element-wise: -log2( c(z+.5) - c(z-.5) ), c = sigmoid(logits_cumulative), under
  tfc.NoisyDeepFactorized (ContinuousBatchedEntropyModel.__call__(z, training=False), mshyper/models.py:249-252),
  with the side of the median chosen so that no two numbers close to 1 are subtracted."""
  z = np.asarray(z_hat, dtype=np.float64)
  lower = deep_factorized_logits(z - 0.5, wts, prefix)
  upper = deep_factorized_logits(z + 0.5, wts, prefix)
  sgn = np.where(lower + upper > 0, -1.0, 1.0)
  u, l = sgn * upper, sgn * lower
  big, small = np.maximum(u, l), np.minimum(u, l)
  lb, ls = _log_sigmoid(big), _log_sigmoid(small)
  with np.errstate(divide="ignore"):
    lp = lb + np.log(-np.expm1(ls - lb))
  return -lp / math.log(2.0)


def rate_bits(wts, raw_sigma, q_y, z_hat=None, prefix="prior"):
  """(bits_y [B], bits_z [B] or None): the rate half of frame_loss_given_latent_rvs(training=False), :278-279, 300-310."""
  i_c, _, _ = scale_indexes(raw_sigma)
  by = noisy_normal_bits(q_y, i_c).reshape(q_y.shape[0], -1).sum(1)
  bz = None
  if z_hat is not None and f"{prefix}.matrix_0" in wts:
    bz = deep_factorized_bits(z_hat, wts, prefix).reshape(z_hat.shape[0], -1).sum(1)
  return by, bz


def mshyper_decode(wts, synthesis_cls, z_hat, q_y, H, W, synthesis_kwargs=None,
                   hyper_cls="HyperSynthesis", hyper_kwargs=None, original_u8=None,
                   dtype=np.float64, gemm_form=False, index_rounding="trunc"):
  """mshyper/models.py:269-317 (training=False), from decoded symbols.

  z_hat: [B, Hp/64, Wp/64, Cz] integer-valued; q_y = round(y - mu): [B, Hp/16, Wp/16, Cy]."""
  hs = hyper_synthesis_by_name(hyper_cls, wts, z_hat, hyper_kwargs, dtype, gemm_form)
  cy = hs.shape[-1] // 2
  mu, raw_sigma = hs[..., :cy], hs[..., cy:]          # tf.split(..., 2, axis=-1)  :274
  i_c, idx, dist = scale_indexes(raw_sigma, index_rounding)   # :275-276
  y_hat = dequantize(q_y, mu)                           # :278
  recon = synthesis(synthesis_cls, wts, y_hat.astype(dtype), synthesis_kwargs, dtype, gemm_form)   # :297
  recon = unpad_images(recon, H, W)                     # :298
  out = dict(mu=mu, raw_sigma=raw_sigma, i_c=i_c, idx=idx, idx_dist=dist, y_hat=y_hat,
             recon=recon, recon_u8=floats_to_pixels(recon))                                       # :314
  if original_u8 is not None:
    out["mse"], out["psnr"] = mse_psnr(original_u8, out["recon_u8"])                              # :315
  out["bits_y"], out["bits_z"] = rate_bits(wts, raw_sigma, q_y, z_hat)                            # :278-279, 300-310
  return out


def factorized_decode(wts, synthesis_cls, q_y, H, W, synthesis_kwargs=None, original_u8=None,
                      dtype=np.float64, gemm_form=False):
  """factorized/models.py:101-141 (training=False): y_hat = round(y) (offset 0) -> synthesis -> crop -> u8."""
  y_hat = np.asarray(q_y, dtype=np.float32)
  recon = synthesis(synthesis_cls, wts, y_hat.astype(dtype), synthesis_kwargs, dtype, gemm_form)
  recon = unpad_images(recon, H, W)
  out = dict(y_hat=y_hat, recon=recon, recon_u8=floats_to_pixels(recon))
  if original_u8 is not None:
    out["mse"], out["psnr"] = mse_psnr(original_u8, out["recon_u8"])
  return out


# --------------------------------------------------------------------------
# structural known-answers (parameter and MAC counts) -- checked against results/*.csv in tests

def count_params(wts, prefix):
  return int(sum(int(np.prod(v.shape)) for k, v in wts.items() if k.startswith(prefix + ".")))


def convt_macs(h, w, k, s, cin, cout):
  """MACs of a transposed conv as TF's profiler counts them: every (input pixel, tap) pair."""
  return h * w * k * k * cin * cout


# --------------------------------------------------------------------------
# LPIPS (evaluate-loop metric): lpips_tf2/lpips_tensorflow.py:14-72 as called by mshyper/models.py:334-340

VGG_BLOCKS = ((64, 64), (128, 128), (256, 256, 256), (512, 512, 512), (512, 512, 512))   # Keras VGG16 conv stacks, block1 .. block5
LPIPS_SCALE = (0.458, 0.448, 0.450)     # image_preprocess, lpips_tensorflow.py:18-19
LPIPS_SHIFT = (-0.030, -0.088, -0.188)


def lpips_variable_shapes():
  """name -> shape: 13 VGG16 convs ``lpips.conv_i.kernel`` [3,3,Cin,Cout] / ``.bias`` (Keras Conv2D layout) and the five
  1x1 ``lpips.lin_l.kernel`` [C_l] (Conv2D(1, 1, use_bias=False) over channels)."""
  v, cin, i = {}, 3, 0
  for block in VGG_BLOCKS:
    for cout in block:
      v[f"lpips.conv_{i}.kernel"] = (3, 3, cin, cout)
      v[f"lpips.conv_{i}.bias"] = (cout,)
      cin, i = cout, i + 1
  for l, block in enumerate(VGG_BLOCKS):
    v[f"lpips.lin_{l}.kernel"] = (block[-1],)
  return v


def conv2d_same(x, kernel_io, bias, dtype=np.float64):
  """Keras Conv2D(3x3, padding='same', strides 1): out[o] = sum_a x[o + a - 1] * K[a]  (cross-correlation), kernel [kh,kw,Cin,Cout]."""
  x = np.asarray(x, dtype=dtype)
  k = np.asarray(kernel_io, dtype=dtype)
  B, h, w, cin = x.shape
  kh, kw = k.shape[:2]
  ph, pw = (kh - 1) // 2, (kw - 1) // 2
  xp = np.zeros((B, h + kh - 1, w + kw - 1, cin), dtype=dtype)
  xp[:, ph:ph + h, pw:pw + w] = x
  out = np.zeros((B, h, w, k.shape[3]), dtype=dtype)
  for ay in range(kh):
    for ax in range(kw):
      out += xp[:, ay:ay + h, ax:ax + w] @ k[ay, ax]
  return out + np.asarray(bias, dtype=dtype)


def maxpool2(x):
  """Keras MaxPooling2D(2, 2, 'valid'): odd trailing rows / columns are dropped."""
  B, h, w, c = x.shape
  x = x[:, :h // 2 * 2, :w // 2 * 2]
  return x.reshape(B, h // 2, 2, w // 2, 2, c).max(axis=(2, 4))


def vgg16_features(wts, x, dtype=np.float64, prefix="lpips"):
  """perceptual_model (lpips_tensorflow.py:124-136): outputs of block1_conv2, block2_conv2, block3_conv3, block4_conv3, block5_conv3."""
  feats, i = [], 0
  for b, block in enumerate(VGG_BLOCKS):
    if b > 0:
      x = maxpool2(x)
    for _ in block:
      x = relu(conv2d_same(x, wts[f"{prefix}.conv_{i}.kernel"], wts[f"{prefix}.conv_{i}.bias"], dtype))
      i += 1
    feats.append(x)
  return feats


def lpips(wts, image_a, image_b, dtype=np.float64, prefix="lpips", return_layers=False):
  """learned_perceptual_metric_model([a, b]) for images [B,H,W,3] in [0, 255] (the reference passes image_batch and
  reconstruction as they are, mshyper/models.py:339): preprocess, VGG16 features, unit-normalise over channels (x * rsqrt(sum x^2),
  NO epsilon: an all-zero feature vector gives NaN, as in the reference), squared difference, 1x1 lin, spatial mean, sum over layers."""
  def pre(im):
    x = np.asarray(im, dtype=dtype) / 127.5 - 1.0
    return (x - np.asarray(LPIPS_SHIFT, dtype=dtype)) / np.asarray(LPIPS_SCALE, dtype=dtype)
  fa, fb = vgg16_features(wts, pre(image_a), dtype, prefix), vgg16_features(wts, pre(image_b), dtype, prefix)
  per_layer = []
  with np.errstate(divide="ignore", invalid="ignore"):
    for l, (a, b) in enumerate(zip(fa, fb)):
      na = a * (1.0 / np.sqrt(np.sum(a * a, axis=-1, keepdims=True)))
      nb = b * (1.0 / np.sqrt(np.sum(b * b, axis=-1, keepdims=True)))
      d = (na - nb) ** 2 @ np.asarray(wts[f"{prefix}.lin_{l}.kernel"], dtype=dtype)
      per_layer.append(d.mean(axis=(1, 2)))
  total = np.sum(per_layer, axis=0)
  return (total, np.stack(per_layer, axis=-1)) if return_layers else total


def lpips_weights_from_reference_checkpoints(vgg_prefix, lin_prefix, prefix="lpips"):
  """The vendored checkpoints of the reference (lpips_tf2/models/{vgg,lin}/exported, read with oracle/tf_checkpoint.py) under
  the names above.  Keras tracks ``layer_with_weights-i`` in layer order: VGG16's 13 convs, the linear model's five 1x1 convs."""
  from . import tf_checkpoint
  vgg, lin = tf_checkpoint.load_checkpoint(vgg_prefix), tf_checkpoint.load_checkpoint(lin_prefix)
  w = {}
  for i in range(13):
    w[f"{prefix}.conv_{i}.kernel"] = vgg[f"layer_with_weights-{i}/kernel/.ATTRIBUTES/VARIABLE_VALUE"]
    w[f"{prefix}.conv_{i}.bias"] = vgg[f"layer_with_weights-{i}/bias/.ATTRIBUTES/VARIABLE_VALUE"]
  for l in range(5):
    w[f"{prefix}.lin_{l}.kernel"] = lin[f"layer_with_weights-{l}/kernel/.ATTRIBUTES/VARIABLE_VALUE"].reshape(-1)
  return w
