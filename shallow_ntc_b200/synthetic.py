"""Seeded synthetic inputs and random-init weights (checkpoints and datasets are unavailable offline).

Per-image seed = base_seed + global image index, so sharding images over GPUs never changes the data
(SURVEY 8(d)).  Two weight sets: "init" (framework default initialisers) and "stress" (non-zero biases,
dense GDN gammas, and a hyper-synthesis head scaled so that exp(raw_sigma) spans both clamps and all 64
scale-table rows -- with "init" weights raw_sigma ~ 0 and idx == 1 everywhere, which tests nothing)."""
from __future__ import annotations

import numpy as np

BASE_SEED = 20231003
WEIGHT_SEED = 1234


def make_latents(z_shape, y_shape, first_index=0, base_seed=BASE_SEED):
  """z_hat = clip(rint(N(0, 1.5^2)), -16, 16); q_y = clip(rint(N(0, s_c^2)), -127, 127) with per-channel
  s_c = exp(U(ln 0.05, ln 8)): sparse, trained-model-like (many all-zero channels, a few reaching +-30).
  Returns (z_hat float32 | None, q_y float32); q_y is integer-valued and int8-representable."""
  B = y_shape[0]
  q = np.empty(y_shape, dtype=np.float32)
  z = np.empty(z_shape, dtype=np.float32) if z_shape is not None else None
  for b in range(B):
    rng = np.random.default_rng(base_seed + first_index + b)
    if z is not None:
      z[b] = np.clip(np.rint(rng.normal(0.0, 1.5, size=z_shape[1:])), -16, 16)
    s_c = np.exp(rng.uniform(np.log(0.05), np.log(8.0), size=y_shape[-1]))
    q[b] = np.clip(np.rint(rng.normal(0.0, 1.0, size=y_shape[1:]) * s_c), -127, 127)
  return z, q


def _glorot(rng, shape, fan):
  lim = np.sqrt(6.0 / fan)
  return rng.uniform(-lim, lim, size=shape).astype(np.float32)


# gain on the last synthesis layer so that the synthetic reconstruction mostly stays inside the pixel
# range instead of saturating (random weights are not image-like); keyed by (synthesis class, kind)
OUT_GAIN = {
  ("TwoLayerResSynthesis", "init"): 0.6, ("TwoLayerResSynthesis", "stress"): 0.5,
  ("TwoLayerSynthesis", "init"): 0.7, ("TwoLayerSynthesis", "stress"): 0.55,
  # MBT2018Synthesis: three IGDN1 stages (tfc.GDN defaults, x * (beta + |x| @ gamma)) grow the signal ~25x more than the
  # sqrt form the first round assumed (0.07 is the gain that suits gdn_form="classic")
  ("MBT2018Synthesis", "stress"): 0.003, ("BLS2017Synthesis", "stress"): 0.04, ("CNNSynthesis", "stress"): 0.3,
}


def make_weights(variable_shapes: dict, kind="init", seed=WEIGHT_SEED, synthesis_cls=None, out_gain=None):
  """variable_shapes: name -> shape (Model.variable_shapes()).  Kernels Glorot-uniform over
  k*k*(Cin+Cout); biases 0; GDN beta = 1, gamma = 0.1*I (tfc defaults).  ``synthesis_cls`` selects
  the OUT_GAIN applied to the last synthesis layer (``out_gain`` overrides it)."""
  assert kind in ("init", "stress")
  rng = np.random.default_rng(seed)
  rng_prior = np.random.default_rng(seed + 1)   # own stream: the other weights do not depend on whether a prior is present
  w = {}
  for name in sorted(variable_shapes):
    shape = tuple(variable_shapes[name])
    if name.endswith(".kernel"):
      k = shape[0]
      w[name] = _glorot(rng, shape, k * k * (shape[2] + shape[3]))
    elif name.endswith(".bias"):
      w[name] = (rng.uniform(-0.05, 0.05, size=shape) if kind == "stress" else np.zeros(shape)).astype(np.float32)
    elif name.endswith(".beta"):
      w[name] = (1.0 + (rng.uniform(0, 0.5, size=shape) if kind == "stress" else 0.0) * np.ones(shape)).astype(np.float32)
    elif name.startswith("prior."):
      w[name] = _deep_factorized_init(name, shape, rng_prior, kind)
    elif name.endswith(".gamma"):
      g = 0.1 * np.eye(shape[0])
      if kind == "stress":
        g = g + rng.uniform(0, 0.02, size=shape)
      w[name] = g.astype(np.float32)
    else:
      raise KeyError(name)
  if kind == "stress":
    # last hyper-synthesis layer: sigma half (second half of the output channels) spans (0.2, 120) after exp
    heads = [n for n in w if n.startswith("hyper_synthesis.") and n.endswith(".kernel")]
    if heads:
      _scale_sigma_head(w, sorted(heads)[-1], rng)
  gain = out_gain if out_gain is not None else OUT_GAIN.get((synthesis_cls, kind))
  if gain is not None:
    last = {"TwoLayerResSynthesis": "out_conv", "TwoLayerSynthesis": "conv2"}.get(synthesis_cls)
    if last is None:
      last = sorted(n for n in w if n.startswith("synthesis.layer_") and n.endswith(".kernel"))[-1][len("synthesis."):-len(".kernel")]
    w[f"synthesis.{last}.kernel"] *= np.float32(gain)
    w[f"synthesis.{last}.bias"] *= np.float32(gain)
  return w


def _deep_factorized_init(name, shape, rng, kind, init_scale=10.0, num_filters=(3, 3, 3)):
  """tfc.DeepFactorized variable initialisers: matrix_i = log(expm1(1 / scale / f_out)) with
  scale = init_scale ** (1 / (len(num_filters) + 1)), bias_i ~ U(-.5, .5), factor_i = 0.  "stress" perturbs matrices
  and gives non-zero factors so that the tanh gates matter."""
  f_out = shape[1]
  if ".matrix_" in name:
    scale = init_scale ** (1.0 / (len(num_filters) + 1))
    m = np.full(shape, np.log(np.expm1(1.0 / scale / f_out)))
    if kind == "stress":
      m = m + rng.uniform(-0.3, 0.3, size=shape)
    return m.astype(np.float32)
  if ".bias_" in name:
    return rng.uniform(-0.5, 0.5, size=shape).astype(np.float32)
  return (rng.uniform(-0.8, 0.8, size=shape) if kind == "stress" else np.zeros(shape)).astype(np.float32)


def _scale_sigma_head(w, kname, rng):
  """Multiply the sigma-half of the head kernel by 4 and give it a bias in U(-1, 4.5)."""
  bname = kname[:-len(".kernel")] + ".bias"
  cout = w[bname].shape[0]
  half = cout // 2
  kern = w[kname]
  if kern.shape[2] == cout:       # Keras Conv2DTranspose [kh,kw,Cout,Cin]
    kern[:, :, half:, :] *= 4.0
  else:                           # tfc SignalConv2D [kh,kw,Cin,Cout]
    kern[:, :, :, half:] *= 4.0
  w[bname][half:] = rng.uniform(-1.0, 4.5, size=cout - half).astype(np.float32)


def make_original(recon_u8, first_index=0, base_seed=BASE_SEED, noise_std=6.0):
  """A fixed synthetic "original" = reconstruction + seeded noise, so PSNR is finite (30-40 dB)."""
  out = np.empty_like(recon_u8)
  for b in range(recon_u8.shape[0]):
    rng = np.random.default_rng(base_seed + 7919 + first_index + b)
    out[b] = np.clip(np.rint(recon_u8[b].astype(np.float64) + rng.normal(0, noise_std, size=recon_u8.shape[1:])), 0, 255).astype(np.uint8)
  return out


def shard_range(n_items: int, rank: int, world: int):
  """Contiguous block partition of the image list: item i -> rank floor(i * world / n_items)."""
  lo = (rank * n_items + world - 1) // world
  hi = ((rank + 1) * n_items + world - 1) // world
  return lo, hi
