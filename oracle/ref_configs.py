"""Benchmark configurations, variable shapes and seeded inputs for the CPU arms.  TEST INFRASTRUCTURE ONLY.

``bench.py --impl reference`` and the ``cpu_baseline`` leg must not load the product (``shallow_ntc_b200`` maps
``libsntc.so`` on import), so everything the CPU restatements need is derived here from the reference's own config
files: the ``transform_config`` dicts (``mshyper/configs/*.py``, ``factorized/configs/bls2017.py``) and the variable
layouts of the layer classes they name (``common/transforms.py``).  ``tests/test_oracle.py`` checks that these shapes are
exactly the ones the product asks for, so the two arms cannot drift apart.

The seeded generators themselves live in ``shallow_ntc_b200/synthetic.py`` (pure numpy); that single FILE is loaded by
path, without importing the package.
"""
from __future__ import annotations

import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = {
  "jpegl": dict(bottleneck=320, hyperprior=True,  # mshyper/configs/jpegl.py:36-39
                synthesis=dict(cls="JPEGLikeSynthesis", kernel_size=18, strides=16)),
  "two_layer_syn": dict(bottleneck=320, hyperprior=True,  # mshyper/configs/two_layer_syn.py:36-40
                        synthesis=dict(cls="TwoLayerResSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5),
                                       activation_type="igdn", res_type="conv")),
  "two_layer_syn2": dict(bottleneck=320, hyperprior=True,  # mshyper/configs/two_layer_syn2.py:47-50
                         synthesis=dict(cls="TwoLayerSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5), activation_type="igdn")),
  "mbt2018": dict(bottleneck=320, hyperprior=True,  # mshyper/configs/mbt2018.py:34-39
                  synthesis=dict(cls="MBT2018Synthesis", channels_base=192, output_channels=3)),
  "bls2017": dict(bottleneck=256, hyperprior=False,  # factorized/configs/bls2017.py:35-38
                  synthesis=dict(cls="BLS2017Synthesis", num_filters=256)),
}


def get_config(name: str) -> dict:
  base = name.split(":")[0]
  cfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in CONFIGS[base].items()}
  if ":" in name:   # two_layer_syn2:24 -> hidden width (mshyper/configs/two_layer_syn2.py:87-89)
    cfg["synthesis"]["channels"] = (int(name.split(":")[1]), 3)
  return cfg


def _keras(prefix, name, k, cin, cout, bias=True):
  v = {f"{prefix}.{name}.kernel": (k, k, cout, cin)}     # Conv2DTranspose kernel [kh, kw, Cout, Cin]
  if bias:
    v[f"{prefix}.{name}.bias"] = (cout,)
  return v


def _tfc(prefix, name, k, cin, cout):
  return {f"{prefix}.{name}.kernel": (k, k, cin, cout), f"{prefix}.{name}.bias": (cout,)}   # SignalConv2D kernel [kh, kw, Cin, Cout]


def _gdn(prefix, name, c):
  return {f"{prefix}.{name}.beta": (c,), f"{prefix}.{name}.gamma": (c, c)}


def variable_shapes(name: str, prior: bool = False) -> dict:
  """name -> shape of every decode-side variable of the config, in the reference's native layouts."""
  cfg = get_config(name)
  C, syn = cfg["bottleneck"], cfg["synthesis"]
  v = {}
  if cfg["hyperprior"]:   # HyperSynthesis(bottleneck_size)   common/transforms.py:222-232, mshyper/models.py:126-129
    v.update(_keras("hyper_synthesis", "layer_0", 5, C, C))
    v.update(_keras("hyper_synthesis", "layer_1", 5, C, int(C * 1.5)))
    v.update(_keras("hyper_synthesis", "layer_2", 3, int(C * 1.5), C * 2))
  cls = syn["cls"]
  if cls == "JPEGLikeSynthesis":            # :265-295
    v.update(_keras("synthesis", "conv", syn["kernel_size"], C + int(syn.get("use_offset", False)), 3, syn.get("use_bias", True)))
  elif cls == "TwoLayerSynthesis":          # :298-317
    c1, co = syn["channels"]
    v.update(_keras("synthesis", "conv1", syn["kernel_sizes"][0], C, c1))
    v.update(_gdn("synthesis", "activation", c1))
    v.update(_keras("synthesis", "conv2", syn["kernel_sizes"][1], c1, co))
  elif cls == "TwoLayerResSynthesis":       # :320-361
    c1, co = syn["channels"]
    v.update(_keras("synthesis", "base_conv", syn["kernel_sizes"][0], C, c1))
    v.update(_keras("synthesis", "res", syn["kernel_sizes"][0], C, c1))
    v.update(_gdn("synthesis", "activation", c1))
    v.update(_keras("synthesis", "out_conv", syn["kernel_sizes"][1], c1, co))
  elif cls == "MBT2018Synthesis":           # :158-175
    cb, cin = syn["channels_base"], C
    for i in range(4):
      v.update(_tfc("synthesis", f"layer_{i}", 5, cin, 3 if i == 3 else cb))
      if i < 3:
        v.update(_gdn("synthesis", f"igdn_{i}", cb))
      cin = cb
  elif cls == "BLS2017Synthesis":           # :115-134
    nf = syn["num_filters"]
    v.update(_tfc("synthesis", "layer_0", 5, C, nf)); v.update(_gdn("synthesis", "igdn_0", nf))
    v.update(_tfc("synthesis", "layer_1", 5, nf, nf)); v.update(_gdn("synthesis", "igdn_1", nf))
    v.update(_tfc("synthesis", "layer_2", 9, nf, 3))
  else:
    raise KeyError(cls)
  if prior and cfg["hyperprior"]:   # tfc.NoisyDeepFactorized(batch_shape=(Cz,)), num_filters (3,3,3)   mshyper/models.py:135
    f = (1, 3, 3, 3, 1)
    for i in range(4):
      v[f"prior.matrix_{i}"] = (C, f[i + 1], f[i])
      v[f"prior.bias_{i}"] = (C, f[i + 1], 1)
      if i < 3:
        v[f"prior.factor_{i}"] = (C, f[i + 1], 1)
  return v


def latent_shapes(name: str, batch: int, H: int, W: int):
  """(z shape | None, y shape): pad_images to a multiple of 64 (mshyper/models.py:137-140, 218) or 16 (factorized/models.py:30, 76)."""
  cfg = get_config(name)
  d = 64 if cfg["hyperprior"] else 16
  Hp, Wp = -(-H // d) * d, -(-W // d) * d
  C = cfg["bottleneck"]
  y = (batch, Hp // 16, Wp // 16, C)
  z = (batch, Hp // 64, Wp // 64, C) if cfg["hyperprior"] else None
  return z, y


def load_synthetic():
  """shallow_ntc_b200/synthetic.py as a stand-alone module (numpy only): the package __init__ -- and with it libsntc.so -- is
  NOT imported."""
  spec = importlib.util.spec_from_file_location("_sntc_synthetic_standalone", os.path.join(ROOT, "shallow_ntc_b200", "synthetic.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


def make_case(name: str, batch: int, H: int, W: int, kind: str = "stress", first_index: int = 0):
  """(cfg, weights, z_hat | None, q_y) -- the same seeded weights and symbols the GPU arm decodes."""
  syn = load_synthetic()
  cfg = get_config(name)
  wts = syn.make_weights(variable_shapes(name), kind, synthesis_cls=cfg["synthesis"]["cls"])
  zs, ys = latent_shapes(name, batch, H, W)
  z, q = syn.make_latents(zs, ys, first_index=first_index)
  return cfg, wts, z, q
