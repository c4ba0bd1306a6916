"""not-gpu: host-side mirror of the reference's plugin interface (registry, constructor kwargs, config
dicts, geometry) and the DLPack / array-interface tensor hand-off."""
import ctypes
import numpy as np
import pytest

from shallow_ntc_b200 import class_builder, build_config, CONFIGS, Model, FactorizedModel, as_tensor, _lib
from shallow_ntc_b200 import transforms as T
from shallow_ntc_b200 import synthetic


def test_registry_has_the_decoder_classes_of_the_reference():
  """common/transforms.py:383-393 minus the encoder-side classes."""
  for name in ("BLS2017Synthesis", "CNNSynthesis", "HyperSynthesis", "MBT2018Synthesis", "HyperSynthesisSmall",
               "JPEGLikeSynthesis", "TwoLayerSynthesis", "TwoLayerResSynthesis", "JPEGLikeHyperSynthesis"):
    assert name in class_builder
  t = class_builder.build("TwoLayerResSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5),
                          activation_type="igdn", res_type="conv")
  assert t.upsample == 16 and t.out_channels == 3
  with pytest.raises(KeyError):
    class_builder.build("ElicAnalysis")
  d2s = class_builder.build("TwoLayerResSynthesis", res_type="d2s")       # common/transforms.py:339-348
  v = d2s.variable_shapes(320)
  assert v["synthesis.res.conv_0.kernel"] == (1, 1, 80, 192) and v["synthesis.res.conv_1.kernel"] == (1, 1, 48, 48)
  assert "synthesis.res.kernel" not in v and d2s.desc(320).kind == 16
  with pytest.raises(NotImplementedError):
    class_builder.build("TwoLayerResSynthesis", res_type="bilinear")     # the reference raises NotImplementedError too (:349-350)
  with pytest.raises(ValueError):
    class_builder.build("TwoLayerResSynthesis", strides=(4, 4), res_type="d2s")
  with pytest.raises(NotImplementedError):
    T.activation_code("prelu")


def test_constructor_defaults_match_the_reference_signatures():
  assert T.JPEGLikeSynthesis().kernel_size == 16 and T.JPEGLikeSynthesis().strides == 16
  assert T.TwoLayerSynthesis().channels == (24, 3) and T.TwoLayerResSynthesis().channels == (12, 3)
  assert T.JPEGLikeHyperSynthesis(320).kernel_size == 6
  assert T.MBT2018Synthesis(192).n_layers == 4 and T.MBT2018Synthesis(192).output_channels == 3
  assert T.CNNSynthesis(192).activation_type == "leaky_relu"
  assert T.HyperSynthesis(320).activation_type == "relu"


def test_variable_layouts():
  v = T.HyperSynthesis(320).variable_shapes(320)
  assert v["hyper_synthesis.layer_1.kernel"] == (5, 5, 480, 320)     # Keras [kh,kw,Cout,Cin]
  assert v["hyper_synthesis.layer_2.kernel"] == (3, 3, 640, 480)
  v = T.BLS2017Synthesis(256).variable_shapes(256)
  assert v["synthesis.layer_2.kernel"] == (9, 9, 256, 3)             # tfc [kh,kw,Cin,Cout]
  assert v["synthesis.igdn_0.gamma"] == (256, 256)
  v = T.JPEGLikeSynthesis(kernel_size=18, strides=16, use_offset=True).variable_shapes(320)
  assert v["synthesis.conv.kernel"] == (18, 18, 3, 321)


def test_model_geometry_follows_the_reference():
  m = build_config("two_layer_syn")
  assert m.downsample_factor == 64 and m.latent_channels == 320 and m.hyper_channels == 320
  assert m.latent_shapes(24, 512, 768) == ((24, 8, 12, 320), (24, 32, 48, 320))
  assert m.latent_shapes(1, 500, 700) == ((1, 8, 11, 320), (1, 32, 44, 320))       # pad to a multiple of 64
  f = build_config("bls2017")
  assert isinstance(f, FactorizedModel) and f.downsample_factor == 16 and f.latent_channels == 256
  assert set(CONFIGS) == {"jpegl", "two_layer_syn", "two_layer_syn2", "mbt2018", "bls2017"}
  assert build_config("two_layer_syn2:48")._synthesis.channels == (48, 3)
  custom = Model(dict(analysis=dict(cls="ElicAnalysis", channels=(192, 192, 192, 320)),
                      synthesis=dict(cls="JPEGLikeSynthesis", kernel_size=18, strides=16),
                      hyper_synthesis=dict(cls="JPEGLikeHyperSynthesis", bottleneck_size=320, kernel_size=6)))
  assert custom.downsample_factor == 64


def test_descriptors_carry_the_kwargs():
  d = build_config("two_layer_syn")._synthesis.desc(320)
  assert d.kind == _lib.T_TWO_LAYER_RES and list(d.channels) == [12, 3] and list(d.kernel_sizes) == [13, 5]
  assert list(d.strides) == [8, 2] and d.activation == _lib.ACT_IGDN1 and d.in_channels == 320
  d = T.JPEGLikeSynthesis(kernel_size=18, strides=16).desc(320)
  assert d.kind == _lib.T_JPEG_LIKE_SYNTHESIS and d.kernel_sizes[0] == 18 and d.strides[0] == 16 and d.use_bias == 1


def test_tensor_handoff_numpy_and_dlpack():
  a = np.arange(24, dtype=np.float32).reshape(1, 2, 3, 4)
  t = as_tensor(a)
  assert t.t.device_type == _lib.DL_CPU and t.shape == (1, 2, 3, 4) and t.t.dtype_code == _lib.DL_FLOAT and t.t.dtype_bits == 32
  assert t.t.data == a.ctypes.data
  with pytest.raises(ValueError):
    as_tensor(a.transpose(0, 2, 1, 3))
  with pytest.raises(TypeError):
    as_tensor(a.astype(np.float64))
  # a DLPack producer (torch stands in for TensorFlow's tf.experimental.dlpack.to_dlpack): zero-copy, borrowed
  import torch
  x = torch.arange(24, dtype=torch.int16).reshape(1, 2, 3, 4)
  t = as_tensor(x)
  assert t.t.device_type == _lib.DL_CPU and t.shape == (1, 2, 3, 4) and t.t.dtype_code == _lib.DL_INT and t.t.dtype_bits == 16
  assert t.t.data + t.t.byte_offset == x.data_ptr()
  cap = torch.utils.dlpack.to_dlpack(torch.zeros(2, 2, 2, 2, dtype=torch.uint8))
  assert as_tensor(cap).t.dtype_code == _lib.DL_UINT


def test_cuda_array_interface_objects_are_device_tensors():
  class Fake:
    __cuda_array_interface__ = dict(shape=(1, 2, 2, 4), typestr="<f4", data=(0x7f0000000000, False), version=3, strides=None)
  t = as_tensor(Fake(), device_id=3)
  assert t.t.device_type == _lib.DL_CUDA and t.t.device_id == 3 and t.t.data == 0x7f0000000000


def test_bind_host_to_gpu_uses_the_gpus_numa_cpus(tmp_path, monkeypatch):
  """parallel.bind_host_to_gpu: CPU list of the GPU's PCI device from sysfs -> sched_setaffinity; a box without NUMA
  information (numa_node = -1, the 8-GPU VM of this round) or without sysfs is left alone and nothing raises."""
  import os
  from shallow_ntc_b200 import parallel

  class Ctx:
    pci_bus_id = "0000:1B:00.0"

  assert sorted(parallel._parse_cpulist("0-3,8,10-11\n")) == [0, 1, 2, 3, 8, 10, 11]
  dev = tmp_path / "0000:1b:00.0"
  dev.mkdir()
  allowed = sorted(os.sched_getaffinity(0))
  calls = []
  monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append(set(cpus)))
  # no NUMA information
  (dev / "numa_node").write_text("-1\n"); (dev / "local_cpulist").write_text(f"{allowed[0]}\n")
  info = parallel.bind_host_to_gpu(Ctx(), sysfs=str(tmp_path))
  assert info["bound"] is False and info["numa_node"] == -1 and not calls
  # NUMA node with a strict subset of the allowed CPUs -> bound to it
  if len(allowed) > 1:
    (dev / "numa_node").write_text("1\n"); (dev / "local_cpulist").write_text(f"{allowed[0]}\n")
    info = parallel.bind_host_to_gpu(Ctx(), sysfs=str(tmp_path))
    assert info["bound"] is True and calls == [{allowed[0]}]
  # missing sysfs entry: reported, not raised
  info = parallel.bind_host_to_gpu(Ctx(), sysfs=str(tmp_path / "nope"))
  assert info["bound"] is False and "error" in info


# ---- intra-frame band split (tiling.py): geometry against the oracle ---------------------------------------------------
def test_input_rows_formula_against_brute_force():
  """Receptive rows of a conv chain, checked by enumerating o = n*s + a - p directly."""
  from shallow_ntc_b200.tiling import input_rows
  for chain in ([(5, 2, 1)], [(5, 2, 2), (9, 4, 4)], [(13, 8, 2), (5, 2, 1)], [(5, 2, 1), (5, 2, 1), (3, 1, 1)], [(18, 16, 1)], [(6, 4, 1)]):
    for lo, hi in ((0, 0), (5, 9), (64, 127), (17, 17)):
      rows = set(range(lo, hi + 1))
      for k, s, p in reversed(chain):
        rows = {n for o in rows for a in range(k) for n in [(o + p - a) // s] if (o + p - a) % s == 0}
      assert (min(rows), max(rows)) == input_rows(chain, lo, hi), (chain, lo, hi)


@pytest.mark.parametrize("name,H,W,n_bands", [("bls2017", 200, 48, 3), ("two_layer_syn", 320, 64, 2), ("jpegl", 300, 64, 4), ("mbt2018", 192, 64, 3)])
def test_banded_oracle_decode_equals_whole_frame(name, H, W, n_bands):
  """The band plan (rows + halo) applied to the ORACLE: decoding each band's sub-tensors and keeping its rows reproduces
  the whole-frame decode (float64, so equal up to BLAS blocking: 1e-11), image rows and scale-table rows alike."""
  from oracle import ntc_oracle as O
  m = build_config(name)
  cfg = m._transform_config["synthesis"]
  kw = {k: v for k, v in cfg.items() if k != "cls"}
  w = synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=cfg["cls"])
  zs, ys = m.latent_shapes(1, H, W)
  z, q = synthetic.make_latents(zs, ys)
  dec = (lambda z_, q_, h: O.mshyper_decode(w, cfg["cls"], z_, q_, h, W, kw)) if m.hyperprior else (lambda z_, q_, h: O.factorized_decode(w, cfg["cls"], q_, h, W, kw))
  whole = dec(z, q, H)
  bands = m.band_plan((H, W), n_bands)
  assert bands[0].rows[0] == 0 and bands[-1].rows[1] == H and all(a.rows[1] == b.rows[0] for a, b in zip(bands, bands[1:]))
  for b in bands:
    zb, qb = m.band_inputs(z, q, b)
    part = dec(zb, qb, b.sub_h)
    assert np.abs(part["recon"][:, b.keep[0]:b.keep[1]] - whole["recon"][:, b.rows[0]:b.rows[1]]).max() < 1e-11
    if m.hyperprior:
      c0, c1 = b.y_core[0] - b.y_rows[0], b.y_core[1] - b.y_rows[0]
      assert np.array_equal(part["idx"][:, c0:c1], whole["idx"][:, b.y_core[0]:b.y_core[1]])
  # the halo is tight: one row less on either side changes the band's pixels
  b = bands[1]
  if not m.hyperprior:
    qs = np.ascontiguousarray(q[:, b.y_rows[0] + 1:b.y_rows[1]])
    part = dec(None, qs, b.sub_h - 16)
    assert np.abs(part["recon"][:, b.keep[0] - 16:b.keep[1] - 16] - whole["recon"][:, b.rows[0]:b.rows[1]]).max() > 1e-6
