"""LPIPS.  The reference holds a real known answer for it -- lpips_tf2/test.py:17-19 records the official values
0.569 / 0.422 for the image pairs it ships -- so this is the one place where the oracle is pinned to numbers the reference itself
states, with the reference's own weights (oracle/_ref/lpips_weights.npz, made from the vendored checkpoints by oracle/make_ref.py)."""
import os

import numpy as np
import pytest

from oracle import ntc_oracle as O
from oracle import make_ref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lpips_reference.npz"))


def _real_weights():
  w = make_ref.load_lpips_weights()
  if w is None:
    pytest.skip("neither oracle/_ref/lpips_weights.npz nor the reference tree is available")
  return w


def test_oracle_lpips_reproduces_the_references_known_answers():
  w = _real_weights()
  a = np.stack([GOLD["ex_ref"], GOLD["ex_ref"]])
  b = np.stack([GOLD["ex_p0"], GOLD["ex_p1"]])
  val, layers = O.lpips(w, a, b, return_layers=True)
  assert [round(float(v), 3) for v in val] == GOLD["official"].tolist() == [0.569, 0.422]      # lpips_tf2/test.py:17-19
  assert np.abs(val - GOLD["oracle_value"]).max() < 1e-12 and np.abs(layers - GOLD["oracle_layers"]).max() < 1e-12
  assert np.abs(O.lpips(w, a, a)).max() == 0.0                                                   # identical images
  f32 = O.lpips(w, a, b, dtype=np.float32)
  assert np.abs(f32 - val).max() < 1e-5


def test_checkpoint_reader_on_the_vendored_lpips_checkpoints():
  root = "/root/reference/lpips_tf2/models"
  if not os.path.isdir(root):
    pytest.skip("reference tree not mounted")
  from oracle import tf_checkpoint
  entries, shards = tf_checkpoint.read_index(os.path.join(root, "vgg", "exported"))
  assert shards == 2 and sum(k.endswith("VARIABLE_VALUE") and "layer_with_weights" in k for k in entries) == 26
  w = O.lpips_weights_from_reference_checkpoints(os.path.join(root, "vgg", "exported"), os.path.join(root, "lin", "exported"))
  assert {k: tuple(v.shape) for k, v in w.items()} == O.lpips_variable_shapes()
  ref = make_ref.load_lpips_weights()
  assert all(np.array_equal(w[k], ref[k]) for k in w)
  assert sum(int(np.prod(v.shape)) for v in w.values()) == 14714688 + 1472              # VGG16 without top + the five lin layers


def test_oracle_lpips_against_an_independent_torch_statement():
  """conv2d / max_pool2d from torch (float64) on random weights, odd image size (pooling drops the trailing row / column)."""
  import torch
  import torch.nn.functional as F
  from shallow_ntc_b200 import lpips as L
  assert L.variable_shapes() == O.lpips_variable_shapes()
  w = L.random_weights()
  rng = np.random.default_rng(5)
  a = rng.integers(0, 256, size=(2, 37, 53, 3)).astype(np.uint8)
  b = np.clip(a.astype(int) + rng.integers(-40, 41, size=a.shape), 0, 255).astype(np.uint8)
  val, layers = O.lpips(w, a, b, return_layers=True)

  def feats(im):
    x = torch.from_numpy(im.astype(np.float64)).permute(0, 3, 1, 2) / 127.5 - 1.0
    x = (x - torch.tensor(O.LPIPS_SHIFT, dtype=torch.float64).view(1, 3, 1, 1)) / torch.tensor(O.LPIPS_SCALE, dtype=torch.float64).view(1, 3, 1, 1)
    out, i = [], 0
    for bi, block in enumerate(O.VGG_BLOCKS):
      if bi:
        x = F.max_pool2d(x, 2)
      for _ in block:
        k = torch.from_numpy(w[f"lpips.conv_{i}.kernel"].astype(np.float64)).permute(3, 2, 0, 1)
        x = F.relu(F.conv2d(x, k, torch.from_numpy(w[f"lpips.conv_{i}.bias"].astype(np.float64)), padding=1))
        i += 1
      out.append(x)
    return out
  tot = 0
  for l, (fa, fb) in enumerate(zip(feats(a), feats(b))):
    na, nb = fa * torch.rsqrt((fa * fa).sum(1, keepdim=True)), fb * torch.rsqrt((fb * fb).sum(1, keepdim=True))
    d = ((na - nb) ** 2 * torch.from_numpy(w[f"lpips.lin_{l}.kernel"].astype(np.float64)).view(1, -1, 1, 1)).sum(1).mean((1, 2))
    assert np.abs(d.numpy() - layers[:, l]).max() < 1e-12
    tot = tot + d
  assert np.abs(tot.numpy() - val).max() < 1e-12


# --------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["tc", "fp32"])
def test_gpu_lpips_matches_oracle_on_random_weights(gpu_ctx, precision):
  from shallow_ntc_b200 import lpips as L
  w = L.random_weights()
  rng = np.random.default_rng(6)
  a = rng.integers(0, 256, size=(3, 100, 150, 3)).astype(np.uint8)
  b = np.clip(a.astype(int) + rng.integers(-30, 31, size=a.shape), 0, 255).astype(np.uint8)
  ref, ref_layers = O.lpips(w, a, b, return_layers=True)
  net = L.Lpips(gpu_ctx, w, precision)
  n0 = gpu_ctx.launch_counts
  got, layers = net(a, b, return_layers=True)
  n = {k: v - n0[k] for k, v in gpu_ctx.launch_counts.items()}
  # differences of two unit vectors cancel: feature errors of ~1e-6 (13 layers of fp32-class accumulation) show up ~10-100x larger
  tol = 2e-4 if precision == "tc" else 5e-5
  assert np.abs(got / ref - 1).max() < tol and np.abs(layers / ref_layers - 1).max() < 2 * tol, (got, ref)
  assert (n["band_tc"] == 12 and n["band_f32"] == 1) if precision == "tc" else (n["band_tc"] == 0 and n["band_f32"] == 13), n
  # float32 images in [0, 255], device-resident inputs, identical images, determinism
  assert np.array_equal(net(a.astype(np.float32), b.astype(np.float32)), got)
  assert np.array_equal(net(gpu_ctx.to_device(a), gpu_ctx.to_device(b)), got)
  assert np.abs(net(a, a)).max() == 0.0
  assert np.array_equal(net(a[1:2], b[1:2]), got[1:2])


@pytest.mark.gpu
def test_gpu_lpips_reproduces_the_references_known_answers(gpu_ctx):
  """The reference's own weights and test images: 0.569 / 0.422 (lpips_tf2/test.py:17-19)."""
  from shallow_ntc_b200 import lpips as L
  w = _real_weights()
  a = np.stack([GOLD["ex_ref"], GOLD["ex_ref"]])
  b = np.stack([GOLD["ex_p0"], GOLD["ex_p1"]])
  for precision in ("tc", "fp32"):
    got, layers = L.Lpips(gpu_ctx, w, precision)(a, b, return_layers=True)
    assert [round(float(v), 3) for v in got] == [0.569, 0.422], (precision, got)
    assert np.abs(got - GOLD["oracle_value"]).max() < 1e-4 and np.abs(layers - GOLD["oracle_layers"]).max() < 5e-5, (precision, got)


@pytest.mark.gpu
def test_evaluate_records_carry_lpips(gpu_ctx):
  from shallow_ntc_b200 import lpips as L, synthetic
  from helpers import make_case
  H, W = 64, 128
  model, wts, z, q = make_case("two_layer_syn", 3, H, W, "stress", "tc", gpu_ctx)
  img = model.decompress(z, q, (H, W))["image"]
  orig = synthetic.make_original(img)
  w = L.random_weights()
  recs = list(model.evaluate(z, q, orig, batch_size=2, lpips=L.Lpips(gpu_ctx, w)))
  ref = O.lpips(w, orig, img)
  assert all(abs(r["lpips"] / ref[i] - 1) < 2e-4 for i, r in enumerate(recs))
  assert full_size_smoke(gpu_ctx, w)


def full_size_smoke(gpu_ctx, w):
  """A 512 x 768 pair in several passes' worth of workspace: finite, positive, and equal to the sum of its layer terms."""
  from shallow_ntc_b200 import lpips as L
  rng = np.random.default_rng(7)
  a = rng.integers(0, 256, size=(2, 512, 768, 3)).astype(np.uint8)
  b = np.clip(a.astype(int) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8)
  got, layers = L.Lpips(gpu_ctx, w)(a, b, return_layers=True)
  return bool(np.all(np.isfinite(got)) and np.all(got > 0) and np.allclose(layers.sum(1), got, rtol=1e-12))
