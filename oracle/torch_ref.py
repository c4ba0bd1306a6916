"""torch-CPU (oneDNN) statement of the decode path.  TEST INFRASTRUCTURE ONLY -- a second CPU stand-in for the reference's
TF-2.10 CPU decode: TensorFlow's CPU ``Conv2DBackpropInput`` runs on oneDNN too, so
``torch.nn.functional.conv_transpose2d`` in float32 is the closest thing to it that can run here.  Used by ``bench.py``'s
CPU arms and checked against ``oracle/ntc_oracle.py`` in ``tests/test_oracle.py``.  Never imported by the product.

Reference lines restated: ``mshyper/models.py:269-298, 313-314`` (hyper-synthesis, split / exp, q + mu, synthesis, crop,
pixels), ``factorized/models.py:101-141``, layer stacks ``common/transforms.py:115-134, 158-175, 222-232, 265-361``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _nchw(x):
  return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)


class _ConvT:
  """out[o] += in[n] * W[a], o = n*s + a - p, cropped to n_in*s (A1 / A2 of the oracle): conv_transpose2d without padding,
  then the crop.  ``kernel`` comes in the reference's layout."""

  def __init__(self, kernel, bias, s, keras: bool):
    k = kernel.shape[0]
    w = np.asarray(kernel, dtype=np.float32)
    # torch wants [Cin, Cout, kh, kw]; Keras kernel is [kh, kw, Cout, Cin], tfc SignalConv2D [kh, kw, Cin, Cout]
    w = w.transpose(3, 2, 0, 1) if keras else w.transpose(2, 3, 0, 1)
    self.w = torch.from_numpy(np.ascontiguousarray(w))
    self.b = torch.from_numpy(np.asarray(bias, dtype=np.float32)) if bias is not None else None
    self.s = s
    self.p = max(k - s, 0) // 2 if keras else (k - 1) // 2

  def __call__(self, x):
    n_h, n_w = x.shape[2] * self.s, x.shape[3] * self.s
    y = F.conv_transpose2d(x, self.w, self.b, stride=self.s)
    return y[:, :, self.p:self.p + n_h, self.p:self.p + n_w]


def _gdn1(x, beta, gamma, inverse=True):
  """common/transforms.py:27-63: norm = beta + |x| @ gamma (a 1x1 convolution over channels)."""
  norm = F.conv2d(x.abs(), gamma, beta)
  return x * norm if inverse else x / norm


class TorchDecoder:
  """Weights converted once (as a framework would hold them); ``__call__`` is the timed decode."""

  def __init__(self, cfg: dict, wts: dict):
    self.cfg = cfg
    syn = cfg["synthesis"]
    self.cls = syn["cls"]
    g = lambda n: wts[n]
    if cfg["hyperprior"]:
      self.hyper = [_ConvT(g(f"hyper_synthesis.layer_{i}.kernel"), g(f"hyper_synthesis.layer_{i}.bias"), s, True) for i, s in enumerate((2, 2, 1))]
    gd = lambda pre: (torch.from_numpy(np.asarray(g(pre + ".beta"), np.float32)),
                      torch.from_numpy(np.ascontiguousarray(np.asarray(g(pre + ".gamma"), np.float32).T[:, :, None, None])))   # conv2d weight [out, in, 1, 1]
    if self.cls == "JPEGLikeSynthesis":
      self.layers = [_ConvT(g("synthesis.conv.kernel"), g("synthesis.conv.bias"), syn["strides"], True)]
    elif self.cls == "TwoLayerResSynthesis":
      s1, s2 = syn["strides"]
      self.base = _ConvT(g("synthesis.base_conv.kernel"), g("synthesis.base_conv.bias"), s1, True)
      self.res = _ConvT(g("synthesis.res.kernel"), g("synthesis.res.bias"), s1, True)
      self.act = gd("synthesis.activation")
      self.out = _ConvT(g("synthesis.out_conv.kernel"), g("synthesis.out_conv.bias"), s2, True)
    elif self.cls == "TwoLayerSynthesis":
      s1, s2 = syn["strides"]
      self.base = _ConvT(g("synthesis.conv1.kernel"), g("synthesis.conv1.bias"), s1, True)
      self.act = gd("synthesis.activation")
      self.out = _ConvT(g("synthesis.conv2.kernel"), g("synthesis.conv2.bias"), s2, True)
    elif self.cls in ("MBT2018Synthesis", "BLS2017Synthesis"):
      strides = (2, 2, 2, 2) if self.cls == "MBT2018Synthesis" else (2, 2, 4)
      self.layers = [_ConvT(g(f"synthesis.layer_{i}.kernel"), g(f"synthesis.layer_{i}.bias"), s, False) for i, s in enumerate(strides)]
      self.gdns = [gd(f"synthesis.igdn_{i}") for i in range(len(strides) - 1)]
    else:
      raise KeyError(self.cls)

  def synthesis(self, y):
    if self.cls == "JPEGLikeSynthesis":
      return self.layers[0](y)
    if self.cls == "TwoLayerResSynthesis":
      return self.out(_gdn1(self.base(y), *self.act) + self.res(y))
    if self.cls == "TwoLayerSynthesis":
      return self.out(_gdn1(self.base(y), *self.act))
    x = y
    for i, layer in enumerate(self.layers):
      x = layer(x)
      if i < len(self.gdns):
        x = _gdn1(x, *self.gdns[i])
    return x

  @torch.no_grad()
  def __call__(self, z_hat, q_y, H, W, index_rounding="trunc"):
    out = {}
    q = _nchw(q_y)
    if self.cfg["hyperprior"]:
      x = _nchw(z_hat)
      for i, layer in enumerate(self.hyper):
        x = layer(x)
        if i < 2:
          x = F.relu(x)
      cy = x.shape[1] // 2
      mu, raw = x[:, :cy], x[:, cy:]
      i_c = torch.exp(raw).clamp(0.0, 63.0)
      idx = torch.floor(i_c) if index_rounding == "trunc" else torch.round(i_c)
      out["idx"] = idx.to(torch.uint8).permute(0, 2, 3, 1).contiguous().numpy()
      y_hat = q + mu
    else:
      y_hat = q
    x = self.synthesis(y_hat)[:, :, :H, :W]
    v = (x + 0.5) * 255.0
    out["recon"] = x.permute(0, 2, 3, 1).contiguous().numpy()
    out["image"] = torch.round(v).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().numpy()
    return out
