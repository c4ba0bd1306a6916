"""LPIPS on the device: the ``lpips`` column of the evaluate records.

``Lpips(ctx, weights)(image_batch, reconstruction)`` is ``learned_perceptual_metric_model(im_size)([image_batch, reconstruction])``
of the vendored lpips_tf2 (``lpips_tf2/lpips_tensorflow.py:14-72``), as called in validation mode by
``mshyper/models.py:334-340`` / ``factorized/models.py:158-164``: images [B,H,W,3] in [0, 255] (uint8 or float32), one value per
image.  The 13 VGG16 convolutions run on the band-GEMM kernels of the decode path (``csrc/sntc_kernels_lpips.cuh``).

Weights: ``lpips.conv_i.kernel`` [3,3,Cin,Cout] / ``lpips.conv_i.bias`` (Keras Conv2D layout, the 13 VGG16 convs in order) and
``lpips.lin_l.kernel`` [C_l] (the five 1x1 convs of the linear model).  From the reference's Keras models:
``{f"lpips.conv_{i}.kernel": l.kernel.numpy() ...}`` over ``[l for l in vgg16.layers if l.weights]`` and
``lin.get_layer(f"lin{l}").kernel.numpy().reshape(-1)``.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from ._lib import lib, check
from .tensors import Context, as_tensor

VGG_BLOCKS = ((64, 64), (128, 128), (256, 256, 256), (512, 512, 512), (512, 512, 512))


def variable_shapes() -> dict:
  v, cin, i = {}, 3, 0
  for block in VGG_BLOCKS:
    for cout in block:
      v[f"lpips.conv_{i}.kernel"] = (3, 3, cin, cout)
      v[f"lpips.conv_{i}.bias"] = (cout,)
      cin, i = cout, i + 1
  for l, block in enumerate(VGG_BLOCKS):
    v[f"lpips.lin_{l}.kernel"] = (block[-1],)
  return v


def random_weights(seed=4321) -> dict:
  """He-normal kernels, small biases, positive lin weights (the trained ones are non-negative): random-init stand-in for tests."""
  rng = np.random.default_rng(seed)
  w = {}
  for name, shape in variable_shapes().items():
    if name.endswith(".kernel") and len(shape) == 4:
      w[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / (9 * shape[2]))).astype(np.float32)
    elif name.endswith(".bias"):
      w[name] = rng.uniform(-0.05, 0.05, size=shape).astype(np.float32)
    else:
      w[name] = rng.uniform(0.0, 0.5, size=shape).astype(np.float32)
  return w


class Lpips:
  def __init__(self, ctx: Context, weights: dict, precision: str = "tc"):
    self.ctx = ctx
    self.handle = C.c_void_p()
    prec = {"fp32": _lib.PRECISION_FP32, "tc": _lib.PRECISION_TC_F16X3}[precision]
    check(lib.sntc_lpips_create(ctx.handle, prec, C.byref(self.handle)))
    for name, shape in variable_shapes().items():
      if name not in weights:
        raise KeyError(f"missing variable {name} {shape}")
      w = np.ascontiguousarray(weights[name], dtype=np.float32)
      shp = (C.c_int64 * w.ndim)(*w.shape)
      check(lib.sntc_lpips_load_weights(self.handle, name.encode(), w.ctypes.data_as(C.POINTER(C.c_float)), shp, w.ndim))
    check(lib.sntc_lpips_finalize(self.handle))

  def __call__(self, image_a, image_b, return_layers=False):
    """(lpips [B]) or (lpips [B], per-layer terms [B, 5]) as float64 numpy; inputs numpy / DeviceArray / DLPack, uint8 or float32."""
    a, b = as_tensor(image_a, self.ctx.device), as_tensor(image_b, self.ctx.device)
    B = int(a.shape[0])
    out = (C.c_double * max(B, 1))()
    lay = (C.c_double * max(5 * B, 1))()
    check(lib.sntc_image_lpips(self.handle, a.byref(), b.byref(), out, lay, None))
    val = np.array(out[:B], dtype=np.float64)
    if return_layers:
      return val, np.array(lay[:5 * B], dtype=np.float64).reshape(B, 5)
    return val

  def __del__(self):
    try:
      if self.handle and self.ctx.handle:
        lib.sntc_lpips_destroy(self.handle)
    except Exception:
      pass
