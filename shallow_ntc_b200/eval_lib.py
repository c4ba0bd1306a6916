"""Decode-side mirror of the evaluate loop: ``Model.evaluate`` (mshyper/models.py:415-433) yields one
metrics record per image and ``eval_lib.eval_workdir`` (common/eval_lib.py:91-105) dumps them as a flat JSON list.

Here the loop starts from decoded symbols instead of images (the encoder is out of scope): every image is decoded by
libsntc -- batched, which the reference does not do (it evaluates image by image) -- and the per-image record carries
the scalars the reference's results files hold and this path can produce on the device: ``bpp``, ``psnr``, ``mse``,
``rd_loss`` (= bpp + rd_lambda * mse, :343), ``msssim`` / ``msssim_db`` (:321-332, ``Context.msssim``), ``lpips`` (:334-340, when an
``lpips.Lpips`` object holding the network's weights is passed), ``instance_id``."""
from __future__ import annotations

import json
import numpy as np


def msssim_defined(H, W):
  """Sizes tf.image.ssim / ssim_multiscale accept: every scale must hold the 11x11 window (single scale when both sides
  are < 160 px, mshyper/models.py:325-327; otherwise 5 scales, each half the size of the previous one, rounded up)."""
  scales = 1 if (H < 160 and W < 160) else 5
  for _ in range(scales - 1):
    H, W = (H + 1) // 2, (W + 1) // 2
  return H >= 11 and W >= 11


def evaluate_symbols(model, z_hat, q_y, originals, image_hw=None, batch_size=24, rd_lambda=None, extra=None, msssim=True, lpips=None):
  """Yields one dict per image, in order.  z_hat / q_y: arrays (or lists of per-image arrays of one shape) of decoded
  symbols; originals: uint8 [N,H,W,3].  ``extra``: hyper-parameters added to every record (parse_runname's role)."""
  originals = np.asarray(originals)
  N = originals.shape[0]
  H, W = image_hw if image_hw is not None else originals.shape[1:3]
  for lo in range(0, N, batch_size):
    hi = min(N, lo + batch_size)
    z = None if z_hat is None else np.ascontiguousarray(z_hat[lo:hi])
    out = model.decompress(z, np.ascontiguousarray(q_y[lo:hi]), (H, W), original=np.ascontiguousarray(originals[lo:hi]),
                           return_bits=model.hyperprior)
    ms = None
    if msssim and msssim_defined(H, W):
      ms = model.ctx.msssim(np.ascontiguousarray(originals[lo:hi]), out["image"])
    lp = lpips(np.ascontiguousarray(originals[lo:hi]), out["image"]) if lpips is not None else None   # lpips_model([image_batch, reconstruction])
    for i in range(hi - lo):
      rec = dict(instance_id=lo + i, psnr=float(out["psnr"][i]), mse=float(out["mse"][i]))
      if ms is not None:
        rec.update(msssim=float(ms[0][i]), msssim_db=float(ms[1][i]))
      if lp is not None:
        rec["lpips"] = float(lp[i])
      if "bpp" in out:
        rec.update(bpp=float(out["bpp"][i]), latent_bpp=float(out["bits_y"][i] / (H * W)), hyper_latent_bpp=float(out["bits_z"][i] / (H * W)))
        if rd_lambda is not None:
          rec["rd_loss"] = rec["bpp"] + float(rd_lambda) * rec["mse"]        # mshyper/models.py:343
      if extra:
        rec.update(extra)
      yield rec


def dump_json(records, path):
  """common/eval_lib.py:103 (utils.dump_json): a flat list of per-image dicts."""
  with open(path, "w") as f:
    json.dump(list(records), f, indent=2)
  return path
