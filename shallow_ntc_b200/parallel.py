"""Multi-GPU driver logic: one process per GPU, images sharded as independent units, no data-path
collective; one all-reduce of the metric sums at the very end.  ``dist`` is any object with
the torch.distributed API (NCCL on the GPU box, gloo in the CPU tests) -- plumbing, not the product."""
from __future__ import annotations

import numpy as np

from .synthetic import shard_range  # noqa: F401  (re-exported)


def reduce_metric_sums(dist, sums, device=None):
  """all-reduce(SUM) of [sum psnr, sum mse, sum bits_y, sum bits_z, n_images] (float64).
  Reference semantics: per-image metrics, then an arithmetic mean (mshyper/models.py:306-317,
  train_lib.py:64-68)."""
  sums = np.asarray(sums, dtype=np.float64)
  if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
    return sums
  import torch
  t = torch.tensor(sums, dtype=torch.float64, device=device if device is not None else "cpu")
  dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return t.cpu().numpy()


def max_over_ranks(dist, values, device=None):
  """Device times are reported as the max over ranks."""
  values = np.asarray(values, dtype=np.float64)
  if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
    return values
  import torch
  t = torch.tensor(values, dtype=torch.float64, device=device if device is not None else "cpu")
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return t.cpu().numpy()


def mean_metrics(sums):
  n = max(float(sums[-1]), 1.0)
  return dict(psnr=float(sums[0] / n), mse=float(sums[1] / n), n_images=int(sums[-1]))
