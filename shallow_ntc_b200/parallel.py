"""Multi-GPU driver logic: one process per GPU, images sharded as independent units, no data-path
collective; one all-reduce of the metric sums at the very end.  ``dist`` is any object with
the torch.distributed API (NCCL on the GPU box, gloo in the CPU tests) -- plumbing, not the product."""
from __future__ import annotations

import os

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


def _parse_cpulist(text):
  cpus = set()
  for part in text.strip().split(","):
    if not part:
      continue
    lo, _, hi = part.partition("-")
    cpus.update(range(int(lo), int(hi or lo) + 1))
  return cpus


def bind_host_to_gpu(ctx, sysfs="/sys/bus/pci/devices"):
  """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that the page-locked staging buffers it
  allocates afterwards (first touch) and the thread that enqueues the copies are local to that GPU's PCIe root.  On an
  8-GPU box every rank otherwise lands wherever the launcher put it and the host<->device streams of all ranks cross the
  socket interconnect.  Returns a dict describing what was done; never raises (a box without sysfs NUMA info is left alone)."""
  info = dict(bound=False)
  try:
    bus = ctx.pci_bus_id.lower()                       # "0000:1b:00.0"
    info["pci_bus_id"] = bus
    base = os.path.join(sysfs, bus)
    node = int(open(os.path.join(base, "numa_node")).read().strip())
    info["numa_node"] = node
    cpus = _parse_cpulist(open(os.path.join(base, "local_cpulist")).read())
    allowed = os.sched_getaffinity(0)
    cpus &= allowed
    if node < 0 or not cpus or cpus == allowed:
      return info
    os.sched_setaffinity(0, cpus)
    info.update(bound=True, cpus=len(cpus))
  except Exception as e:  # sysfs missing (containers), permission, exotic topology: run unbound
    info["error"] = f"{type(e).__name__}: {e}"
  return info
