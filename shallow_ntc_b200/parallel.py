"""Multi-GPU driver logic: one process per GPU, images sharded as independent units, no data-path
collective; one all-reduce of the metric sums at the very end.

On the GPU box the collective is NCCL called by libsntc itself (``NcclGroup``: ``sntc_comm_*`` / ``sntc_allreduce_metrics``,
NVLink 5 / NVSwitch underneath) -- no PyTorch anywhere in the product.  The reduce helpers also accept any object with the
``torch.distributed`` API (gloo in the CPU tests of the sharding logic); that module is the CALLER's, nothing here imports it."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from .synthetic import shard_range  # noqa: F401  (re-exported)


def exchange_bytes(path: str, rank: int, payload: bytes | None, nbytes: int, timeout_s: float = 120.0) -> bytes:
  """Single-node rendezvous through the file system: rank 0 publishes `payload` atomically (write + rename), every other rank
  polls until the file has `nbytes` bytes.  Used to hand the 128-byte NCCL unique id to the ranks a launcher started."""
  if rank == 0:
    tmp = f"{path}.{os.getpid()}.tmp"
    with open(tmp, "wb") as f:
      f.write(payload)
    os.replace(tmp, path)
    return payload
  t0 = time.time()
  while True:
    try:
      data = open(path, "rb").read()
      if len(data) == nbytes:
        return data
    except FileNotFoundError:
      pass
    if time.time() - t0 > timeout_s:
      raise TimeoutError(f"rank {rank}: no rendezvous file {path} after {timeout_s:.0f} s")
    time.sleep(0.01)


class NcclGroup:
  """The ranks of one launch (one process per GPU) with the collective done by libsntc over NCCL."""

  def __init__(self, ctx, rank: int, world: int, id_path: str):
    from ._lib import lib, check, COMM_ID_BYTES
    self.ctx, self.rank, self.world, self.id_path = ctx, int(rank), int(world), id_path
    uid = None
    if self.rank == 0:
      buf = C.create_string_buffer(COMM_ID_BYTES)
      check(lib.sntc_comm_unique_id(buf))
      uid = buf.raw
    uid = exchange_bytes(id_path, self.rank, uid, COMM_ID_BYTES)
    self.handle = C.c_void_p()
    check(lib.sntc_comm_create(ctx.handle, uid, self.rank, self.world, C.byref(self.handle)))

  @classmethod
  def from_env(cls, ctx):
    """RANK / WORLD_SIZE as exported by ``python -m torch.distributed.run`` (or any launcher).  The id file is keyed by the
    launcher's pid and MASTER_PORT, so concurrent launches on one box do not collide."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    key = f"{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', 'none')}_{os.getppid()}"
    return cls(ctx, rank, world, os.path.join(os.environ.get("SNTC_RENDEZVOUS_DIR", "/tmp"), f"sntc_nccl_{key}.id"))

  def _allreduce(self, values, op):
    from ._lib import lib, check
    v = np.ascontiguousarray(values, dtype=np.float64).copy()
    check(lib.sntc_comm_allreduce_f64(self.handle, v.ctypes.data_as(C.POINTER(C.c_double)), int(v.size), op))
    return v

  def allreduce_sum(self, values):
    from ._lib import REDUCE_SUM
    return self._allreduce(values, REDUCE_SUM)

  def allreduce_max(self, values):
    from ._lib import REDUCE_MAX
    return self._allreduce(values, REDUCE_MAX)

  def allreduce_metrics(self, sums5):
    """sntc_allreduce_metrics: [sum psnr, sum mse, sum bits_y, sum bits_z, n_images]."""
    from ._lib import lib, check
    v = np.ascontiguousarray(sums5, dtype=np.float64).copy()
    assert v.size == 5
    check(lib.sntc_allreduce_metrics(self.handle, v.ctypes.data_as(C.POINTER(C.c_double))))
    return v

  def barrier(self):
    self.allreduce_sum(np.zeros(1))

  def close(self):
    from ._lib import lib
    if getattr(self, "handle", None):
      lib.sntc_comm_destroy(self.handle)
      self.handle = None
      if self.rank == 0:
        try:
          os.remove(self.id_path)
        except OSError:
          pass

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


def _caller_tensor(dist, values, device):
  """float64 tensor of the framework `dist` belongs to (the caller imported it; this package never does)."""
  import sys
  fw = sys.modules[dist.__name__.split(".")[0]]
  return fw.tensor(values, dtype=fw.float64, device=device if device is not None else "cpu")


def reduce_metric_sums(dist, sums, device=None):
  """all-reduce(SUM) of [sum psnr, sum mse, sum bits_y, sum bits_z, n_images] (float64).
  Reference semantics: per-image metrics, then an arithmetic mean (mshyper/models.py:306-317,
  train_lib.py:64-68).  ``dist``: a NcclGroup, None (single process), or a torch.distributed-like module."""
  sums = np.asarray(sums, dtype=np.float64)
  if isinstance(dist, NcclGroup):
    return dist.allreduce_metrics(sums) if sums.size == 5 else dist.allreduce_sum(sums)
  if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
    return sums
  t = _caller_tensor(dist, sums, device)
  dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return t.cpu().numpy()


def max_over_ranks(dist, values, device=None):
  """Device times are reported as the max over ranks."""
  values = np.asarray(values, dtype=np.float64)
  if isinstance(dist, NcclGroup):
    return dist.allreduce_max(values)
  if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
    return values
  t = _caller_tensor(dist, values, device)
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
