"""Context, device/pinned buffers and zero-copy tensor hand-off (numpy, DLPack capsules,
``__cuda_array_interface__``) for libsntc.  Producer-agnostic: a TensorFlow tensor goes through
``tf.experimental.dlpack.to_dlpack(t)``; anything exposing ``__dlpack__`` works unchanged."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from ._lib import lib, check, Tensor

_DTYPES = {
  np.dtype(np.float32): (_lib.DL_FLOAT, 32), np.dtype(np.uint8): (_lib.DL_UINT, 8),
  np.dtype(np.int16): (_lib.DL_INT, 16), np.dtype(np.int8): (_lib.DL_INT, 8),
}
_DTYPES_INV = {v: k for k, v in _DTYPES.items()}


class Context:
  """One per GPU (sntc_create).  Raises SntcError when no sm_100 device is usable."""

  def __init__(self, device: int = 0):
    h = C.c_void_p()
    check(lib.sntc_create(int(device), C.byref(h)))
    self.handle = h
    self.device = int(device)

  def close(self):
    if getattr(self, "handle", None):
      lib.sntc_destroy(self.handle)
      self.handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def sync(self):
    check(lib.sntc_sync(self.handle))

  @property
  def stream(self):
    return lib.sntc_stream(self.handle)

  @property
  def name(self):
    buf = C.create_string_buffer(256)
    check(lib.sntc_device_name(self.handle, buf, 256))
    return buf.value.decode()

  @property
  def pci_bus_id(self) -> str:
    buf = C.create_string_buffer(32)
    check(lib.sntc_device_pci_bus_id(self.handle, buf, 32))
    return buf.value.decode()

  @property
  def launch_count(self) -> int:
    return int(lib.sntc_launch_count(self.handle))

  @property
  def launch_counts(self) -> dict:
    """Kernels launched so far by family (sntc_launch_counts): total, band_tc (tcgen05 band GEMM), band_f32 (FFMA band GEMM),
    tail_mma, tail_tc, final_f32.  Lets a caller see which path served a ``precision='tc'`` model."""
    out = (C.c_uint64 * len(_lib.LAUNCH_KINDS))()
    check(lib.sntc_launch_counts(self.handle, out))
    return {k: int(out[i]) for i, k in enumerate(_lib.LAUNCH_KINDS)}

  def msssim(self, a_u8, b_u8):
    """Per-image MS-SSIM of two uint8 batches [B,H,W,C] (numpy, DeviceArray, DLPack ...) computed on the device the way
    the reference's validation branch does (mshyper/models.py:321-332: tf.image.ssim_multiscale(max_val=255.), or
    tf.image.ssim when both sides are < 160 px).  Returns (msssim [B], msssim_db [B]) as float64 numpy arrays."""
    a, b = as_tensor(a_u8, self.device), as_tensor(b_u8, self.device)
    B = int(a.shape[0])
    out = (C.c_double * max(B, 1))()
    check(lib.sntc_image_msssim(self.handle, a.byref(), b.byref(), out, None))
    val = np.array(out[:B], dtype=np.float64)
    with np.errstate(divide="ignore"):
      db = -10.0 * np.log(1.0 - val) / np.log(10.0)        # :330
    return val, db

  # --- memory ---
  def empty(self, shape, dtype) -> "DeviceArray":
    return DeviceArray(self, shape, dtype)

  def to_device(self, arr: np.ndarray) -> "DeviceArray":
    arr = np.ascontiguousarray(arr)
    d = DeviceArray(self, arr.shape, arr.dtype)
    d.copy_from_host(arr)
    return d

  def pinned_empty(self, shape, dtype, write_combined=False) -> np.ndarray:
    """numpy array backed by page-locked host memory (cudaHostAlloc).  ``write_combined``: for buffers the host only writes
    (symbols on their way to the device); reading such an array on the host is very slow."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib.sntc_host_alloc_flags(self.handle, max(n, 1), _lib.HOST_WRITE_COMBINED if write_combined else 0, C.byref(p)))
    owner = _PinnedOwner(self, p)
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = owner
    return arr

  def pinned_like(self, arr: np.ndarray) -> np.ndarray:
    out = self.pinned_empty(arr.shape, arr.dtype)
    out[...] = arr
    return out

  # --- events (CUDA-event timing on the launching stream) ---
  def event(self) -> "Event":
    return Event(self)


_PINNED = {}


class _PinnedOwner:
  def __init__(self, ctx, p):
    self.ctx, self.p = ctx, p

  def __del__(self):
    try:
      if self.ctx.handle:
        lib.sntc_host_free(self.ctx.handle, self.p)
    except Exception:
      pass


class Event:
  def __init__(self, ctx):
    self.ctx = ctx
    h = C.c_void_p()
    check(lib.sntc_event_create(ctx.handle, C.byref(h)))
    self.handle = h

  def record(self, stream=None):
    check(lib.sntc_event_record(self.ctx.handle, self.handle, stream))

  def elapsed_ms(self, stop: "Event") -> float:
    ms = C.c_float()
    check(lib.sntc_event_elapsed_ms(self.ctx.handle, self.handle, stop.handle, C.byref(ms)))
    return float(ms.value)

  def __del__(self):
    try:
      if self.ctx.handle and self.handle:
        lib.sntc_event_destroy(self.ctx.handle, self.handle)
    except Exception:
      pass


class DeviceArray:
  """Dense device buffer owned by a Context; exposes ``__cuda_array_interface__`` (v3)."""

  def __init__(self, ctx: Context, shape, dtype):
    self.ctx = ctx
    self.shape = tuple(int(s) for s in shape)
    self.dtype = np.dtype(dtype)
    self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
    p = C.c_void_p()
    check(lib.sntc_malloc(ctx.handle, max(self.nbytes, 1), C.byref(p)))
    self.ptr = p

  @property
  def __cuda_array_interface__(self):
    return dict(shape=self.shape, typestr=self.dtype.str, data=(self.ptr.value, False), version=3, strides=None)

  def __dlpack__(self, stream=None):
    return to_dlpack(self)

  def __dlpack_device__(self):
    return (_lib.DL_CUDA, self.ctx.device)

  def copy_from_host(self, arr: np.ndarray, stream=None):
    arr = np.ascontiguousarray(arr, dtype=self.dtype)
    assert arr.shape == self.shape, (arr.shape, self.shape)
    check(lib.sntc_memcpy_h2d(self.ctx.handle, self.ptr, arr.ctypes.data_as(C.c_void_p), self.nbytes, stream))
    self.ctx.sync()

  def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
      out = np.empty(self.shape, dtype=self.dtype)
    check(lib.sntc_memcpy_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p), self.ptr, self.nbytes, None))
    self.ctx.sync()
    return out

  def fill_bytes(self, value: int):
    check(lib.sntc_memset(self.ctx.handle, self.ptr, int(value), self.nbytes, None))

  def free(self):
    if self.ptr is not None and self.ctx.handle:
      lib.sntc_free(self.ctx.handle, self.ptr)
    self.ptr = None

  def __del__(self):
    try:
      self.free()
    except Exception:
      pass


class DeviceView:
  """Non-owning dense view of a byte range of a DeviceArray (``base`` keeps the allocation alive): lets several tensors share
  one allocation, so that one host<->device copy moves all of them (DecodePipeline's packed staging buffers)."""

  def __init__(self, base: DeviceArray, byte_offset: int, shape, dtype):
    self.base, self.ctx = base, base.ctx
    self.shape = tuple(int(s) for s in shape)
    self.dtype = np.dtype(dtype)
    self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
    if byte_offset < 0 or byte_offset + self.nbytes > base.nbytes:
      raise ValueError("view exceeds the allocation")
    self.ptr = C.c_void_p(base.ptr.value + int(byte_offset))

  @property
  def __cuda_array_interface__(self):
    return dict(shape=self.shape, typestr=self.dtype.str, data=(self.ptr.value, False), version=3, strides=None)

  def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
      out = np.empty(self.shape, dtype=self.dtype)
    check(lib.sntc_memcpy_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p), self.ptr, self.nbytes, None))
    self.ctx.sync()
    return out


# --- DLPack capsule parsing (borrowed: we never call the deleter, never rename the capsule) ---
class _DLManagedTensor(C.Structure):
  _fields_ = [("dl_tensor", Tensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


# Private prototypes: ctypes.pythonapi's function objects are process-wide singletons whose argtypes any other package may
# reassign, so none of them is configured or relied upon here.
_capsule_is_valid = C.PYFUNCTYPE(C.c_int, C.py_object, C.c_char_p)(("PyCapsule_IsValid", C.pythonapi))
_capsule_get_pointer = C.PYFUNCTYPE(C.c_void_p, C.py_object, C.c_char_p)(("PyCapsule_GetPointer", C.pythonapi))
_capsule_get_pointer_raw = C.PYFUNCTYPE(C.c_void_p, C.c_void_p, C.c_char_p)(("PyCapsule_GetPointer", C.pythonapi))
_capsule_get_name_raw = C.PYFUNCTYPE(C.c_char_p, C.c_void_p)(("PyCapsule_GetName", C.pythonapi))
_capsule_new = C.PYFUNCTYPE(C.py_object, C.c_void_p, C.c_char_p, C.c_void_p)(("PyCapsule_New", C.pythonapi))


class TensorRef:
  """An ``sntc_tensor`` plus whatever must stay alive while the call runs."""

  def __init__(self, t: Tensor, keep):
    self.t = t
    self.keep = keep

  @property
  def shape(self):
    return tuple(self.t.shape[i] for i in range(self.t.ndim))

  def byref(self):
    return C.byref(self.t)


def _from_parts(ptr, device_type, device_id, shape, np_dtype, keep):
  code, bits = _DTYPES[np.dtype(np_dtype)]
  shp = (C.c_int64 * len(shape))(*[int(s) for s in shape])
  t = Tensor(C.c_void_p(ptr), device_type, device_id, len(shape), code, bits, 1, shp, None, 0)
  return TensorRef(t, (keep, shp))


def _is_capsule(obj):
  return type(obj).__name__ == "PyCapsule"


def as_tensor(obj, device_id: int = 0) -> TensorRef | None:
  """numpy array -> host tensor; DeviceArray / __cuda_array_interface__ -> device tensor (zero-copy);
  __dlpack__ object or DLPack capsule -> whatever device it says (zero-copy)."""
  if obj is None:
    return None
  if isinstance(obj, TensorRef):
    return obj
  if isinstance(obj, np.ndarray):
    if not obj.flags.c_contiguous:
      raise ValueError("numpy inputs must be C-contiguous (NHWC dense)")
    if obj.dtype not in _DTYPES:
      raise TypeError(f"unsupported dtype {obj.dtype}")
    kind = _lib.DL_CUDA_HOST if obj.__array_interface__["data"][0] in _PINNED else _lib.DL_CPU
    return _from_parts(obj.ctypes.data, kind, 0, obj.shape, obj.dtype, obj)
  if hasattr(obj, "__cuda_array_interface__"):
    cai = obj.__cuda_array_interface__
    if cai.get("strides") is not None:
      st, acc = cai["strides"], np.dtype(cai["typestr"]).itemsize
      for n, s in zip(reversed(cai["shape"]), reversed(st)):
        if n != 1 and s != acc:
          raise ValueError("device inputs must be dense")
        acc *= n
    dev = getattr(getattr(obj, "ctx", None), "device", None)
    if dev is None:
      dev = getattr(getattr(obj, "device", None), "index", None)
    return _from_parts(cai["data"][0], _lib.DL_CUDA, device_id if dev is None else dev, cai["shape"], np.dtype(cai["typestr"]), obj)
  capsule = obj if _is_capsule(obj) else (obj.__dlpack__() if hasattr(obj, "__dlpack__") else None)
  if capsule is not None:
    if not _capsule_is_valid(capsule, b"dltensor"):
      raise ValueError("not a live 'dltensor' capsule (already consumed?)")
    p = _capsule_get_pointer(capsule, b"dltensor")
    mt = _DLManagedTensor.from_address(p)
    src = mt.dl_tensor
    t = Tensor()
    C.memmove(C.byref(t), C.byref(src), C.sizeof(Tensor))
    return TensorRef(t, (obj, capsule))
  raise TypeError(f"cannot interpret {type(obj)} as a tensor")


# --- DLPack export: hand a DeviceArray / DeviceView / numpy array to a consumer (tf.experimental.dlpack.from_dlpack, ...) zero-copy ---
_DL_DELETER = C.CFUNCTYPE(None, C.c_void_p)
_DL_CAPSULE_DTOR = C.CFUNCTYPE(None, C.c_void_p)
_DL_LIVE = {}     # address of the DLManagedTensor -> everything that must outlive the consumer's use of it



@_DL_DELETER
def _dl_deleter(managed_ptr):
  _DL_LIVE.pop(managed_ptr, None)     # drops the owner reference: the buffer may be freed from here on


@_DL_CAPSULE_DTOR
def _dl_capsule_destructor(capsule_ptr):
  # a capsule that was never consumed still owns the tensor (consumers rename theirs to "used_dltensor")
  if _capsule_get_name_raw(capsule_ptr) == b"dltensor":
    _DL_LIVE.pop(_capsule_get_pointer_raw(capsule_ptr, b"dltensor"), None)


def to_dlpack(obj):
  """'dltensor' capsule of a DeviceArray / DeviceView (kDLCUDA) or a C-contiguous numpy array (kDLCPU; page-locked ones as
  kDLCUDAHost).  The producer object stays alive until the consumer calls the deleter: zero-copy hand-off of decode outputs,
  e.g. ``tf.experimental.dlpack.from_dlpack(to_dlpack(out["image"]))``."""
  if isinstance(obj, np.ndarray):
    if not obj.flags.c_contiguous or obj.dtype not in _DTYPES:
      raise ValueError("numpy arrays must be C-contiguous float32 / uint8 / int16 / int8")
    ptr, shape, dtype = obj.ctypes.data, obj.shape, obj.dtype
    dev_type, dev_id = (_lib.DL_CUDA_HOST if ptr in _PINNED else _lib.DL_CPU), 0
  elif isinstance(obj, (DeviceArray, DeviceView)):
    ptr, shape, dtype, dev_type, dev_id = obj.ptr.value, obj.shape, obj.dtype, _lib.DL_CUDA, obj.ctx.device
  else:
    raise TypeError(f"cannot export {type(obj)} through DLPack")
  code, bits = _DTYPES[np.dtype(dtype)]
  shp = (C.c_int64 * max(len(shape), 1))(*[int(s) for s in shape])
  mt = _DLManagedTensor()
  mt.dl_tensor = Tensor(C.c_void_p(ptr), dev_type, dev_id, len(shape), code, bits, 1, shp, None, 0)
  mt.manager_ctx = None
  mt.deleter = C.cast(_dl_deleter, C.c_void_p)
  addr = C.addressof(mt)
  _DL_LIVE[addr] = (mt, shp, obj)
  return _capsule_new(addr, b"dltensor", C.cast(_dl_capsule_destructor, C.c_void_p))


def empty_like_kind(ctx: Context, like, shape, dtype):
  """Output buffer of the same kind as `like`: numpy for host inputs, DeviceArray for device inputs."""
  if isinstance(like, np.ndarray) or like is None:
    return np.empty(shape, dtype=dtype)
  return DeviceArray(ctx, shape, dtype)
