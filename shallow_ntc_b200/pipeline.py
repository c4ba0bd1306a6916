"""Streaming decode of HOST-resident symbol batches with copy/compute overlap.

The range decoder hands the integer symbols over in host memory and the application wants the pixels back
in host memory; PCIe (not the GPU) is then the bottleneck unless the three phases overlap.  ``DecodePipeline``
keeps ``depth`` sets of device buffers and runs, on three streams,

    copy-in stream :  H2D(z_i | q_i)                         (waits until decode i-depth released the set)
    compute stream :  sntc_decode(z_i, q_i) -> image_i, idx_i (waits for H2D i and for D2H i-depth)
    copy-out stream:  D2H(image_i [| idx_i])                  (waits for decode i)

so that the upload of batch i+1 and the download of batch i-1 hide behind the decode of batch i.

Defaults follow what a decoder actually holds at this point:
  * symbols travel as **int16** (``q_dtype``): they come out of an integer range decoder; int8 when the producer knows
    |q| <= 127 (``codec.decompress`` checks), float32 (the reference's tensor dtype) still accepted -- bit-identical results;
  * the scale-table rows ``idx`` are NOT downloaded (``return_idx=False``): whoever holds decoded symbols already has the
    rows (phase 1, ``Model.decode_hyper``, produced them for the range decoder).  They are still computed and stay on the
    device (``slot['idx']``); ``return_idx=True`` adds the download;
  * every direction is ONE copy: z and q share one device allocation and one page-locked host buffer per slot
    (``host_slot(i)`` returns numpy views the producer fills in place), image (and idx) likewise.  Arrays from elsewhere are
    accepted too (one copy per array).  ``write_combined=True`` makes the upload buffers write-combined (host writes only).
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from ._lib import lib, check
from .tensors import Context, DeviceArray, DeviceView, Event


def _align(n, a=256):
  return (int(n) + a - 1) // a * a


class _Stream:
  def __init__(self, ctx: Context):
    self.ctx = ctx
    h = C.c_void_p()
    check(lib.sntc_stream_create(ctx.handle, C.byref(h)))
    self.handle = h

  def wait(self, ev: Event):
    check(lib.sntc_stream_wait_event(self.ctx.handle, self.handle, ev.handle))

  def sync(self):
    check(lib.sntc_stream_sync(self.ctx.handle, self.handle))

  def __del__(self):
    try:
      if self.ctx.handle and self.handle:
        lib.sntc_stream_destroy(self.ctx.handle, self.handle)
    except Exception:
      pass


class DecodePipeline:
  def __init__(self, model, batch: int, image_hw, q_dtype=np.int16, depth: int = 2, return_idx: bool = False,
               host_slots: int | None = None, write_combined: bool = False):
    model._ensure_native()
    self.model, self.ctx = model, model.ctx
    self.H, self.W = int(image_hw[0]), int(image_hw[1])
    self.B, self.depth = int(batch), int(depth)
    self.q_dtype = np.dtype(q_dtype)
    zs, ys = model.latent_shapes(batch, self.H, self.W)
    self.zs, self.ys = zs, ys
    self.has_z = zs is not None
    self.return_idx = bool(return_idx) and self.has_z
    ctx = self.ctx
    Co = model._synthesis.out_channels
    self.img_shape = (batch, self.H, self.W, Co)
    # packed layouts: [z f32 | q] up, [image u8 | idx u8] down (idx only when it is downloaded)
    self.z_bytes = int(np.prod(zs)) * 4 if self.has_z else 0
    self.q_off = _align(self.z_bytes)
    self.q_bytes = int(np.prod(ys)) * self.q_dtype.itemsize
    self.in_bytes = self.q_off + self.q_bytes
    self.img_bytes = int(np.prod(self.img_shape))
    self.idx_off = _align(self.img_bytes)
    self.idx_bytes = int(np.prod(ys)) if self.has_z else 0
    self.out_bytes = self.idx_off + self.idx_bytes          # device side always holds idx (the decode writes it)
    self.down_bytes = self.out_bytes if self.return_idx else self.img_bytes
    self.slots = []
    for _ in range(depth):
      din, dout = DeviceArray(ctx, (self.in_bytes,), np.uint8), DeviceArray(ctx, (self.out_bytes,), np.uint8)
      s = dict(din=din, dout=dout, ev_in=Event(ctx), ev_done=Event(ctx), ev_out=Event(ctx), used=False,
               q=DeviceView(din, self.q_off, ys, self.q_dtype), image=DeviceView(dout, 0, self.img_shape, np.uint8),
               z=DeviceView(din, 0, zs, np.float32) if self.has_z else None,
               idx=DeviceView(dout, self.idx_off, ys, np.uint8) if self.has_z else None)
      self.slots.append(s)
    self.s_in, self.s_out = _Stream(ctx), _Stream(ctx)
    self.n = 0
    self._host = []
    self._host_by_addr = {}
    for _ in range(host_slots if host_slots is not None else depth):
      self._new_host_slot(write_combined)

  # -- packed page-locked host buffers ----------------------------------------------------------------
  def _new_host_slot(self, write_combined):
    up = self.ctx.pinned_empty((self.in_bytes,), np.uint8, write_combined=write_combined)
    down = self.ctx.pinned_empty((self.out_bytes,), np.uint8)
    h = dict(up=up, down=down,
             q=up[self.q_off:self.q_off + self.q_bytes].view(self.q_dtype).reshape(self.ys),
             z=up[:self.z_bytes].view(np.float32).reshape(self.zs) if self.has_z else None,
             image=down[:self.img_bytes].reshape(self.img_shape),
             idx=down[self.idx_off:self.idx_off + self.idx_bytes].reshape(self.ys) if self.has_z else None)
    self._host.append(h)
    self._host_by_addr[h["q"].ctypes.data] = h
    self._host_by_addr[h["image"].ctypes.data] = h
    return h

  def host_slot(self, i: int) -> dict:
    """Views ``z`` / ``q`` (to fill) and ``image`` / ``idx`` (to read) into the i-th packed page-locked buffer pair: passing
    them to ``submit`` makes each direction a single copy."""
    return self._host[i % len(self._host)]

  # -- copies ------------------------------------------------------------------------------------------
  def _h2d_raw(self, dptr, src: np.ndarray, nbytes):
    check(lib.sntc_memcpy_h2d(self.ctx.handle, dptr, src.ctypes.data_as(C.c_void_p), nbytes, self.s_in.handle))

  def _d2h_raw(self, dst: np.ndarray, dptr, nbytes):
    check(lib.sntc_memcpy_d2h(self.ctx.handle, dst.ctypes.data_as(C.c_void_p), dptr, nbytes, self.s_out.handle))

  def _h2d(self, dst, src: np.ndarray):
    assert src.flags.c_contiguous and src.nbytes == dst.nbytes and src.dtype == dst.dtype, (src.shape, src.dtype, dst.shape, dst.dtype)
    self._h2d_raw(dst.ptr, src, dst.nbytes)

  def _d2h(self, dst: np.ndarray, src):
    assert dst.flags.c_contiguous and dst.nbytes == src.nbytes and dst.dtype == src.dtype
    self._d2h_raw(dst, src.ptr, src.nbytes)

  def submit(self, z_host, q_host, image_host, idx_host=None) -> int:
    """Enqueue one batch; returns a ticket for wait().  Nothing blocks the host.  ``idx_host`` is filled only when the
    pipeline was built with ``return_idx=True``."""
    if self.has_z and z_host is None:
      raise ValueError("this model has a hyperprior: submit() needs z_host")
    s = self.slots[self.n % self.depth]
    if s["used"]:
      self.s_in.wait(s["ev_done"])                       # the decode that last read this input set
    hp = self._host_by_addr.get(q_host.ctypes.data)
    if hp is not None and (not self.has_z or (z_host is not None and z_host.ctypes.data == hp["z"].ctypes.data)):
      self._h2d_raw(s["din"].ptr, hp["up"], self.in_bytes)                      # z | q in one copy
    else:
      if self.has_z:
        self._h2d(s["z"], z_host)
      self._h2d(s["q"], q_host)
    s["ev_in"].record(self.s_in.handle)
    check(lib.sntc_stream_wait_event(self.ctx.handle, None, s["ev_in"].handle))
    if s["used"]:
      check(lib.sntc_stream_wait_event(self.ctx.handle, None, s["ev_out"].handle))   # the D2H that last read this output set
    out = dict(image=s["image"])
    if s["idx"] is not None:
      out["idx"] = s["idx"]
    self.model.decompress(s["z"], s["q"], (self.H, self.W), return_idx=s["idx"] is not None, out=out, sync=False)
    s["ev_done"].record(None)
    self.s_out.wait(s["ev_done"])
    want_idx = self.return_idx and idx_host is not None
    hq = self._host_by_addr.get(image_host.ctypes.data)
    if hq is not None and (not want_idx or idx_host.ctypes.data == hq["idx"].ctypes.data):
      self._d2h_raw(hq["down"], s["dout"].ptr, self.down_bytes if want_idx else self.img_bytes)   # image [| idx] in one copy
    else:
      self._d2h(image_host, s["image"])
      if want_idx:
        self._d2h(idx_host, s["idx"])
    s["ev_out"].record(self.s_out.handle)
    s["used"] = True
    self.n += 1
    return self.n - 1

  def wait(self, ticket: int):
    """Block until the outputs of `ticket` are in host memory (valid while fewer than `depth` newer submits)."""
    ev = self.slots[ticket % self.depth]["ev_out"]
    check(lib.sntc_event_elapsed_ms(self.ctx.handle, ev.handle, ev.handle, C.byref(C.c_float())))

  def drain(self):
    self.s_in.sync()
    self.ctx.sync()
    self.s_out.sync()
