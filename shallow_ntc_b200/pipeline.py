"""Streaming decode of HOST-resident symbol batches with copy/compute overlap.

The range decoder hands the integer symbols over in host memory and the application wants the pixels back
in host memory; PCIe (not the GPU) is then the bottleneck unless the three phases overlap.  ``DecodePipeline``
keeps ``depth`` sets of device buffers and runs, on three streams,

    copy-in stream :  H2D(z_i, q_i)                          (waits until decode i-depth released the set)
    compute stream :  sntc_decode(z_i, q_i) -> image_i, idx_i (waits for H2D i and for D2H i-depth)
    copy-out stream:  D2H(image_i, idx_i)                     (waits for decode i)

so that the upload of batch i+1 and the download of batch i-1 hide behind the decode of batch i.  Host
arrays should be page-locked (``Context.pinned_empty``) for the copies to be asynchronous.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from ._lib import lib, check
from .tensors import Context, DeviceArray, Event


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
  def __init__(self, model, batch: int, image_hw, q_dtype=np.float32, depth: int = 2, return_idx: bool = True):
    model._ensure_native()
    self.model, self.ctx = model, model.ctx
    self.H, self.W = int(image_hw[0]), int(image_hw[1])
    self.B, self.depth = int(batch), int(depth)
    zs, ys = model.latent_shapes(batch, self.H, self.W)
    self.return_idx = return_idx and zs is not None
    ctx = self.ctx
    self.slots = []
    for _ in range(depth):
      s = dict(q=DeviceArray(ctx, ys, q_dtype), image=DeviceArray(ctx, (batch, self.H, self.W, model._synthesis.out_channels), np.uint8),
               ev_in=Event(ctx), ev_done=Event(ctx), ev_out=Event(ctx), used=False)
      s["z"] = DeviceArray(ctx, zs, np.float32) if zs is not None else None
      s["idx"] = DeviceArray(ctx, ys, np.uint8) if self.return_idx else None
      self.slots.append(s)
    self.s_in, self.s_out = _Stream(ctx), _Stream(ctx)
    self.n = 0

  def _h2d(self, dst: DeviceArray, src: np.ndarray):
    assert src.flags.c_contiguous and src.nbytes == dst.nbytes and src.dtype == dst.dtype, (src.shape, src.dtype, dst.shape, dst.dtype)
    check(lib.sntc_memcpy_h2d(self.ctx.handle, dst.ptr, src.ctypes.data_as(C.c_void_p), dst.nbytes, self.s_in.handle))

  def _d2h(self, dst: np.ndarray, src: DeviceArray):
    assert dst.flags.c_contiguous and dst.nbytes == src.nbytes and dst.dtype == src.dtype
    check(lib.sntc_memcpy_d2h(self.ctx.handle, dst.ctypes.data_as(C.c_void_p), src.ptr, src.nbytes, self.s_out.handle))

  def submit(self, z_host, q_host, image_host, idx_host=None) -> int:
    """Enqueue one batch; returns a ticket for wait().  Nothing blocks the host."""
    s = self.slots[self.n % self.depth]
    if s["used"]:
      self.s_in.wait(s["ev_done"])                       # the decode that last read this input set
    if s["z"] is not None:
      self._h2d(s["z"], z_host)
    self._h2d(s["q"], q_host)
    s["ev_in"].record(self.s_in.handle)
    check(lib.sntc_stream_wait_event(self.ctx.handle, None, s["ev_in"].handle))
    if s["used"]:
      check(lib.sntc_stream_wait_event(self.ctx.handle, None, s["ev_out"].handle))   # the D2H that last read this output set
    out = dict(image=s["image"])
    if s["idx"] is not None:
      out["idx"] = s["idx"]
    self.model.decompress(s["z"], s["q"], (self.H, self.W), return_idx=self.return_idx, out=out, sync=False)
    s["ev_done"].record(None)
    self.s_out.wait(s["ev_done"])
    self._d2h(image_host, s["image"])
    if s["idx"] is not None and idx_host is not None:
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
