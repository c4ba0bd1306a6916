"""Bitstream either side of the GPU path: host range coder (libsntc, CPU) + wire format.

The reference never produces a bitstream (``compression=False`` on every entropy-model call); this module
defines the ``compress`` / ``decompress`` pair its models lack, in the manner of the tensorflow-compression example
models: per image one string for the hyper-latent z (per-channel tables of the NoisyDeepFactorized prior) and one for
the latent symbols q = round(y - mu) (NoisyNormal scale-table rows picked by idx).  Decoding is two-phase because the
rows are only known after hyper-synthesis (``Model.decode_hyper`` -> idx -> range-decode q -> ``Model.decode_latents``).

Container v2 (little endian): magic ``SNTC`` u8 version=2 | u8 index_rounding (0 rint, 1 trunc) | u8 precision class
(0 fp32, 1 tensor-core) | u16 libsntc version | u32 B H W hz wz Cz hy wy Cy | B x (u32 len_z, u32 len_y, u32 crc32(idx)) |
payloads.  The y string of an image is only decodable with EXACTLY the scale-table rows the encoder used: the rows come
from the GPU hyper-synthesis, and they may differ by one on elements near a rounding boundary between the fp32 and the
tensor-core kernels or between library versions.  The header therefore records what produced them, ``decompress`` refuses a
stream made with another index rule / precision class, and the per-image CRC of idx turns any remaining difference
into a loud error instead of garbage symbols.
"""
from __future__ import annotations

import ctypes as C
import struct
import time
import zlib
import numpy as np

from ._lib import lib, check

MAGIC = b"SNTC\x02"
_ROUNDING_CODE = {"rint": 0, "trunc": 1}
_PRIOR_ORDER = [("matrix", 0), ("bias", 0), ("factor", 0), ("matrix", 1), ("bias", 1), ("factor", 1),
                ("matrix", 2), ("bias", 2), ("factor", 2), ("matrix", 3), ("bias", 3)]


class EntropyCoder:
  """CDF tables (tfc-style: tail_mass, range_coder_precision, overflow symbol) + range coder, on the host."""

  def __init__(self, num_scales=64, scale_min=0.11, scale_max=256.0, tail_mass=2.0 ** -8, precision=12, prior_weights=None,
               prefix="prior"):
    self.handle = C.c_void_p()
    check(lib.sntc_coder_create(num_scales, scale_min, scale_max, tail_mass, precision, C.byref(self.handle)))
    self.precision = precision
    self.num_channels = 0
    if prior_weights is not None:
      self.set_prior(prior_weights, prefix)

  def set_prior(self, weights, prefix="prior"):
    """Raw tfc.DeepFactorized variables (matrix_i [Cz,fo,fi], bias_i [Cz,fo,1], factor_i [Cz,fo,1])."""
    Cz = np.asarray(weights[f"{prefix}.matrix_0"]).shape[0]
    packed = np.concatenate([np.asarray(weights[f"{prefix}.{kind}_{i}"], dtype=np.float32).reshape(Cz, -1) for kind, i in _PRIOR_ORDER], axis=1)
    assert packed.shape == (Cz, 43)
    packed = np.ascontiguousarray(packed)
    check(lib.sntc_coder_set_prior(self.handle, Cz, packed.ctypes.data_as(C.POINTER(C.c_float))))
    self.num_channels = Cz

  def table(self, kind, row):
    """(offset, cdf uint32[nsym + 2]) of one table; the last symbol is the overflow (escape) symbol."""
    off, n, p = C.c_int32(), C.c_int32(), C.POINTER(C.c_uint32)()
    check(lib.sntc_coder_table(self.handle, kind, row, C.byref(off), C.byref(n), C.byref(p)))
    return off.value, np.ctypeslib.as_array(p, shape=(n.value + 2,)).copy()

  def _encode(self, kind, symbols, rows):
    sym = np.ascontiguousarray(np.rint(symbols).astype(np.int32).ravel())
    r = None if rows is None else np.ascontiguousarray(rows, dtype=np.uint8).ravel()
    assert r is None or r.size == sym.size
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    check(lib.sntc_coder_encode(self.handle, kind, sym.ctypes.data_as(C.POINTER(C.c_int32)),
                                r.ctypes.data_as(C.POINTER(C.c_uint8)) if r is not None else None, sym.size, C.byref(out), C.byref(n)))
    try:
      return C.string_at(out, n.value)
    finally:
      lib.sntc_coder_free(out)

  def _decode(self, kind, data, rows, n):
    r = None if rows is None else np.ascontiguousarray(rows, dtype=np.uint8).ravel()
    sym = np.empty(n, dtype=np.int32)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data) if len(data) else None
    check(lib.sntc_coder_decode(self.handle, kind, buf, len(data), r.ctypes.data_as(C.POINTER(C.c_uint8)) if r is not None else None,
                                n, sym.ctypes.data_as(C.POINTER(C.c_int32))))
    return sym

  def encode_y(self, q, idx):
    return self._encode(0, q, idx)

  def decode_y(self, data, idx):
    idx = np.asarray(idx)
    return self._decode(0, data, idx, idx.size).reshape(idx.shape)

  def encode_z(self, z):
    assert self.num_channels and np.asarray(z).shape[-1] == self.num_channels
    return self._encode(1, z, None)

  def decode_z(self, data, shape):
    assert self.num_channels and shape[-1] == self.num_channels
    return self._decode(1, data, None, int(np.prod(shape))).reshape(shape)

  def __del__(self):
    try:
      if self.handle:
        lib.sntc_coder_destroy(self.handle)
    except Exception:
      pass


def _model_tag(model):
  """(index_rounding code, precision class) of the model whose hyper-synthesis picks the scale-table rows."""
  return _ROUNDING_CODE[model.index_rounding], 0 if model.precision == "fp32" else 1


def pack(strings, B, H, W, z_shape, y_shape, tag=(1, 1), crcs=None):
  head = MAGIC + struct.pack("<BBH", tag[0], tag[1], lib.sntc_version() & 0xFFFF)
  head += struct.pack("<9I", B, H, W, z_shape[1], z_shape[2], z_shape[3], y_shape[1], y_shape[2], y_shape[3])
  crcs = crcs if crcs is not None else [0] * len(strings)
  lens = b"".join(struct.pack("<3I", len(sz), len(sy), c) for (sz, sy), c in zip(strings, crcs))
  return head + lens + b"".join(sz + sy for sz, sy in strings)


def unpack(blob, with_meta=False):
  if blob[:4] != MAGIC[:4]:
    raise ValueError("not an SNTC container")
  if blob[:5] != MAGIC:
    raise ValueError(f"SNTC container version {blob[4]} is not supported (this library reads version {MAGIC[4]})")
  rounding, pclass, version = struct.unpack_from("<BBH", blob, 5)
  B, H, W, hz, wz, Cz, hy, wy, Cy = struct.unpack_from("<9I", blob, 9)
  pos = 9 + 36
  lens = [struct.unpack_from("<3I", blob, pos + 12 * b) for b in range(B)]
  pos += 12 * B
  if pos + sum(a + b for a, b, _ in lens) != len(blob):
    raise ValueError("SNTC container: payload length does not match the header (truncated or corrupt)")
  strings = []
  for lz, ly, _ in lens:
    strings.append((blob[pos:pos + lz], blob[pos + lz:pos + lz + ly]))
    pos += lz + ly
  out = (strings, (B, H, W), (B, hz, wz, Cz), (B, hy, wy, Cy))
  if with_meta:
    return out + (dict(index_rounding=rounding, precision_class=pclass, lib_version=version, idx_crc=[c for _, _, c in lens]),)
  return out


def compress(model, coder, z_hat, q_y, image_hw):
  """Symbols -> container bytes.  The scale-table rows come from the GPU (phase 1), exactly as the decoder will see them."""
  H, W = image_hw
  z_hat = np.ascontiguousarray(z_hat, dtype=np.float32)
  idx = model.decode_hyper(z_hat)
  strings = [(coder.encode_z(z_hat[b]), coder.encode_y(q_y[b], idx[b])) for b in range(z_hat.shape[0])]
  crcs = [zlib.crc32(np.ascontiguousarray(idx[b]).tobytes()) for b in range(z_hat.shape[0])]
  return pack(strings, z_hat.shape[0], H, W, z_hat.shape, q_y.shape, _model_tag(model), crcs)


def _map(fn, items, threads):
  """Images are independent strings: decode them on `threads` host threads (the coder calls release the GIL)."""
  if threads <= 1 or len(items) <= 1:
    return [fn(it) for it in items]
  from concurrent.futures import ThreadPoolExecutor
  with ThreadPoolExecutor(max_workers=min(threads, len(items))) as ex:
    return list(ex.map(fn, items))


def decompress(model, coder, blob, timing=None, threads=1, **kw):
  """Container bytes -> decompress() dict (``image`` uint8 [B,H,W,3], ...): range-decode z (host) -> hyper-synthesis
  (GPU) -> idx -> range-decode q (host) -> dequantise + synthesis (GPU).  ``timing``: dict that receives the seconds
  spent in the host coder (``range_decode_s``) and in the GPU calls (``gpu_s``), reported separately (north_star).
  ``threads``: host threads of the range decoder (one image per task; the tables are read-only)."""
  strings, (B, H, W), zs, ys, meta = unpack(blob, with_meta=True)
  tag = _model_tag(model)
  if (meta["index_rounding"], meta["precision_class"]) != tag:
    raise ValueError(f"SNTC container was written with index_rounding code {meta['index_rounding']} / precision class "
                     f"{meta['precision_class']}, this model has {tag[0]} / {tag[1]}: the scale-table rows would not match")
  t0 = time.perf_counter()
  z = np.stack(_map(lambda st: coder.decode_z(st[0], zs[1:]), strings, threads)).astype(np.float32)
  t1 = time.perf_counter()
  idx = model.decode_hyper(z)
  for b in range(B):
    if zlib.crc32(np.ascontiguousarray(idx[b]).tobytes()) != meta["idx_crc"][b]:
      raise ValueError(f"SNTC container: the scale-table rows of image {b} differ from the encoder's (stream written by libsntc "
                       f"version {meta['lib_version']}, this is {lib.sntc_version()}): the y string is not decodable with this build")
  t2 = time.perf_counter()
  q = np.stack(_map(lambda bs: coder.decode_y(bs[1][1], idx[bs[0]]), list(enumerate(strings)), threads))
  qmax = int(np.abs(q).max(initial=0))   # escape-coded symbols can exceed any fixed width: pick the narrowest exact type
  q = q.astype(np.int8) if qmax <= 127 else (q.astype(np.int16) if qmax <= 32767 else q.astype(np.float32))
  if qmax >= 1 << 24:
    raise ValueError("decoded symbol magnitude exceeds what float32 represents exactly (corrupt stream?)")
  t3 = time.perf_counter()
  out = model.decode_latents(q, (H, W), **kw)
  t4 = time.perf_counter()
  out["idx"], out["z_hat"], out["q_y"] = idx, z, q
  if timing is not None:
    timing.update(range_decode_s=(t1 - t0) + (t3 - t2), gpu_s=(t2 - t1) + (t4 - t3))
  return out
