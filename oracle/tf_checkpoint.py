"""Minimal reader of TensorFlow V2 checkpoints (tensor bundles) without TensorFlow.  TEST INFRASTRUCTURE ONLY.

Used to read the LPIPS weights the reference vendors (``lpips_tf2/models/{vgg,lin}/exported.*``) so that the oracle's LPIPS can
be pinned to the known answers the reference records for them (``lpips_tf2/test.py:17-19``).

Format (tensorflow/core/util/tensor_bundle): ``<prefix>.index`` is an uncompressed leveldb table (blocks of prefix-compressed
key / value entries + restart array, 5-byte trailer per block, 48-byte footer with the index block's handle) mapping a tensor
name to a ``BundleEntryProto`` {1: dtype, 2: shape {2: dim {1: size}}, 3: shard_id, 4: offset, 5: size}; the key "" holds the
``BundleHeaderProto`` {1: num_shards}.  ``<prefix>.data-SSSSS-of-NNNNN`` are the raw little-endian tensor bytes.
"""
from __future__ import annotations

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 4: np.uint8}


def _varint(buf, pos):
  out, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    out |= (b & 0x7F) << shift
    if not b & 0x80:
      return out, pos
    shift += 7


def _block_entries(buf, offset, size):
  """(key, value) pairs of one table block [offset, offset + size)."""
  blk = buf[offset:offset + size]
  n_restarts = int.from_bytes(blk[-4:], "little")
  end = len(blk) - 4 - 4 * n_restarts
  pos, key = 0, b""
  while pos < end:
    shared, pos = _varint(blk, pos)
    non_shared, pos = _varint(blk, pos)
    vlen, pos = _varint(blk, pos)
    key = key[:shared] + bytes(blk[pos:pos + non_shared])
    pos += non_shared
    yield key, bytes(blk[pos:pos + vlen])
    pos += vlen


def _proto_fields(buf):
  """Flat decode of one protobuf message: list of (field number, wire type, value)."""
  pos, out = 0, []
  while pos < len(buf):
    tag, pos = _varint(buf, pos)
    field, wt = tag >> 3, tag & 7
    if wt == 0:
      v, pos = _varint(buf, pos)
    elif wt == 2:
      n, pos = _varint(buf, pos)
      v = buf[pos:pos + n]
      pos += n
    elif wt == 5:
      v = int.from_bytes(buf[pos:pos + 4], "little")
      pos += 4
    elif wt == 1:
      v = int.from_bytes(buf[pos:pos + 8], "little")
      pos += 8
    else:
      raise ValueError(f"unsupported protobuf wire type {wt}")
    out.append((field, wt, v))
  return out


def read_index(prefix):
  """name -> dict(dtype, shape, shard_id, offset, size), and the number of shards."""
  buf = open(prefix + ".index", "rb").read()
  footer = buf[-48:]
  if int.from_bytes(footer[-8:], "little") != _MAGIC:
    raise ValueError("not a leveldb table (bad magic)")
  pos = 0
  _, pos = _varint(footer, pos)          # metaindex handle
  _, pos = _varint(footer, pos)
  ioff, pos = _varint(footer, pos)       # index block handle
  isize, pos = _varint(footer, pos)
  entries, shards = {}, 1
  for _, handle in _block_entries(buf, ioff, isize):
    boff, p = _varint(handle, 0)
    bsize, p = _varint(handle, p)
    if buf[boff + bsize] != 0:
      raise ValueError("compressed table blocks are not supported")
    for key, val in _block_entries(buf, boff, bsize):
      f = _proto_fields(val)
      if key == b"":
        shards = next((v for n, _, v in f if n == 1), 1)
        continue
      e = dict(dtype=1, shape=(), shard_id=0, offset=0, size=0)
      for n, wt, v in f:
        if n == 1:
          e["dtype"] = v
        elif n == 2:
          e["shape"] = tuple(next((vv for nn, _, vv in _proto_fields(d) if nn == 1), 0) for fn, _, d in _proto_fields(v) if fn == 2)
        elif n == 3:
          e["shard_id"] = v
        elif n == 4:
          e["offset"] = v
        elif n == 5:
          e["size"] = v
      entries[key.decode()] = e
  return entries, shards


def load_checkpoint(prefix, numeric_only=True):
  """name -> numpy array for every tensor of a known numeric dtype."""
  entries, shards = read_index(prefix)
  files = {}
  out = {}
  for name, e in entries.items():
    if e["dtype"] not in _DTYPES:
      if numeric_only:
        continue
      raise ValueError(f"{name}: dtype {e['dtype']} not supported")
    sid = e["shard_id"]
    if sid not in files:
      files[sid] = np.memmap(f"{prefix}.data-{sid:05d}-of-{shards:05d}", dtype=np.uint8, mode="r")
    raw = files[sid][e["offset"]:e["offset"] + e["size"]]
    out[name] = np.frombuffer(raw.tobytes(), dtype=_DTYPES[e["dtype"]]).reshape(e["shape"])
  return out
