"""Intra-frame sharding: split ONE frame into latent-row bands that decode independently (BASELINE configs[4]).

Every transposed convolution of the path is ``out[o] += in[n] * W[a]`` with ``o = n*s + a - p``, ``a in [0, k)``
(``common/transforms.py``: Keras ``Conv2DTranspose(padding="SAME")`` has p = max(k-s,0)//2, tfc ``SignalConv2D(same_zeros)``
p = (k-1)//2; the pointwise stages -- relu, GDN, q + mu, exp / clamp -- do not mix rows).  Output row o therefore reads the input
rows ``ceil((o + p - k + 1) / s) .. floor((o + p) / s)``.  Walking the layer chain backwards from a band of image rows gives
the latent rows the band needs; decoding the sub-tensor made of exactly those rows (a *halo* of recomputed rows around the
band's own) yields, on the band's rows, bit-for-bit what the whole-frame decode yields there: rows outside the sub-tensor are
either outside the frame too (zero padding in both cases) or out of reach of the rows that are kept.  No communication: the
halo is recomputed from symbols every rank can be handed (they are what the range decoder produces).

Halos that come out of the formula (rows of the sub-tensor beyond the band's own, before / after):
  bls2017 (5x5 up2, 5x5 up2, 9x9 up4), bands on the y grid:            y rows 1 / 2
  two-layer synthesis (k13 s8, k5 s2) and jpegl (k18 s16):             y rows 1 / 1, and through the hyper-synthesis
  (k5 s2, k5 s2, k3 s1; bands on the z grid, 64 image rows per z row):  z rows 2 / 2
  mbt2018 (4 x 5x5 up2):                                               y rows 1 / 2 -> z rows 2 / 2
A z row is 64 image rows, so a hyperprior band carries 256 halo rows: banding pays for large frames (4K: 34 z rows), not
for Kodak-sized ones; the factorized bls2017 frame (no hyperprior, 135 y rows at 4K) carries 3 y rows = 48 image rows.
"""
from __future__ import annotations

from dataclasses import dataclass


def _ceil_div(a, b):
  return -((-a) // b)


def input_rows(chain, lo, hi):
  """Rows [lo_in, hi_in] (inclusive) of the INPUT of a conv chain [(k, s, p), ...] (forward order) that output rows [lo, hi]
  depend on -- before clamping to the tensor's extent."""
  for k, s, p in reversed(chain):
    lo, hi = _ceil_div(lo + p - k + 1, s), (hi + p) // s
  return lo, hi


@dataclass
class Band:
  index: int
  rows: tuple        # image rows [r0, r1) this band is responsible for (clipped to H)
  y_rows: tuple      # latent rows [y0, y1) of the sub-tensor handed to the decode (band + halo)
  z_rows: tuple      # hyper-latent rows [z0, z1) of the sub-tensor, or None for the factorized model
  y_core: tuple      # latent rows [c0, c1) owned by the band (for idx / y_hat outputs)
  sub_h: int         # image height to ask the sub-decode for
  keep: tuple        # rows [k0, k1) of the sub-decode's image that are the band's rows


def plan_bands(H, syn_chain, syn_up, hyper_chain, hyper_up, n_bands):
  """Split the frame's coarsest latent grid (z rows for hyperprior models, y rows otherwise) into `n_bands` contiguous bands and
  derive, for each, the sub-tensor rows to decode.  Bands of zero height (more bands than latent rows) are dropped."""
  d = syn_up * (hyper_up if hyper_chain else 1)
  Hp = _ceil_div(H, d) * d
  hy = Hp // syn_up
  n_coarse = Hp // d
  bands = []
  for i in range(n_bands):
    a, b = (i * n_coarse) // n_bands, ((i + 1) * n_coarse) // n_bands      # coarse rows [a, b)
    if b <= a:
      continue
    r0, r1 = a * d, min(b * d, H)
    if r1 <= r0:
      continue
    yc0, yc1 = a * (d // syn_up), b * (d // syn_up)
    ylo, yhi = input_rows(syn_chain, r0, r1 - 1)                            # y rows the band's pixels read
    ylo, yhi = max(min(ylo, yc0), 0), min(max(yhi, yc1 - 1), hy - 1)        # ... and the band's own rows (their idx is an output)
    if hyper_chain:
      zlo, zhi = input_rows(hyper_chain, ylo, yhi)                          # z rows behind mu / sigma of those y rows
      zlo, zhi = max(zlo, 0), min(zhi, n_coarse - 1)
      zlo, zhi = min(zlo, ylo // hyper_up), max(zhi, yhi // hyper_up)       # the y sub-tensor is hyper_up x the z sub-tensor
      y0, y1 = zlo * hyper_up, (zhi + 1) * hyper_up
      z_rows = (zlo, zhi + 1)
    else:
      y0, y1 = ylo, yhi + 1
      z_rows = None
    sub_h = min(H, y1 * syn_up) - y0 * syn_up
    bands.append(Band(i, (r0, r1), (y0, y1), z_rows, (yc0, yc1), sub_h, (r0 - y0 * syn_up, r1 - y0 * syn_up)))
  return bands
