"""Generates tests/golden/msssim_torch.npz with torch-CPU, independently of oracle/ntc_oracle.py.

tf.image.ssim / tf.image.ssim_multiscale (TF 2.10, what mshyper/models.py:321-332 calls on the uint8 images) cannot be
run here (TensorFlow is not installable offline), so this fixture pins the oracle's restatement (assumption A9) to the
documented definition written with a different library and a different formulation: the FULL 2-D 11x11 window obtained
by a softmax over the 2-D log-weights (tf.image's _fspecial_gauss) applied as a grouped F.conv2d, F.avg_pool2d after
replicate-padding odd sizes at the end (SYMMETRIC pad of one), float64.

Run:  python tests/golden/make_golden_msssim.py
"""
import os
import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_default_dtype(torch.float64)
WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def window2d(size=11, sigma=1.5):
  c = torch.arange(size, dtype=torch.float64) - (size - 1) / 2.0
  g = -0.5 * c * c / (sigma * sigma)
  g2 = (g[None, :] + g[:, None]).reshape(1, -1)
  return torch.softmax(g2, dim=-1).reshape(size, size)


def ssim_per_channel(x, y, max_val=1.0):          # x, y: [B, C, H, W]
  C = x.shape[1]
  k = window2d()[None, None].repeat(C, 1, 1, 1)
  red = lambda t: F.conv2d(t, k, groups=C)
  c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
  m0, m1 = red(x), red(y)
  lum = (2 * m0 * m1 + c1) / (m0 * m0 + m1 * m1 + c1)
  cs = (2 * red(x * y) - 2 * m0 * m1 + c2) / (red(x * x + y * y) - m0 * m0 - m1 * m1 + c2)
  return (lum * cs).mean(dim=(2, 3)), cs.mean(dim=(2, 3))


def ms_ssim(a_u8, b_u8):
  x = torch.from_numpy(a_u8.astype(np.float64) / 255.0).permute(0, 3, 1, 2)
  y = torch.from_numpy(b_u8.astype(np.float64) / 255.0).permute(0, 3, 1, 2)
  if x.shape[2] < 160 and x.shape[3] < 160:
    return ssim_per_channel(x, y)[0].mean(dim=1).numpy()
  mcs = []
  for k in range(5):
    if k > 0:
      ph, pw = x.shape[2] % 2, x.shape[3] % 2
      x, y = (F.avg_pool2d(F.pad(t, (0, pw, 0, ph), mode="replicate"), 2) for t in (x, y))
    s, cs = ssim_per_channel(x, y)
    mcs.append(torch.relu(cs))
  fac = torch.stack(mcs[:-1] + [torch.relu(s)], dim=-1)
  return torch.prod(fac ** torch.tensor(WEIGHTS), dim=-1).mean(dim=1).numpy()


def make_images(seed, B, H, W):
  """Smooth structure + texture, and a distorted copy (blur-free additive noise of growing strength per image)."""
  rng = np.random.default_rng(seed)
  yy, xx = np.mgrid[0:H, 0:W]
  a = np.zeros((B, H, W, 3))
  for b in range(B):
    for c in range(3):
      f = rng.uniform(0.02, 0.2, size=4)
      a[b, :, :, c] = 127 + 60 * np.sin(f[0] * yy + f[1] * xx) + 40 * np.cos(f[2] * yy - f[3] * xx) + rng.normal(0, 12, (H, W))
  a = np.clip(np.rint(a), 0, 255).astype(np.uint8)
  noise = np.stack([rng.normal(0, 3.0 * (b + 1), a.shape[1:]) for b in range(B)])
  b_img = np.clip(np.rint(a + noise), 0, 255).astype(np.uint8)
  return a, b_img


if __name__ == "__main__":
  out = {}
  for name, (B, H, W) in dict(multi_odd=(2, 200, 181), small=(2, 96, 120), multi_even=(1, 192, 256)).items():
    a, b = make_images(hash(name) % 1000 if False else {"multi_odd": 11, "small": 12, "multi_even": 13}[name], B, H, W)
    out[name + "_a"] = a
    out[name + "_b"] = b
    out[name + "_val"] = ms_ssim(a, b)
    print(name, out[name + "_val"])
  np.savez_compressed(os.path.join(HERE, "msssim_torch.npz"), **out)
