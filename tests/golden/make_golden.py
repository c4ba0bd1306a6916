"""Generates tests/golden/*.npz with torch-CPU, independently of oracle/ntc_oracle.py.

The reference (TF 2.10 + tensorflow-compression 2.10) cannot be imported here (not installed, Python
3.12, no network), so these fixtures do NOT pin the oracle to TensorFlow's bits; they pin it to the
*definitions* TensorFlow documents, written with a different library and a different formulation:

* Keras Conv2DTranspose(padding="SAME") = gradient w.r.t. the input of a SAME cross-correlation with
  the same kernel and stride (tf.nn.conv2d_transpose's definition), SAME pad_before = pad_total // 2;
  realised with torch autograd through F.conv2d.
* tfc SignalConv2D(corr=False, strides_up=s, "same_zeros") = insert s-1 zeros after every sample,
  then a true (flipped-kernel) convolution with zero padding that keeps the upsampled size;
  realised with F.conv2d on the zero-stuffed signal.
* GDN1 / classic GDN as in common/transforms.py:27-63 and tfc.GDN, with torch matmul.

Run:  python tests/golden/make_golden.py     (writes next to this file; a few hundred KB)
"""
import os
import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_default_dtype(torch.float64)


def keras_convt_same(x_nhwc, kernel_kkoi, bias, s):
  """gradient-of-SAME-conv definition (float64 autograd)."""
  x = torch.from_numpy(x_nhwc).permute(0, 3, 1, 2)                # [B, Cin_t, h, w]  (Cin of the transpose)
  k = kernel_kkoi.shape[0]
  w = torch.from_numpy(kernel_kkoi).permute(3, 2, 0, 1)           # forward conv: out-channels = Cin_t, in-channels = Cout_t
  B, _, h, wd = x.shape
  big = torch.zeros(B, kernel_kkoi.shape[2], h * s, wd * s, requires_grad=True)
  total = max(k - s, 0)
  pb, pa = total // 2, total - total // 2
  y = F.conv2d(F.pad(big, (pb, pa, pb, pa)), w, stride=s)         # SAME correlation, output h x w
  assert y.shape[-2:] == x.shape[-2:], (y.shape, x.shape)
  (y * x).sum().backward()
  out = big.grad.permute(0, 2, 3, 1).numpy()
  return out + (bias if bias is not None else 0.0)


def tfc_signal_conv_up(x_nhwc, kernel_kkio, bias, s):
  """zero-stuff by s (length N*s), then true convolution keeping the size (odd k, centred)."""
  x = torch.from_numpy(x_nhwc).permute(0, 3, 1, 2)
  B, C, h, w = x.shape
  up = torch.zeros(B, C, h * s, w * s)
  up[:, :, ::s, ::s] = x
  k = kernel_kkio.shape[0]
  wt = torch.from_numpy(kernel_kkio).permute(3, 2, 0, 1).flip(2, 3)   # convolution = correlation with flipped kernel
  out = F.conv2d(F.pad(up, (k // 2, (k - 1) // 2, k // 2, (k - 1) // 2)), wt)
  return out.permute(0, 2, 3, 1).numpy() + bias


def gdn1(x, beta, gamma, inverse):
  n = torch.from_numpy(np.abs(x)) @ torch.from_numpy(gamma) + torch.from_numpy(beta)
  return x * n.numpy() if inverse else x / n.numpy()


def gdn_classic(x, beta, gamma, inverse):
  n = torch.sqrt(torch.from_numpy(x * x) @ torch.from_numpy(gamma) + torch.from_numpy(beta)).numpy()
  return x * n if inverse else x / n


def main():
  rng = np.random.default_rng(42)
  out = {}
  # ---- single layers, every (k, s) on the decode path + awkward ones -------------------------------
  keras_cases = [(13, 8), (5, 2), (18, 16), (3, 1), (6, 4), (16, 16), (4, 2), (7, 3)]
  for k, s in keras_cases:
    x = rng.normal(size=(2, 3, 4, 5))
    w = rng.normal(size=(k, k, 4, 5)) * 0.2
    b = rng.normal(size=(4,))
    out[f"keras_k{k}s{s}_x"], out[f"keras_k{k}s{s}_w"], out[f"keras_k{k}s{s}_b"] = x, w, b
    out[f"keras_k{k}s{s}_y"] = keras_convt_same(x, w, b, s)
  for k, s in [(5, 2), (9, 4), (3, 1)]:
    x = rng.normal(size=(2, 4, 3, 6))
    w = rng.normal(size=(k, k, 6, 5)) * 0.2
    b = rng.normal(size=(5,))
    out[f"tfc_k{k}s{s}_x"], out[f"tfc_k{k}s{s}_w"], out[f"tfc_k{k}s{s}_b"] = x, w, b
    out[f"tfc_k{k}s{s}_y"] = tfc_signal_conv_up(x, w, b, s)
  # ---- tiny end-to-end transforms ------------------------------------------------------------------
  C, C1 = 8, 4
  y_hat = rng.normal(size=(1, 3, 4, C)) * 3
  wts = {
    "synthesis.base_conv.kernel": rng.normal(size=(13, 13, C1, C)) * 0.05, "synthesis.base_conv.bias": rng.normal(size=(C1,)) * 0.1,
    "synthesis.res.kernel": rng.normal(size=(13, 13, C1, C)) * 0.05, "synthesis.res.bias": rng.normal(size=(C1,)) * 0.1,
    "synthesis.activation.beta": 1 + rng.uniform(0, .5, size=(C1,)), "synthesis.activation.gamma": 0.1 * np.eye(C1) + rng.uniform(0, .05, size=(C1, C1)),
    "synthesis.out_conv.kernel": rng.normal(size=(5, 5, 3, C1)) * 0.2, "synthesis.out_conv.bias": rng.normal(size=(3,)) * 0.1,
  }
  base = keras_convt_same(y_hat, wts["synthesis.base_conv.kernel"], wts["synthesis.base_conv.bias"], 8)
  base = gdn1(base, wts["synthesis.activation.beta"], wts["synthesis.activation.gamma"], True)
  res = keras_convt_same(y_hat, wts["synthesis.res.kernel"], wts["synthesis.res.bias"], 8)
  out["tlr_yhat"] = y_hat
  for k_, v in wts.items():
    out["tlr_w:" + k_] = v
  out["tlr_out"] = keras_convt_same(base + res, wts["synthesis.out_conv.kernel"], wts["synthesis.out_conv.bias"], 2)

  # hyper-synthesis (k5s2, k5s2, k3s1 with relu) on a tiny grid
  Cz = 4
  hw = {
    "hyper_synthesis.layer_0.kernel": rng.normal(size=(5, 5, Cz, Cz)) * 0.2, "hyper_synthesis.layer_0.bias": rng.normal(size=(Cz,)) * 0.1,
    "hyper_synthesis.layer_1.kernel": rng.normal(size=(5, 5, 6, Cz)) * 0.2, "hyper_synthesis.layer_1.bias": rng.normal(size=(6,)) * 0.1,
    "hyper_synthesis.layer_2.kernel": rng.normal(size=(3, 3, 8, 6)) * 0.2, "hyper_synthesis.layer_2.bias": rng.normal(size=(8,)) * 0.1,
  }
  z = np.rint(rng.normal(size=(1, 2, 3, Cz)) * 1.5)
  h = np.maximum(keras_convt_same(z, hw["hyper_synthesis.layer_0.kernel"], hw["hyper_synthesis.layer_0.bias"], 2), 0)
  h = np.maximum(keras_convt_same(h, hw["hyper_synthesis.layer_1.kernel"], hw["hyper_synthesis.layer_1.bias"], 2), 0)
  out["hs_z"] = z
  for k_, v in hw.items():
    out["hs_w:" + k_] = v
  out["hs_out"] = keras_convt_same(h, hw["hyper_synthesis.layer_2.kernel"], hw["hyper_synthesis.layer_2.bias"], 1)

  # bls2017-style: 5x5 up2 + IGDN1, 5x5 up2 + IGDN1, 9x9 up4 ; mbt2018-style 2 layers with classic IGDN
  F_ = 6
  bw = {}
  cin = C
  for i, (k, co) in enumerate([(5, F_), (5, F_), (9, 3)]):
    bw[f"synthesis.layer_{i}.kernel"] = rng.normal(size=(k, k, cin, co)) * 0.1
    bw[f"synthesis.layer_{i}.bias"] = rng.normal(size=(co,)) * 0.1
    if i < 2:
      bw[f"synthesis.igdn_{i}.beta"] = 1 + rng.uniform(0, .5, size=(co,))
      bw[f"synthesis.igdn_{i}.gamma"] = 0.1 * np.eye(co) + rng.uniform(0, .05, size=(co, co))
    cin = co
  x = rng.normal(size=(1, 2, 3, C)) * 2
  out["bls_yhat"] = x
  for i, s in enumerate((2, 2, 4)):
    x = tfc_signal_conv_up(x, bw[f"synthesis.layer_{i}.kernel"], bw[f"synthesis.layer_{i}.bias"], s)
    if i < 2:
      x = gdn1(x, bw[f"synthesis.igdn_{i}.beta"], bw[f"synthesis.igdn_{i}.gamma"], True)
  out["bls_out"] = x
  x = out["bls_yhat"]
  for i in range(2):
    x = tfc_signal_conv_up(x, bw[f"synthesis.layer_{i}.kernel"], bw[f"synthesis.layer_{i}.bias"], 2)
    x = gdn_classic(x, bw[f"synthesis.igdn_{i}.beta"], bw[f"synthesis.igdn_{i}.gamma"], True)
  out["mbt2_out"] = x          # two SignalConv + classic IGDN stages (prefix of MBT2018Synthesis)
  for k_, v in bw.items():
    out["bls_w:" + k_] = v
  np.savez_compressed(os.path.join(HERE, "torch_definitions.npz"), **{k: np.asarray(v, dtype=np.float64) for k, v in out.items()})
  print("wrote", os.path.join(HERE, "torch_definitions.npz"), len(out), "arrays")


if __name__ == "__main__":
  main()
