"""GPU edge cases beyond the pytest shapes: smallest images, a large single image, prime-ish sizes; tensor-core path against the
fp32 CUDA-core path (both already pinned to the oracle on the pytest shapes).  python tools/edge_check.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallow_ntc_b200 import build_config, synthetic, Context
ctx = Context(0)
worst = 0.0
for name in ("two_layer_syn", "jpegl", "two_layer_syn2:24", "two_layer_syn2:48"):
  ms = {}
  for prec in ("tc", "fp32"):
    m = build_config(name, precision=prec, ctx=ctx)
    m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=m._transform_config["synthesis"]["cls"]))
    ms[prec] = m
  for (B, H, W) in ((1, 16, 16), (1, 1, 1), (3, 63, 65), (1, 64, 64), (2, 257, 511), (1, 2048, 2048), (1, 1999, 1201)):
    zs, ys = ms["tc"].latent_shapes(B, H, W)
    z, q = synthetic.make_latents(zs, ys)
    a = ms["tc"].decompress(z, q, (H, W), return_float=True)
    b = ms["fp32"].decompress(z, q, (H, W), return_float=True)
    fast = ms["tc"].decompress(z, q, (H, W))
    err = float(np.abs(a["float"] - b["float"]).max())
    du8 = int(np.abs(a["image"].astype(int) - b["image"].astype(int)).max())
    didx = float((a["idx"] != b["idx"]).mean())
    same = bool(np.array_equal(fast["image"], a["image"]))
    worst = max(worst, err)
    print(f"{name:18s} B={B} {H}x{W}: tc-fp32 max|df|={err:.2e} max|du8|={du8} idx diff frac={didx:.2e} u8-only path same bytes={same}", flush=True)
    assert err < 1e-4 and du8 <= 1 and didx < 2e-2 and same
print("edge cases ok, worst float diff", worst)
