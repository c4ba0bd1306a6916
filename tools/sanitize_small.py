"""Small decodes of every config for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallow_ntc_b200 import build_config, synthetic, Context
ctx = Context(0)
for name, B, H, W in (("two_layer_syn", 2, 97, 149), ("jpegl", 1, 100, 150), ("two_layer_syn2:24", 1, 64, 64), ("two_layer_syn2:48", 1, 64, 64),
                      ("mbt2018", 1, 64, 64), ("bls2017", 1, 48, 80)):
  m = build_config(name, precision="tc", ctx=ctx, prior=not name.startswith("bls"))
  m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=m._transform_config["synthesis"]["cls"]))
  zs, ys = m.latent_shapes(B, H, W)
  z, q = synthetic.make_latents(zs, ys)
  out = m.decompress(z, q, (H, W), return_bits=m.hyperprior)
  orig = synthetic.make_original(out["image"])
  m.decompress(z, q, (H, W), original=orig, return_float=True)
  print(name, "ok", out["image"].shape, flush=True)
a = np.random.default_rng(0).integers(0, 256, (1, 192, 181, 3), dtype=np.uint8)
print("msssim", ctx.msssim(a, a)[0])
