"""Debug: print the per-item timeline (clock64) of the tensor-core layer kernels for one bench-shaped decode.
SNTC_TC_TRACE=1 python tools/tc_trace.py [config] [B]"""
import os, sys
os.environ["SNTC_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallow_ntc_b200 import build_config, synthetic, Context
cfgname = sys.argv[1] if len(sys.argv) > 1 else "two_layer_syn"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
H, W = 512, 768
ctx = Context(0)
m = build_config(cfgname, precision="tc", ctx=ctx)
cfg = m._transform_config["synthesis"]
m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=cfg["cls"]))
zs, ys = m.latent_shapes(B, H, W)
z, q = synthetic.make_latents(zs, ys)
dz, dq = ctx.to_device(z), ctx.to_device(q)
for i in range(2):
  print(f"---- decode {i}", file=sys.stderr, flush=True)
  m.decompress(dz, dq, (H, W))
