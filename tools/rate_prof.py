import sys, time; sys.path.insert(0, '.')
import numpy as np
from shallow_ntc_b200 import build_config, synthetic, Context
ctx = Context(0)
m = build_config("two_layer_syn", precision="tc", ctx=ctx, prior=True)
m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls="TwoLayerResSynthesis"))
B,H,W = 24,512,768
zs, ys = m.latent_shapes(B,H,W)
z,q = synthetic.make_latents(zs, ys)
dz,dq = ctx.to_device(z), ctx.to_device(q)
out = dict(image=ctx.empty((B,H,W,3), np.uint8), idx=ctx.empty(ys, np.uint8))
for bits in (False, True):
  for i in range(5): m.decompress(dz,dq,(H,W),out=out,return_bits=bits,sync=False)
  ctx.sync()
  m.profile_layers(True)
  t0=time.perf_counter()
  for i in range(50): m.decompress(dz,dq,(H,W),out=out,return_bits=bits,sync=False)
  ctx.sync(); dt=(time.perf_counter()-t0)/50*1e3
  m.profile_layers(False)
  print("return_bits", bits, "wall ms/step", round(dt,4), {k: round(v["ms"]/v["n"],4) for k,v in m.layer_profile().items()})
