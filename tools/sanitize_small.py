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
# round 2: device-resident small-batch decodes (CUDA-graph capture + replay, programmatic dependent launch), the band split,
# the other hyper-synthesis classes, LPIPS
from shallow_ntc_b200 import Model, lpips as L
m = build_config("two_layer_syn", precision="tc", ctx=ctx)
m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls="TwoLayerResSynthesis"))
zs, ys = m.latent_shapes(2, 97, 149)
z, q = synthetic.make_latents(zs, ys)
dz, dq = ctx.to_device(z), ctx.to_device(q.astype(np.int16))
out = dict(image=ctx.empty((2, 97, 149, 3), np.uint8), idx=ctx.empty(ys, np.uint8))
for _ in range(4):
  m.decompress(dz, dq, (97, 149), out=out)
print("graph replay ok", flush=True)
t = m.decompress_tiled(z, q, (97, 149), 2)
print("tiled ok", t["image"].shape, flush=True)
for hcfg in (dict(cls="JPEGLikeHyperSynthesis", bottleneck_size=320, kernel_size=6), dict(cls="HyperSynthesisSmall", bottleneck_size=320)):
  cfg = dict(analysis=dict(cls="ElicAnalysis", channels=(192, 192, 192, 320)), synthesis=dict(cls="JPEGLikeSynthesis", kernel_size=18, strides=16), hyper_synthesis=hcfg)
  mm = Model(cfg, precision="tc", ctx=ctx)
  mm.load_weights(synthetic.make_weights(mm.variable_shapes(), "stress", synthesis_cls="JPEGLikeSynthesis"))
  zs, ys = mm.latent_shapes(1, 70, 90)
  z, q = synthetic.make_latents(zs, ys)
  mm.decompress(z, q, (70, 90))
  print(hcfg["cls"], "ok", flush=True)
b = np.clip(a.astype(int) + 9, 0, 255).astype(np.uint8)
print("lpips", L.Lpips(ctx, L.random_weights())(np.ascontiguousarray(a[:, :64, :80]), np.ascontiguousarray(b[:, :64, :80])))
