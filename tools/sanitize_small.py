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
# round 2, later: TwoLayerResSynthesis(res_type="d2s") and the decoder backward (forward + backward of both transforms, fp32 and tc)
cfg = dict(analysis=dict(cls="ElicAnalysis", channels=(192, 192, 192, 320)),
           synthesis=dict(cls="TwoLayerResSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5), activation_type="igdn", res_type="d2s"))
md = Model(cfg, precision="tc", ctx=ctx)
md.load_weights(synthetic.make_weights(md.variable_shapes(), "stress", synthesis_cls="TwoLayerResSynthesis"))
zs, ys = md.latent_shapes(1, 70, 90)
z, q = synthetic.make_latents(zs, ys)
md.decompress(z, q, (70, 90))
print("d2s ok", flush=True)
rng = np.random.default_rng(1)
for name, prec in (("two_layer_syn", "tc"), ("two_layer_syn", "fp32"), ("mbt2018", "tc"), ("bls2017", "tc")):
  mv = build_config(name, precision=prec, ctx=ctx, vjp=True)
  mv.load_weights(synthetic.make_weights(mv.variable_shapes(), "stress", synthesis_cls=mv._transform_config["synthesis"]["cls"]))
  zs, ys = mv.latent_shapes(1, 64, 80)
  up = mv._synthesis.upsample
  gy = mv.synthesis_vjp(rng.standard_normal(ys).astype(np.float32), rng.standard_normal((1, ys[1] * up, ys[2] * up, 3)).astype(np.float32))
  if zs is not None:
    mv.hyper_synthesis_vjp(rng.standard_normal(zs).astype(np.float32), rng.standard_normal((1, ys[1], ys[2], 2 * ys[3])).astype(np.float32))
  print("vjp", name, prec, "ok", float(np.abs(gy).max()), flush=True)
