"""Decoder backward (sntc_synthesis_vjp + sntc_hyper_synthesis_vjp) timing on one GPU: device-resident tensors, CUDA events.
One itinf_train_step (mshyper/models.py:401-408) spends its heavy work in exactly these two calls (forward + backward of both
transforms).  The CPU comparison line (torch float32 autograd of the same two transforms on the host cores, the closest stand-in for
TF's tape) is produced by tests/vjp_cpu_timing.py -- it times the oracle, and only tests/ may touch oracle/.
  python tools/vjp_bench.py [config] [batch] [H] [W]   ->  JSON lines (fp32, tc)
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def gpu(config, B, H, W, precision, steps=30):
  from shallow_ntc_b200 import build_config, synthetic, Context
  ctx = Context(0)
  m = build_config(config, precision=precision, ctx=ctx, vjp=True)
  m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=m._transform_config["synthesis"]["cls"]))
  zs, ys = m.latent_shapes(B, H, W)
  rng = np.random.default_rng(0)
  up = m._synthesis.upsample
  y = ctx.to_device((rng.standard_normal(ys) * 2).astype(np.float32))
  gx = ctx.to_device(rng.standard_normal((B, ys[1] * up, ys[2] * up, 3)).astype(np.float32))
  gy = ctx.empty(ys, np.float32)
  args = [(m.synthesis_vjp, y, gx, gy)]
  if zs is not None:
    z = ctx.to_device(np.rint(rng.standard_normal(zs) * 1.5).astype(np.float32))
    gh = ctx.to_device(rng.standard_normal((B, ys[1], ys[2], 2 * ys[3])).astype(np.float32))
    gz = ctx.empty(zs, np.float32)
    args.append((m.hyper_synthesis_vjp, z, gh, gz))
  def step():
    for fn, x, g, gi in args:
      fn(x, g, grad_in=gi, sync=False)
  for _ in range(3):
    step()
  ctx.sync()
  a, b = ctx.event(), ctx.event()
  a.record()
  for _ in range(steps):
    step()
  b.record(); ctx.sync()
  ms = a.elapsed_ms(b) / steps
  m.profile_layers(True)
  for _ in range(5):
    step()
  ctx.sync(); m.profile_layers(False)
  prof = {k: round(v["ms"] / v["n"], 4) for k, v in m.layer_profile().items()}
  counts = ctx.launch_counts
  return dict(impl=precision, config=config, batch=B, H=H, W=W, ms_per_step=round(ms, 4), mpx_s=round(B * H * W / ms / 1e3, 1), layers_ms=prof,
              launches_by_family={k: int(v) for k, v in counts.items()})


if __name__ == "__main__":
  config = sys.argv[1] if len(sys.argv) > 1 else "two_layer_syn"
  B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
  H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
  W = int(sys.argv[4]) if len(sys.argv) > 4 else 768
  for p in ("fp32", "tc"):
    print(json.dumps(gpu(config, B, H, W, p)), flush=True)
