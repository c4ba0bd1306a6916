"""CPU comparison line for tools/vjp_bench.py: torch float32 autograd (oneDNN) of the same two transforms on the host cores -- the
oracle's autograd statement (oracle/torch_vjp.py) timed as the closest stand-in for TensorFlow's GradientTape through the decoder.
Lives under tests/ because it executes oracle/ (checker code, never the product).  Not collected by pytest.
  python tests/vjp_cpu_timing.py [config] [batch] [H] [W]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def cpu(config, B, H, W, steps=2):
  import torch
  from oracle import ref_configs, torch_vjp as V
  from shallow_ntc_b200 import build_config, synthetic   # shapes / weights only (no GPU work)
  V.NP_DT = np.float32
  m = build_config(config)
  cfg = m._transform_config["synthesis"]
  wts = synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=cfg["cls"])
  zs, ys = m.latent_shapes(B, H, W)
  rng = np.random.default_rng(0)
  up = m._synthesis.upsample
  y = rng.standard_normal(ys) * 2
  gx = rng.standard_normal((B, ys[1] * up, ys[2] * up, 3))
  kw = {k: v for k, v in cfg.items() if k not in ("cls", "channels", "kernel_sizes")}
  t0 = None
  for i in range(steps + 1):
    if i == 1:
      t0 = time.perf_counter()
    V.transform_vjp(cfg["cls"], wts, y, gx, kw)
    if zs is not None:
      z = np.rint(rng.standard_normal(zs) * 1.5)
      gh = rng.standard_normal((B, ys[1], ys[2], 2 * ys[3]))
      V.transform_vjp("HyperSynthesis", wts, z, gh, {})
  ms = (time.perf_counter() - t0) / steps * 1e3
  return dict(impl="cpu torch autograd float32 (oneDNN)", cores=os.cpu_count(), threads=torch.get_num_threads(), config=config, batch=B, H=H, W=W,
              ms_per_step=round(ms, 1), mpx_s=round(B * H * W / ms / 1e3, 3))



if __name__ == "__main__":
  config = sys.argv[1] if len(sys.argv) > 1 else "two_layer_syn"
  B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
  H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
  W = int(sys.argv[4]) if len(sys.argv) > 4 else 768
  print(json.dumps(cpu(config, B, H, W)), flush=True)
