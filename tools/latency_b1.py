"""Single-image decode latency (BASELINE configs[0] regime): device-resident, back to back, CUDA events.
  python tools/latency_b1.py            -> one line per (config, variant); variants run in subprocesses (env switches are read once)
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(config, batch):
  import numpy as np
  from shallow_ntc_b200 import build_config, synthetic, Context
  ctx = Context(0)
  m = build_config(config, precision="tc", ctx=ctx)
  m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=m._transform_config["synthesis"]["cls"]))
  zs, ys = m.latent_shapes(batch, 512, 768)
  z, q = synthetic.make_latents(zs, ys)
  dz, dq = ctx.to_device(z), ctx.to_device(q.astype(np.int16))
  out = dict(image=ctx.empty((batch, 512, 768, 3), np.uint8), idx=ctx.empty(ys, np.uint8))
  for _ in range(20):
    m.decompress(dz, dq, (512, 768), out=out, sync=False)
  ctx.sync()
  a, b = ctx.event(), ctx.event()
  n = 300
  import time
  t0 = time.perf_counter()
  a.record()
  for _ in range(n):
    m.decompress(dz, dq, (512, 768), out=out, sync=False)
  host = (time.perf_counter() - t0) / n * 1e3
  b.record(); ctx.sync()
  ms = a.elapsed_ms(b) / n
  m.profile_layers(True)
  for _ in range(50):
    m.decompress(dz, dq, (512, 768), out=out, sync=False)
  ctx.sync(); m.profile_layers(False)
  prof = {k: round(v["ms"] / v["n"] * 1e3, 1) for k, v in m.layer_profile().items()}
  print(json.dumps(dict(config=config, batch=batch, env={k: v for k, v in os.environ.items() if k.startswith("SNTC_")}, us_per_decode=round(ms * 1e3, 1),
                        host_enqueue_us=round(host * 1e3, 1), layers_us_eager=prof, layers_sum_us=round(sum(prof.values()), 1))))


if __name__ == "__main__":
  if len(sys.argv) > 1:
    one(sys.argv[1], int(sys.argv[2]))
  else:
    for config, batch in (("jpegl", 1), ("two_layer_syn", 1), ("two_layer_syn", 4)):
      for env in ({}, {"SNTC_GRAPH": "0"}, {"SNTC_GRAPH": "0", "SNTC_TC_PDL": "0"}, {"SNTC_TC_PDL": "0"}):
        subprocess.run([sys.executable, __file__, config, str(batch)], env=dict(os.environ, **env))
