"""GPU diagnostic: tensor-core path vs fp32 CUDA-core path vs float64 oracle, stage by stage."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from shallow_ntc_b200 import Context, synthetic
from helpers import make_case, oracle_decode
from oracle import ntc_oracle as O

ctx = Context(0)
name = sys.argv[1] if len(sys.argv) > 1 else "two_layer_syn"
B, H, W = (int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (2, 128, 192)))
kind = sys.argv[5] if len(sys.argv) > 5 else "stress"
m32, wts, z, q = make_case(name, B, H, W, kind, "fp32", ctx)
mtc, _, _, _ = make_case(name, B, H, W, kind, "tc", ctx)
ref = oracle_decode(m32, wts, z, q, H, W)

def stat(tag, a, b):
  d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
  print(f"  {tag:34s} max|d|={d.max():.3e} mean|d|={d.mean():.3e} ref max={np.abs(b).max():.3e}", flush=True)

if m32.hyperprior:
  hs_ref = np.concatenate([ref["mu"], ref["raw_sigma"]], -1)
  print("hyper_synthesis (transform call):")
  hs32 = m32.hyper_synthesis(z); stat("fp32 vs oracle", hs32, hs_ref)
  hstc = mtc.hyper_synthesis(z); stat("tc   vs oracle", hstc, hs_ref)
  bad = np.argwhere(np.abs(hstc - hs_ref) > 1e-3)
  if len(bad):
    print("   first bad entries (b,y,x,c):", bad[:10].tolist(), " count", len(bad), "of", hstc.size)
print("synthesis (transform call) on oracle y_hat:")
x_ref = O.synthesis(m32._transform_config["synthesis"]["cls"], wts, ref["y_hat"].astype(np.float64),
                    {k: v for k, v in m32._transform_config["synthesis"].items() if k != "cls"})
x32 = m32.synthesis(ref["y_hat"]); stat("fp32 vs oracle", x32, x_ref)
xtc = mtc.synthesis(ref["y_hat"]); stat("tc   vs oracle", xtc, x_ref)
print("fused decode:")
for tag, m in (("fp32", m32), ("tc", mtc)):
  got = m.decompress(z, q, (H, W), return_float=True, return_yhat=m.hyperprior)
  stat(tag + " recon", got["float"], ref["recon"])
  if m.hyperprior:
    stat(tag + " y_hat", got["y_hat"], ref["y_hat"])
    far = ref["idx_dist"] > 2e-4 * np.maximum(1, ref["i_c"])
    print(f"  {tag} idx mismatches outside margin: {(got['idx'][far] != ref['idx'][far]).sum()} (in margin {(~far).sum()}, total mism {(got['idx'] != ref['idx']).sum()})")
  print(f"  {tag} u8 max diff {np.abs(got['image'].astype(int) - ref['recon_u8'].astype(int)).max()}")
# timing
for tag, m in (("fp32", m32), ("tc", mtc)):
  dz = ctx.to_device(z) if z is not None else None
  dq = ctx.to_device(q)
  out = None
  for _ in range(3):
    out = m.decompress(dz, dq, (H, W), out=out, sync=False) if m.hyperprior else m.decompress(dq, (H, W), out=out, sync=False)
  ctx.sync()
  e0, e1 = ctx.event(), ctx.event()
  e0.record()
  for _ in range(10):
    out = m.decompress(dz, dq, (H, W), out=out, sync=False) if m.hyperprior else m.decompress(dq, (H, W), out=out, sync=False)
  e1.record(); ctx.sync()
  print(f"{tag}: {e0.elapsed_ms(e1) / 10:.3f} ms per decode of {B}x{H}x{W}")
