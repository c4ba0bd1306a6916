#!/usr/bin/env python
"""bench.py -- decoded Mpx/s of the shallow-ntc decode hot path on N B200s (one process per GPU).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--config two_layer_syn] [--batch 24]

Workload (BASELINE.json configs[1]): mshyper two_layer_syn decode of a batch of 24 synthetic
Kodak-shaped 768x512 images per GPU ("weak" scaling: every rank decodes its own 24-image shard of the
seeded image list, no data-path collective; NCCL only sums the PSNR at the end).  A step = one
sntc_decode of the batch.  `value` = un-padded pixels decoded by all ranks / max-over-ranks device time
with inputs resident in HBM; `e2e` = the same through the public API with pinned HOST buffers
(host->device of the symbols and device->host of image + index map inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

H, W = 512, 768
FLOP_PER_PX = {"two_layer_syn": 40940.0}   # 2*MAC of the convs, reference counting


def load_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    d = json.load(open(p))
    return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), tflops_burst=d["bf16_tflops"], source="measured")
  return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source="fallback")


def ncu_traffic(label, batch):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind `label`, from the committed
  `ncu --set full` capture (profiles/r01_ncu_traffic.json, taken at batch 24); None when no capture matches."""
  p = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
  try:
    d = json.load(open(p))
    e = d["kernels"].get(label)
    return float(e["dram_bytes_per_launch"]) if e and d.get("batch") == batch else None
  except Exception:
    return None


class ClockSampler:
  """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    self.gpu, self.rows, self.proc = gpu_index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self):
    if not self.proc:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      try:
        sm.append(float(r[1])); mx.append(float(r[2]))
        for n, v in zip(names, r[5:9]):
          if v.lower().startswith("active"):
            reasons.add(n)
      except Exception:
        pass
    busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
    return dict(sm_mhz=float(np.median(busy)) if busy else None, sm_max_mhz=max(mx) if mx else None,
                reasons=sorted(reasons), samples=len(sm))


def dist_init(n_gpus):
  """torch.distributed is plumbing only (barrier, max-over-ranks, final PSNR sum)."""
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if world == 1:
    return None, 0, 1, 0
  import torch
  import torch.distributed as dist
  torch.cuda.set_device(local)
  os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
  return dist, rank, world, local


def cpu_reference_run(config, n_images, steps, warmup):
  """Times the oracle's float32 GEMM-form restatement (tier T1) of the same decode on the host cores:
  the stand-in for the reference's TF-2.10 CPU decode, which cannot be installed here."""
  from shallow_ntc_b200 import build_config, synthetic
  from oracle import ntc_oracle as O
  model = build_config(config)
  cfg = model._transform_config["synthesis"]
  kw = {k: v for k, v in cfg.items() if k != "cls"}
  wts = synthetic.make_weights(model.variable_shapes(), "stress", synthesis_cls=cfg["cls"])
  zs, ys = model.latent_shapes(n_images, H, W)
  z, q = synthetic.make_latents(zs, ys)

  def one():
    return O.mshyper_decode(wts, cfg["cls"], z, q, H, W, kw, dtype=np.float32, gemm_form=True)
  # all host threads for BLAS, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
  from threadpoolctl import threadpool_limits
  with threadpool_limits(limits=os.cpu_count()):
    for _ in range(warmup):
      one()
    t0 = time.perf_counter()
    for _ in range(steps):
      one()
    dt = time.perf_counter() - t0
  return n_images * H * W * steps / dt / 1e6, dt / steps


def main():
  # Libraries (NCCL's version banner, ...) write to fd 1: keep the real stdout for the ONE JSON line, send the rest to stderr
  sys.stdout.flush()
  json_fd = os.dup(1)
  os.dup2(2, 1)
  json_out = os.fdopen(json_fd, "w")
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="native", choices=["native", "reference"])
  ap.add_argument("--config", default="two_layer_syn")
  ap.add_argument("--batch", type=int, default=24)
  ap.add_argument("--precision", default=os.environ.get("SNTC_PRECISION", "auto"))
  ap.add_argument("--rotate", type=int, default=4, help="distinct input batches cycled through (working set > L2)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-io-stage", action="store_true", help="skip the separately timed host range-decode sample")
  ap.add_argument("--e2e-depth", type=int, default=2, help="device buffer sets of the streaming pipeline (copy/compute overlap); measured 2 / 3 / 4: 8.52 / 8.14 / 8.07 Gpx/s")
  ap.add_argument("--height", type=int, default=512, help="image height (side runs of the other BASELINE configs; the headline is 512x768)")
  ap.add_argument("--width", type=int, default=768)
  args = ap.parse_args()
  global H, W
  H, W = args.height, args.width
  args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
  cores = os.cpu_count()

  if args.impl == "reference":
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
      return 0
    n_img = 2
    v, sec = cpu_reference_run(args.config, n_img, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = dict(impl="reference", metric="decoded Mpx/s", value=v, unit="Mpx/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"mshyper {args.config} decode (BASELINE configs[1]): 24 x 768x512 per GPU, random-init 'stress' weights",
                            images_per_step=n_img, note="CPU arm times a 2-image sample of the 24-image step"),
                cpu_baseline=dict(value=v, unit="Mpx/s", cores=cores, kind="port",
                                  sample=f"{n_img} of the {args.batch} images per step, oracle tier T1 (numpy float32 GEMM-form, BLAS threads = all cores); "
                                         "TF-2.10 itself is not installable offline"),
                e2e=dict(value=v, unit="Mpx/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), file=json_out, flush=True)
    return 0

  dist, rank, world, local = dist_init(args.gpus)
  from shallow_ntc_b200 import build_config, synthetic, Context
  ctx = Context(local)
  from shallow_ntc_b200 import parallel as _par
  numa = _par.bind_host_to_gpu(ctx) if world > 1 else dict(bound=False)   # pinned buffers + enqueue thread next to the GPU
  B = args.batch
  peaks = load_peaks()

  def make_model(precision):
    m = build_config(args.config, precision=precision, ctx=ctx, prior=True)
    cfg = m._transform_config["synthesis"]
    m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=cfg["cls"]))
    m._ensure_native()
    return m

  precision = args.precision
  if precision == "auto":
    try:
      model = make_model("tc")
      precision = "tc"
    except Exception as e:   # tensor-core kernels unavailable: the fp32 CUDA-core path is still a GPU path
      if rank == 0:
        print(f"[bench] tensor-core path unavailable ({e}); using fp32 CUDA-core kernels", file=sys.stderr)
      model = make_model("fp32")
      precision = "fp32"
  else:
    model = make_model(precision)

  zs, ys = model.latent_shapes(B, H, W)
  # this rank's shard of the seeded image list; `rotate` distinct batches so consecutive steps never reuse inputs
  sets = []
  for r in range(args.rotate):
    z, q = synthetic.make_latents(zs, ys, first_index=(rank * args.rotate + r) * B)
    sets.append((z, q))
  hyper = zs is not None                      # the factorized model (bls2017) has no z_hat and no scale indexes
  dev = [(ctx.to_device(z) if hyper else None, ctx.to_device(q)) for z, q in sets]
  out_dev = dict(image=ctx.empty((B, H, W, 3), np.uint8))
  if hyper:
    out_dev["idx"] = ctx.empty(ys, np.uint8)

  def step_dev(i):
    dz, dq = dev[i % args.rotate]
    model.decompress(dz, dq, (H, W), out=out_dev, sync=False)

  def barrier():
    ctx.sync()
    if dist is not None:
      dist.barrier()

  for i in range(args.warmup):
    step_dev(i)
  barrier()
  sampler = ClockSampler(local)
  sampler.start()
  l0 = ctx.launch_count
  e0, e1 = ctx.event(), ctx.event()
  barrier()
  e0.record()
  th0 = time.perf_counter()
  for i in range(args.steps):
    step_dev(i)
  host_ms = (time.perf_counter() - th0) * 1e3 / args.steps   # host time to enqueue one step (must stay below the device time)
  e1.record()
  ctx.sync()
  ms = e0.elapsed_ms(e1)
  barrier()
  launches = ctx.launch_count - l0
  clocks = sampler.stop()
  # per-layer breakdown (CUDA events around every layer) in a separate, untimed pass: the extra event records
  # would otherwise sit between the kernels of the timed region
  model.profile_layers(True)
  for i in range(min(args.steps, 50)):
    step_dev(i)
  ctx.sync()
  model.profile_layers(False)
  prof = model.layer_profile()

  # ---- e2e: pinned HOST buffers through the public streaming API (DecodePipeline): every step uploads its
  # symbols (H2D) and downloads image + index map (D2H) inside the timed region; copies overlap the decode of
  # the neighbouring steps on separate streams ----
  from shallow_ntc_b200 import DecodePipeline

  def run_e2e(q_dtype):
    pin = [(ctx.pinned_like(z) if hyper else None, ctx.pinned_like(q.astype(q_dtype))) for z, q in sets[:2]]
    outs = [dict(image=ctx.pinned_empty((B, H, W, 3), np.uint8), idx=ctx.pinned_empty(ys, np.uint8) if hyper else None) for _ in range(2)]
    pipe = DecodePipeline(model, B, (H, W), q_dtype=q_dtype, depth=args.e2e_depth)
    for i in range(3):
      pipe.submit(pin[i % 2][0], pin[i % 2][1], outs[i % 2]["image"], outs[i % 2]["idx"])
    pipe.drain()
    barrier()
    f0, f1 = ctx.event(), ctx.event()
    f0.record(pipe.s_in.handle)
    for i in range(args.steps):
      pipe.submit(pin[i % 2][0], pin[i % 2][1], outs[i % 2]["image"], outs[i % 2]["idx"])
    f1.record(pipe.s_out.handle)
    pipe.drain()
    return (f0.elapsed_ms(f1), int((pin[0][0].nbytes if hyper else 0) + pin[0][1].nbytes),
            int(outs[0]["image"].nbytes + (outs[0]["idx"].nbytes if hyper else 0)), outs)

  def link_probe(n=10):
    """Host<->device copy rate of THIS box with all ranks copying at once: the same pinned buffers as the e2e steps, H2D and
    D2H concurrently on two streams.  It is the roofline of the e2e number (PCIe / host memory, not the GPU)."""
    from shallow_ntc_b200.pipeline import _Stream
    import ctypes as C
    from shallow_ntc_b200._lib import lib, check
    z, q = sets[0]
    src = ctx.pinned_like(q)
    dsrc = ctx.empty(q.shape, q.dtype)
    dst = ctx.pinned_empty((B, H, W, 3), np.uint8)
    ddst = ctx.empty((B, H, W, 3), np.uint8)
    s_in, s_out = _Stream(ctx), _Stream(ctx)
    res = {}
    for mode in ("h2d", "d2h", "both"):
      barrier()
      a0, a1, b0, b1 = ctx.event(), ctx.event(), ctx.event(), ctx.event()
      a0.record(s_in.handle); b0.record(s_out.handle)
      for _ in range(n):
        if mode != "d2h":
          check(lib.sntc_memcpy_h2d(ctx.handle, dsrc.ptr, src.ctypes.data_as(C.c_void_p), src.nbytes, s_in.handle))
        if mode != "h2d":
          check(lib.sntc_memcpy_d2h(ctx.handle, dst.ctypes.data_as(C.c_void_p), ddst.ptr, dst.nbytes, s_out.handle))
      a1.record(s_in.handle); b1.record(s_out.handle)
      s_in.sync(); s_out.sync()
      if mode != "d2h":
        res[mode + "_up_gbs"] = n * src.nbytes / (a0.elapsed_ms(a1) * 1e-3) / 1e9
      if mode != "h2d":
        res[mode + "_down_gbs"] = n * dst.nbytes / (b0.elapsed_ms(b1) * 1e-3) / 1e9
    return res

  ms_e2e, h2d, d2h, out_hosts = run_e2e(np.float32)
  ms_e2e_i8, h2d_i8, _, _ = run_e2e(np.int8)
  ms_e2e_i16, h2d_i16, _, _ = run_e2e(np.int16)     # what codec.decompress hands over when a symbol exceeds int8
  out_host = out_hosts[(args.steps - 1) % 2]

  # final quality sum over ranks (the only collective; NCCL all-reduce of 3 doubles)
  orig = synthetic.make_original(out_host["image"][:2], first_index=rank * B)
  zq = sets[(args.steps - 1) % 2]
  met = model.decompress(zq[0][:2] if hyper else None, zq[1][:2], (H, W), original=orig, return_bits=hyper)
  # [sum psnr, sum mse, sum bits_y, sum bits_z, n_images]: the reference averages per-image metrics (mshyper/models.py:300-317)
  qsum = np.array([met["psnr"].sum(), met["mse"].sum(), met["bits_y"].sum() if hyper else 0.0, met["bits_z"].sum() if hyper else 0.0,
                   float(len(met["psnr"]))])
  # cost of asking for the rate term as well (bits_y in the hyper-head epilogue + bits_z kernel), device-resident
  n_rd = max(20, args.steps // 4)
  for i in range(3):   # untimed: the first call with the rate term sizes its scratch buffers (cudaMalloc)
    model.decompress(dev[i % args.rotate][0], dev[i % args.rotate][1], (H, W), out=out_dev, return_bits=hyper, sync=False)
  ctx.sync()
  g0, g1 = ctx.event(), ctx.event()
  g0.record()
  for i in range(n_rd):
    dz, dq = dev[i % args.rotate]
    model.decompress(dz, dq, (H, W), out=out_dev, return_bits=hyper, sync=False)
  g1.record()
  ctx.sync()
  ms_rd = g0.elapsed_ms(g1) / n_rd
  link = link_probe()
  link_keys = sorted(link)
  link_sum = np.array([link[k] for k in link_keys])
  if dist is not None:
    from shallow_ntc_b200 import parallel
    link_sum = parallel.reduce_metric_sums(dist, link_sum, device=f"cuda:{local}")   # aggregate GB/s over ranks
    ms, ms_e2e, ms_e2e_i8, ms_e2e_i16 = (float(v) for v in parallel.max_over_ranks(dist, [ms, ms_e2e, ms_e2e_i8, ms_e2e_i16], device=f"cuda:{local}"))
    qsum = parallel.reduce_metric_sums(dist, qsum, device=f"cuda:{local}")     # NCCL: the only collective

  if rank == 0:
    headline = args.config == "two_layer_syn" and (H, W) == (512, 768)
    workload = (f"mshyper two_layer_syn decode (BASELINE configs[1]): {B} x 768x512 per GPU, random-init 'stress' weights" if headline else
                f"{args.config} decode (side run, not the headline workload): {B} x {W}x{H} per GPU, random-init 'stress' weights")
    px_step = world * B * H * W
    value = px_step * args.steps / (ms * 1e-3) / 1e6
    e2e = px_step * args.steps / (ms_e2e * 1e-3) / 1e6
    # dominant kernel = the layer with the largest share of device time
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    roof = None
    # algorithmic HBM bytes per step: symbols in (f32) + image and index map out
    alg_bytes = sets[0][1].nbytes + (sets[0][0].nbytes if hyper else 0) + B * H * W * 3 + (int(np.prod(ys)) if hyper else 0)
    if dom[0] is not None and dom[1]["macs"] > 0:
      per_launch_ms = dom[1]["ms"] / dom[1]["n"]
      ach = 2.0 * dom[1]["macs"] / (per_launch_ms * 1e-3) / 1e12
      roof = dict(bound="tensor", kernel=dom[0], achieved=ach, peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"],
                  traffic=ncu_traffic(dom[0], B), peak_source=peaks["source"] + " bf16 dense sustained", share_of_step=per_launch_ms / (ms / args.steps),
                  ms_per_launch=per_launch_ms, algorithmic_flops_per_launch=2.0 * dom[1]["macs"],
                  hbm_view=dict(algorithmic_gbs=world * alg_bytes * args.steps / (ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"]))
    line = dict(metric="decoded Mpx/s", value=value, unit="Mpx/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32" if precision == "fp32" else "f16x3-split/f32-accum", data="synthetic",
                config=dict(workload=workload,
                            images_per_gpu=B, precision=precision, l2="inputs rotate over %d distinct batches; per-step working set > 126 MB L2" % args.rotate,
                            layers_ms={k: round(v["ms"] / max(v["n"], 1), 4) for k, v in prof.items()},
                            mean_psnr_db=float(qsum[0] / qsum[4]), mean_bpp_synthetic=float((qsum[2] + qsum[3]) / qsum[4] / (H * W)),
                            ms_per_step_with_rate_term=ms_rd, host_enqueue_ms_per_step=round(host_ms, 4)),
                clocks=clocks, gpu_launches=int(launches),
                e2e=dict(value=e2e, unit="Mpx/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e / args.steps,
                         api=f"DecodePipeline.submit (float32 symbols, pinned host buffers, depth {args.e2e_depth})", host_numa_binding=numa,
                         int8_symbols=dict(value=px_step * args.steps / (ms_e2e_i8 * 1e-3) / 1e6, h2d_bytes_per_step=h2d_i8),
                         int16_symbols=dict(value=px_step * args.steps / (ms_e2e_i16 * 1e-3) / 1e6, h2d_bytes_per_step=h2d_i16),
                         host_link=dict({k: round(float(v), 1) for k, v in zip(link_keys, link_sum)},
                                        note="aggregate pinned-copy GB/s over all ranks copying at once (q symbols up, image down; "
                                             "'both' = the two directions concurrently): the e2e roofline of this box",
                                        link_bound_mpx=float(px_step / max(world * h2d / (link_sum[link_keys.index("both_up_gbs")] * 1e9),
                                                                           world * d2h / (link_sum[link_keys.index("both_down_gbs")] * 1e9)) / 1e6))),
                roofline=roof)
    if world == 1 and hyper and not args.no_io_stage:
      # The I/O stage either side of the hot path (north_star: "range decoding ... stays in the host coder ... timed
      # separately"): container bytes -> symbols on ONE host thread, on a 2-image sample of the same workload, with the GPU
      # phases of the two-phase decode (hyper-synthesis -> idx, then dequantise + synthesis) timed beside it.
      try:
        from shallow_ntc_b200 import EntropyCoder, codec
        wts_io = synthetic.make_weights(model.variable_shapes(), "stress", synthesis_cls=model._transform_config["synthesis"]["cls"])
        coder = EntropyCoder(prior_weights=wts_io)
        n_io = min(B, 8)
        z_io, q_io = sets[0][0][:n_io], sets[0][1][:n_io]
        blob = codec.compress(model, coder, z_io, q_io, (H, W))
        tim1, timN = {}, {}
        nthr = min(n_io, cores or 1)
        codec.decompress(model, coder, blob)
        codec.decompress(model, coder, blob, timing=tim1)
        codec.decompress(model, coder, blob, timing=timN, threads=nthr)
        nsym = int(z_io.size + q_io.size)
        line["io_stage"] = dict(range_decode_mpx_s=n_io * H * W / timN["range_decode_s"] / 1e6, range_decode_msym_s=nsym / timN["range_decode_s"] / 1e6,
                                host_threads=nthr, one_thread_mpx_s=n_io * H * W / tim1["range_decode_s"] / 1e6,
                                gpu_two_phase_s=timN["gpu_s"], range_decode_s=timN["range_decode_s"],
                                bits_per_px=8.0 * len(blob) / (n_io * H * W),
                                sample=f"{n_io} images ({nsym} symbols, {len(blob)} container bytes), codec.decompress with timing, one image per "
                                       "host thread; not part of `value` / `e2e` (the hot path starts from decoded symbols)")
      except Exception as e:   # the I/O stage must never take the headline line down
        line["io_stage"] = dict(error=f"{type(e).__name__}: {e}")
    if world == 1 and not args.no_cpu_baseline:
      n_img = 2
      v, sec = cpu_reference_run(args.config, n_img, 3, 1)
      line["cpu_baseline"] = dict(value=v, unit="Mpx/s", cores=cores, kind="port",
                                  sample=f"{n_img} of the {B} images x 3 steps, oracle tier T1 (numpy float32 GEMM-form + col2im), {sec:.2f} s/step")
    print(json.dumps(line), file=json_out, flush=True)
  if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
  return 0


if __name__ == "__main__":
  sys.exit(main())
