#!/usr/bin/env python
"""bench.py -- decoded Mpx/s of the shallow-ntc decode hot path on N B200s (one process per GPU).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--config two_layer_syn] [--batch 24]

Workload (BASELINE.json configs[1]): mshyper two_layer_syn decode of a batch of 24 synthetic
Kodak-shaped 768x512 images per GPU ("weak" scaling: every rank decodes its own 24-image shard of the
seeded image list, no data-path collective; NCCL only sums the metrics at the end).  A step = one
sntc_decode of the batch.  `value` = un-padded pixels decoded by all ranks / max-over-ranks device time
with inputs resident in HBM; `e2e` = the same through the public streaming API with page-locked HOST
buffers (host->device of the symbols and device->host of the image inside the timed region).
Prints ONE JSON line on rank 0.

`--impl reference` times the CPU statements of the same decode (oracle/: numpy GEMM-form and torch-CPU oneDNN) on the
host cores, without loading the product library.
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
MMA_PASSES = 3   # split-fp16 product: a_lo*w_hi + a_hi*w_lo + a_hi*w_hi per algorithmic MAC (the dominant kernel runs all three)


def load_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    d = json.load(open(p))
    return dict(hbm_gbs=d["hbm_gbs"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), tflops_burst=d["bf16_tflops"],
                sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured (MEASURED_PEAKS.json)")
  return dict(hbm_gbs=6650.0, tflops_sustained=1400.0, tflops_burst=1590.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(label, batch):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind `label`, from the committed
  `ncu --set full` captures (profiles/r0N_ncu_traffic.json, taken at batch 24); None when no capture matches."""
  for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
    try:
      d = json.load(open(os.path.join(ROOT, "profiles", name)))
      e = d["kernels"].get(label)
      if e and d.get("batch") == batch:
        return float(e["dram_bytes_per_launch"])
    except Exception:
      pass
  return None


class ClockSampler:
  """SM clock + clock-event reasons of ONE GPU, sampled for the whole run; windows are cut out afterwards by host time.
  NVML (nvidia_ml_py) at ~2 ms when importable -- the timed region of a 20-step run lasts 17 ms, far below what a polling
  `nvidia-smi -lms` can resolve -- else the recipe's nvidia-smi query at its fastest period.  start() returns only after the
  first sample has arrived."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
  BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

  def __init__(self, gpu_index, pci_bus_id=None):
    self.gpu, self.bus = gpu_index, pci_bus_id
    self.rows = []      # (t, sm_mhz, set(reasons), power_w)
    self.max_mhz = None
    self.source = None
    self._stop = False
    self.proc = None

  def start(self):
    try:
      import pynvml
      pynvml.nvmlInit()
      h = pynvml.nvmlDeviceGetHandleByPciBusId(self.bus.encode()) if self.bus else pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
      self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
      get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

      def loop():
        while not self._stop:
          try:
            sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            mask = int(get_reasons(h))
            try:
              pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
              pw = None
            self.rows.append((time.perf_counter(), sm, {n for b, n in self.BITS.items() if mask & b}, pw))
          except Exception:
            pass
          time.sleep(0.002)
      self.source = "nvml"
      self.t = threading.Thread(target=loop, daemon=True)
      self.t.start()
    except Exception:
      self._start_smi()
    t0 = time.perf_counter()
    while not self.rows and time.perf_counter() - t0 < 5.0:
      time.sleep(0.005)
    return self

  def _start_smi(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.source = "nvidia-smi -lms 20"
      names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

      def read():
        for line in self.proc.stdout:
          r = [c.strip() for c in line.split(",")]
          try:
            self.max_mhz = float(r[2])
            self.rows.append((time.perf_counter(), float(r[1]), {n for n, v in zip(names, r[5:9]) if v.lower().startswith("active")},
                              float(r[3]) if r[3].replace(".", "").isdigit() else None))
          except Exception:
            pass
      self.t = threading.Thread(target=read, daemon=True)
      self.t.start()
    except Exception:
      self.proc, self.source = None, "unavailable"

  def window(self, t0, t1, pad=0.004):
    """Summary of the samples taken in [t0, t1] (host perf_counter); when the window is shorter than the sampling period the
    nearest samples either side are used and the record says so."""
    rows = [r for r in self.rows if t0 - pad <= r[0] <= t1 + pad]
    note = None
    if not rows and self.rows:
      mid = 0.5 * (t0 + t1)
      rows = sorted(self.rows, key=lambda r: abs(r[0] - mid))[:2]
      note = "window shorter than the sampling period: nearest samples used"
    if not rows:
      return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=["no clock samples"], samples=0, source=self.source)
    sm = [r[1] for r in rows]
    reasons = sorted(set().union(*[r[2] for r in rows]))
    pw = [r[3] for r in rows if r[3] is not None]
    out = dict(sm_mhz=float(np.median(sm)), sm_min_mhz=float(min(sm)), sm_max_mhz=self.max_mhz, reasons=reasons, samples=len(rows),
               power_w_max=max(pw) if pw else None, source=self.source)
    if out["sm_mhz"] < 0.9 * (self.max_mhz or out["sm_mhz"]) and not reasons:
      # this part holds tensor-heavy kernels below the maximum clock before NVML raises sw_power_cap (its power figure is a ~1 s
      # average); the driver's own cuBLAS run shows the same (MEASURED_PEAKS.json: clocks_under_load 1305 MHz, throttled false)
      note = (note + "; " if note else "") + "below max clock with no reason flagged yet: power management ahead of the averaged sw_power_cap flag, not a clock lock"
    if note:
      out["note"] = note
    return out

  def stop(self):
    self._stop = True
    if self.proc:
      self.proc.terminate()
      try:
        self.proc.wait(timeout=2)
      except Exception:
        self.proc.kill()


def regime_peak(peaks, clocks):
  """The cuBLAS bf16 figure that matches how the kernel was timed: the burst one when the sampled SM clock sat at (>= 97 % of)
  the maximum with no power cap active during the window, the sustained one otherwise."""
  sm, mx = clocks.get("sm_mhz"), clocks.get("sm_max_mhz") or peaks["sm_max_mhz"]
  burst = sm is not None and sm >= 0.97 * mx and "sw_power_cap" not in clocks.get("reasons", [])
  return ("burst", peaks["tflops_burst"]) if burst else ("sustained", peaks["tflops_sustained"])


def dist_init(n_gpus):
  """Process-group plumbing (barrier, max-over-ranks, final metric sum): NCCL through libsntc itself
  (shallow_ntc_b200.parallel.NcclGroup: ncclCommInitRank on an id exchanged through a file rendezvous), no PyTorch."""
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  return rank, world, local


# ----------------------------------------------------------------------------------------------------------------------
# CPU arms (oracle/: test infrastructure; never the product path)

def cpu_arm(config, batch, steps, warmup, budget_s=150.0, which=("numpy", "torch")):
  """Times the CPU statements of the decode on all host cores: oracle tier T1 (numpy float32 GEMM-form + col2im) and the
  torch-CPU oneDNN pipeline (the closest stand-in for TF-2.10's oneDNN CPU kernels; TF itself cannot be installed here,
  SURVEY F10).  A step is the full `batch`-image step when K + W of them fit the time budget, otherwise a bounded sample of it
  (stated).  Inputs come from oracle/ref_configs.py: the product library is not loaded."""
  from oracle import ref_configs as R
  from oracle import ntc_oracle as O
  cfg, wts, z, q = R.make_case(config, batch, H, W, "stress")
  syn = cfg["synthesis"]
  kw = {k: v for k, v in syn.items() if k != "cls"}
  cores = os.cpu_count()
  res = {}

  def numpy_step(n):
    if cfg["hyperprior"]:
      return O.mshyper_decode(wts, syn["cls"], z[:n], q[:n], H, W, kw, dtype=np.float32, gemm_form=True)
    return O.factorized_decode(wts, syn["cls"], q[:n], H, W, kw, dtype=np.float32, gemm_form=True)

  def run(step_fn, label):
    t0 = time.perf_counter()
    step_fn(1)
    t_img = max(time.perf_counter() - t0, 1e-4)          # also the first warm-up
    n = int(max(1, min(batch, budget_s / ((steps + warmup) * t_img))))
    for _ in range(max(warmup - 1, 0) if n == 1 else warmup):
      step_fn(n)
    t0 = time.perf_counter()
    for _ in range(steps):
      step_fn(n)
    dt = (time.perf_counter() - t0) / steps
    res[label] = dict(value=n * H * W / dt / 1e6, s_per_step=dt, images_per_step=n, same_config=n == batch)

  if "numpy" in which:
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=cores):        # all host threads for BLAS, whatever OMP_NUM_THREADS the launcher exported
      run(numpy_step, "numpy_T1")
  if "torch" in which:
    try:
      import torch
      torch.set_num_threads(cores)
      from oracle import torch_ref as T
      dec = T.TorchDecoder(cfg, wts)
      run(lambda n: dec(z[:n] if z is not None else None, q[:n], H, W), "torch_onednn")
    except Exception as e:
      res["torch_onednn"] = dict(error=f"{type(e).__name__}: {e}")
  return res, cores


def parity_sample(config, gpu_out, n=1):
  """The checker leg: oracle tier T0 (float64) on the first `n` images of rank 0's first batch against what the GPU returned
  for them.  Reported, not timed; needs the oracle, hence part of the cpu_baseline leg."""
  from oracle import ref_configs as R
  from oracle import ntc_oracle as O
  cfg, wts, z, q = R.make_case(config, n, H, W, "stress")
  syn = cfg["synthesis"]
  kw = {k: v for k, v in syn.items() if k != "cls"}
  if not cfg["hyperprior"]:
    ref = O.factorized_decode(wts, syn["cls"], q, H, W, kw)
  else:
    ref = O.mshyper_decode(wts, syn["cls"], z, q, H, W, kw, index_rounding=gpu_out["index_rounding"])
  rep = dict(images=n, oracle="T0 float64 scatter definition")
  d = np.abs(gpu_out["image"][:n].astype(np.int16) - ref["recon_u8"].astype(np.int16))
  rep.update(u8_max_diff=int(d.max()), u8_frac_diff=float((d > 0).mean()))
  if "float" in gpu_out:
    rep["recon_max_abs"] = float(np.abs(gpu_out["float"][:n].astype(np.float64) - ref["recon"]).max())
  if cfg["hyperprior"]:
    margin = 2e-4 * np.maximum(1.0, ref["i_c"])          # tests/helpers.py IDX_MARGIN["tc"]
    far = ref["idx_dist"] > margin
    gi = gpu_out["idx"][:n]
    rep.update(index_rounding=gpu_out["index_rounding"], idx_elements=int(far.size), idx_in_margin=int((~far).sum()),
               idx_mismatch_outside_margin=int((gi[far] != ref["idx"][far]).sum()), idx_mismatch_in_margin=int((gi[~far] != ref["idx"][~far]).sum()),
               idx_margin="|i_c - boundary| <= 2e-4 * max(1, i_c)")
  return rep


def reference_main(args, json_out):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return 0
  res, cores = cpu_arm(args.config, args.batch, max(1, args.steps), max(1, args.warmup))
  ok = {k: v for k, v in res.items() if "value" in v}
  best = max(ok, key=lambda k: ok[k]["value"])
  b = ok[best]
  sample = (f"{b['images_per_step']} of the {args.batch} images per step x {args.steps} steps, {best} on {cores} host threads; all CPU statements timed: " +
            ", ".join(f"{k} {v['value']:.2f} Mpx/s ({v['images_per_step']} img/step)" if "value" in v else f"{k} failed" for k, v in res.items()) +
            "; TF-2.10 itself is not installable offline")
  line = dict(impl="reference", metric="decoded Mpx/s", value=b["value"], unit="Mpx/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
              ms_per_step=b["s_per_step"] * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
              config=dict(workload=f"mshyper {args.config} decode (BASELINE configs[1]): {args.batch} x {W}x{H} per GPU, random-init 'stress' weights",
                          images_per_step=b["images_per_step"], same_config=b["same_config"], cpu_statement=best, all=res),
              cpu_baseline=dict(value=b["value"], unit="Mpx/s", cores=cores, kind="port", sample=sample),
              e2e=dict(value=b["value"], unit="Mpx/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
  print(json.dumps(line), file=json_out, flush=True)
  return 0


# ----------------------------------------------------------------------------------------------------------------------
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
  ap.add_argument("--no-side", action="store_true", help="skip the side records (single-image jpegl latency, sustained loop)")
  ap.add_argument("--e2e-depth", type=int, default=2, help="device buffer sets of the streaming pipeline (copy/compute overlap); measured 2 / 3 / 4: 8.52 / 8.14 / 8.07 Gpx/s")
  ap.add_argument("--height", type=int, default=512, help="image height (side runs of the other BASELINE configs; the headline is 512x768)")
  ap.add_argument("--width", type=int, default=768)
  ap.add_argument("--tile-frames", action="store_true", help="intra-frame sharding (BASELINE configs[4]): every frame is split into WORLD_SIZE latent-row bands "
                                                             "(halo recomputed, shallow_ntc_b200/tiling.py) and rank r decodes band r of the same `batch` frames: strong scaling of a frame; "
                                                             "default = whole frames per rank (weak scaling)")
  args = ap.parse_args()
  global H, W
  H, W = args.height, args.width
  args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

  if args.impl == "reference":
    return reference_main(args, json_out)

  rank, world, local = dist_init(args.gpus)
  from shallow_ntc_b200 import build_config, synthetic, Context
  from shallow_ntc_b200 import parallel as par
  ctx = Context(local)
  group = par.NcclGroup.from_env(ctx) if world > 1 else None
  numa = par.bind_host_to_gpu(ctx) if world > 1 else dict(bound=False)   # pinned buffers + enqueue thread next to the GPU
  B = args.batch
  peaks = load_peaks()
  cores = os.cpu_count()
  sampler = ClockSampler(local, ctx.pci_bus_id).start()    # before warm-up; returns after its first sample

  def make_model(precision, config=args.config):
    m = build_config(config, precision=precision, ctx=ctx, prior=True)
    cfg = m._transform_config["synthesis"]
    m.load_weights(synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=cfg["cls"]))
    m._ensure_native()
    return m

  precision = args.precision
  if precision == "auto":
    precision = "tc"
  model = make_model(precision)       # no fallback: a box without the tcgen05 path fails here, loudly

  zs, ys = model.latent_shapes(B, H, W)
  hyper = zs is not None                      # the factorized model (bls2017) has no z_hat and no scale indexes
  frame_h, band = H, None
  # this rank's shard of the seeded image list; `rotate` distinct batches so consecutive steps never reuse inputs
  sets = []
  for r in range(args.rotate):
    z, q = synthetic.make_latents(zs, ys, first_index=((0 if args.tile_frames else rank) * args.rotate + r) * B)
    sets.append((z, q))
  if args.tile_frames and world > 1:
    # every rank holds band `rank` (rows + halo) of the SAME frames: what the range decoder of a tiled stream would hand it
    plan = model.band_plan((H, W), world)
    if len(plan) != world:
      raise SystemExit(f"--tile-frames: a {W}x{H} frame has only {len(plan)} latent-row bands to give to {world} ranks")
    band = plan[rank]
    sets = [model.band_inputs(z, q, band) for z, q in sets]
    H = band.sub_h                            # from here on this rank decodes a (sub_h x W) "image"; its own rows are band.rows
    zs, ys = model.latent_shapes(B, H, W)
    assert tuple(sets[0][1].shape) == tuple(ys), (sets[0][1].shape, ys)
  dev = [(ctx.to_device(z) if hyper else None, ctx.to_device(q)) for z, q in sets]
  out_dev = dict(image=ctx.empty((B, H, W, 3), np.uint8))
  if hyper:
    out_dev["idx"] = ctx.empty(ys, np.uint8)

  def step_dev(i):
    dz, dq = dev[i % args.rotate]
    model.decompress(dz, dq, (H, W), out=out_dev, sync=False)

  def barrier():
    ctx.sync()
    if group is not None:
      group.barrier()

  for i in range(args.warmup):
    step_dev(i)
  barrier()
  k0 = ctx.launch_counts
  e0, e1 = ctx.event(), ctx.event()
  barrier()
  tw0 = time.perf_counter()
  e0.record()
  for i in range(args.steps):
    step_dev(i)
  host_ms = (time.perf_counter() - tw0) * 1e3 / args.steps   # host time to enqueue one step (must stay below the device time)
  e1.record()
  ctx.sync()
  tw1 = time.perf_counter()
  ms = e0.elapsed_ms(e1)
  barrier()
  k1 = ctx.launch_counts
  kinds = {k: k1[k] - k0[k] for k in k1}
  launches = kinds["total"]
  clocks = sampler.window(tw0, tw1)

  # ---- e2e: page-locked HOST buffers through the public streaming API (DecodePipeline): every step uploads its
  # symbols (H2D) and downloads the image (D2H) inside the timed region; copies overlap the decode of the
  # neighbouring steps on separate streams ----
  from shallow_ntc_b200 import DecodePipeline

  def run_e2e(q_dtype, return_idx, write_combined=False):
    pipe = DecodePipeline(model, B, (H, W), q_dtype=q_dtype, depth=args.e2e_depth, return_idx=return_idx, host_slots=2, write_combined=write_combined)
    hs = [pipe.host_slot(i) for i in range(2)]
    for i, h in enumerate(hs):
      if hyper:
        h["z"][...] = sets[i][0]
      h["q"][...] = sets[i][1].astype(q_dtype)
    for i in range(3):
      h = hs[i % 2]
      pipe.submit(h["z"], h["q"], h["image"], h["idx"])
    pipe.drain()
    barrier()
    f0, f1 = ctx.event(), ctx.event()
    f0.record(pipe.s_in.handle)
    for i in range(args.steps):
      h = hs[i % 2]
      pipe.submit(h["z"], h["q"], h["image"], h["idx"])
    f1.record(pipe.s_out.handle)
    pipe.drain()
    last = hs[(args.steps - 1) % 2]
    return dict(ms=f0.elapsed_ms(f1), h2d=int(pipe.in_bytes), d2h=int(pipe.down_bytes if return_idx else pipe.img_bytes),
                image=last["image"].copy(), q_dtype=np.dtype(q_dtype).name, return_idx=bool(return_idx and hyper))

  def link_probe(h2d_bytes, d2h_bytes, n=10):
    """Host<->device copy rate of THIS box with all ranks copying at once: page-locked buffers of the e2e step's own sizes, H2D and
    D2H concurrently on two streams.  It is the roofline of the e2e number (PCIe / host memory, not the GPU)."""
    from shallow_ntc_b200.pipeline import _Stream
    import ctypes as C
    from shallow_ntc_b200._lib import lib, check
    src = ctx.pinned_empty((h2d_bytes,), np.uint8)
    src[...] = 1
    dsrc = ctx.empty((h2d_bytes,), np.uint8)
    dst = ctx.pinned_empty((d2h_bytes,), np.uint8)
    ddst = ctx.empty((d2h_bytes,), np.uint8)
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

  # headline e2e = the pipeline's defaults (int16 symbols up, image down), taken right after the timed region (same clock
  # regime as `value`); then the per-layer profile; the other hand-overs are reported beside it
  e2e_runs = {"int16": run_e2e(np.int16, False)}

  # per-layer breakdown (CUDA events around every layer) in a separate, untimed pass: the extra event records
  # would otherwise sit between the kernels of the timed region
  model.profile_layers(True)
  tp0 = time.perf_counter()
  prof_steps = min(max(args.steps, 20), 50)
  for i in range(prof_steps):
    step_dev(i)
  ctx.sync()
  tp1 = time.perf_counter()
  model.profile_layers(False)
  prof = model.layer_profile()
  clocks_prof = sampler.window(tp0, tp1)

  e2e_runs["int16+idx"] = run_e2e(np.int16, True)
  e2e_runs["float32+idx"] = run_e2e(np.float32, True)      # round-1 headline, kept for continuity
  e2e_runs["int8"] = run_e2e(np.int8, False)
  e2e_runs["int16/wc"] = run_e2e(np.int16, False, write_combined=True)
  head = e2e_runs["int16"]
  link = link_probe(head["h2d"], head["d2h"])

  # final quality sum over ranks (the only collective: one NCCL all-reduce of 5 doubles through libsntc)
  orig = synthetic.make_original(head["image"][:2], first_index=rank * B)
  zq = sets[(args.steps - 1) % 2]
  met = model.decompress(zq[0][:2] if hyper else None, zq[1][:2], (H, W), original=orig, return_bits=hyper)
  # [sum psnr, sum mse, sum bits_y, sum bits_z, n_images]: the reference averages per-image metrics (mshyper/models.py:300-317)
  qsum = np.array([met["psnr"].sum(), met["mse"].sum(), met["bits_y"].sum() if hyper else 0.0, met["bits_z"].sum() if hyper else 0.0,
                   float(len(met["psnr"]))])
  # cost of asking for the rate term as well (bits_y from the hyper-head's raw sigma + bits_z kernel), device-resident
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

  # ---- sustained: the same step back to back for >= 2 s (the timed region above is a short burst at boost clocks; this one runs
  # into the 1 kW power cap).  Last, so that every other number of this line is taken in ONE clock regime, the timed region's ----
  sustained = None
  if not args.no_side:
    n_sus = max(args.steps, int(2000.0 / max(ms / args.steps, 1e-3)) + 1)
    s0, s1 = ctx.event(), ctx.event()
    ts0 = time.perf_counter()
    s0.record()
    for i in range(n_sus):
      step_dev(i)
    s1.record()
    ctx.sync()
    ts1 = time.perf_counter()
    ms_sus = s0.elapsed_ms(s1) / n_sus
    sustained = dict(steps=n_sus, ms_per_step=ms_sus, value=B * H * W / (ms_sus * 1e-3) / 1e6, unit="Mpx/s per GPU (rank 0)",
                     clocks=sampler.window(ts0 + 0.5, ts1))

  link_keys = sorted(link)
  link_sum = np.array([link[k] for k in link_keys])
  e2e_keys = list(e2e_runs)
  times = np.array([ms] + [e2e_runs[k]["ms"] for k in e2e_keys])
  if group is not None:
    link_sum = group.allreduce_sum(link_sum)                 # aggregate GB/s over ranks
    times = group.allreduce_max(times)                       # device times: max over ranks
    qsum = par.reduce_metric_sums(group, qsum)               # NCCL: the only collective on results
  ms = float(times[0])
  for k, t in zip(e2e_keys, times[1:]):
    e2e_runs[k]["ms"] = float(t)

  if rank == 0:
    headline = args.config == "two_layer_syn" and (H, W) == (512, 768)
    workload = (f"mshyper two_layer_syn decode (BASELINE configs[1]): {B} x 768x512 per GPU, random-init 'stress' weights" if headline else
                f"{args.config} decode (side run, not the headline workload): {B} x {W}x{frame_h} per GPU, random-init 'stress' weights")
    if band is not None:
      workload = (f"{args.config} decode, intra-frame sharding: {B} frames of {W}x{frame_h} split into {world} latent-row bands, one band (+ recomputed halo) per GPU; "
                  f"rank 0 decodes rows {band.rows} from a {H}-row sub-frame")
    px_step = B * frame_h * W if band is not None else world * B * H * W      # tiled: all ranks work on the same frames
    mpx = lambda t_ms: px_step * args.steps / (t_ms * 1e-3) / 1e6
    value = mpx(ms)
    # dominant kernel = the layer with the largest share of device time
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    roof = None
    # algorithmic HBM bytes per step: symbols in (f32) + image and index map out
    alg_bytes = sets[0][1].nbytes + (sets[0][0].nbytes if hyper else 0) + B * H * W * 3 + (int(np.prod(ys)) if hyper else 0)
    if dom[0] is not None and dom[1]["macs"] > 0:
      per_launch_ms = dom[1]["ms"] / dom[1]["n"]
      ach = 2.0 * dom[1]["macs"] / (per_launch_ms * 1e-3) / 1e12
      regime, peak = regime_peak(peaks, clocks_prof)
      roof = dict(bound="tensor", kernel=dom[0], achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak,
                  peak_regime=regime, peak_source=f"{peaks['source']}: cuBLAS bf16 dense, {regime} figure chosen from the clocks sampled while the kernel was timed",
                  frac_vs_burst=ach / peaks["tflops_burst"], frac_vs_sustained=ach / peaks["tflops_sustained"],
                  mma_passes=MMA_PASSES, frac_executed=MMA_PASSES * ach / peak,
                  clocks_while_timed=clocks_prof, traffic=ncu_traffic(dom[0], B), share_of_step=per_launch_ms / (ms / args.steps),
                  ms_per_launch=per_launch_ms, algorithmic_flops_per_launch=2.0 * dom[1]["macs"],
                  note="achieved = algorithmic FLOPs (2 * MACs of the layer, reference counting) / CUDA-event time of the layer; the split-fp16 product "
                       "issues 3 tensor-core MACs per algorithmic MAC, so frac <= 1/3 by construction and frac_executed is the tensor-pipe view",
                  hbm_view=dict(algorithmic_gbs=world * alg_bytes * args.steps / (ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"]))
    both_up = link_sum[link_keys.index("both_up_gbs")]
    both_down = link_sum[link_keys.index("both_down_gbs")]

    def e2e_rec(r):
      bound = px_step / max(world * r["h2d"] / (both_up * 1e9), world * r["d2h"] / (both_down * 1e9)) / 1e6
      v = mpx(r["ms"])
      return dict(value=v, ms_per_step=r["ms"] / args.steps, h2d_bytes_per_step=r["h2d"], d2h_bytes_per_step=r["d2h"], symbols=r["q_dtype"],
                  idx_downloaded=r["return_idx"], link_bound_mpx=bound, frac_of_link_bound=v / bound, frac_of_device_value=v / value)
    e2e = e2e_rec(head)
    e2e.update(unit="Mpx/s", api=f"DecodePipeline.submit defaults: int16 symbols from page-locked host memory, image back to page-locked host memory, one copy "
                                 f"per direction, depth {args.e2e_depth}; idx stays on the device (the caller of a symbol-level decode already holds the rows)",
               host_numa_binding=numa, variants={k: e2e_rec(v) for k, v in e2e_runs.items() if k != "int16"},
               host_link=dict({k: round(float(v), 1) for k, v in zip(link_keys, link_sum)},
                              note="aggregate page-locked copy GB/s over all ranks copying at once, buffers of the e2e step's sizes ('both' = the two "
                                   "directions concurrently): the roofline of the e2e numbers on this box"))
    line = dict(metric="decoded Mpx/s", value=value, unit="Mpx/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong" if band is not None else "weak", vs_baseline=None,
                dtype="f32" if precision == "fp32" else "f16x3-split/f32-accum", data="synthetic",
                config=dict(workload=workload,
                            images_per_gpu=B, precision=precision, index_rounding=model.index_rounding,
                            l2="inputs rotate over %d distinct batches; per-step working set > 126 MB L2" % args.rotate,
                            layers_ms={k: round(v["ms"] / prof_steps, 4) for k, v in prof.items()},   # per step (a layer may take several launches)
                            layer_launches_per_step={k: round(v["n"] / prof_steps, 2) for k, v in prof.items() if v["n"] != prof_steps},
                            launches_by_family=kinds,
                            mean_psnr_db=float(qsum[0] / qsum[4]), mean_bpp_synthetic=float((qsum[2] + qsum[3]) / qsum[4] / (H * W)),
                            ms_per_step_with_rate_term=ms_rd, host_enqueue_ms_per_step=round(host_ms, 4),
                            collective="NCCL all-reduce through libsntc (sntc_comm_*)" if group is not None else "none (1 GPU)"),
                clocks=clocks, gpu_launches=int(launches), e2e=e2e, roofline=roof)
    if sustained is not None:
      line["sustained"] = sustained
    if world == 1 and not args.no_side:
      # BASELINE configs[0]: ONE jpegl 768x512 image; device-resident latency of a single decode (launch-bound regime)
      try:
        mj = make_model(precision, "jpegl")
        zj_s, yj_s = mj.latent_shapes(1, 512, 768)
        zj, qj = synthetic.make_latents(zj_s, yj_s)
        dzj, dqj = ctx.to_device(zj), ctx.to_device(qj.astype(np.int16))
        oj = dict(image=ctx.empty((1, 512, 768, 3), np.uint8), idx=ctx.empty(yj_s, np.uint8))
        for _ in range(10):
          mj.decompress(dzj, dqj, (512, 768), out=oj, sync=False)
        ctx.sync()
        j0, j1 = ctx.event(), ctx.event()
        nj = 200
        j0.record()
        for _ in range(nj):
          mj.decompress(dzj, dqj, (512, 768), out=oj, sync=False)
        j1.record()
        ctx.sync()
        msj = j0.elapsed_ms(j1) / nj
        line["side"] = dict(jpegl_b1=dict(workload="mshyper jpegl decode of ONE 768x512 image (BASELINE configs[0]), device-resident, back to back",
                                          ms_per_decode=msj, value=512 * 768 / (msj * 1e-3) / 1e6, unit="Mpx/s", decodes=nj))
      except Exception as e:
        line["side"] = dict(jpegl_b1=dict(error=f"{type(e).__name__}: {e}"))
    if world == 1 and hyper and not args.no_io_stage:
      # The I/O stage either side of the hot path (north_star: "range decoding ... stays in the host coder ... timed
      # separately"): container bytes -> symbols on the host threads, on a sample of the same workload, with the GPU
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
      # bounded CPU sample (a reported baseline, not the target) + the checker leg on one image
      res, ncores = cpu_arm(args.config, B, 2, 1, budget_s=20.0)
      ok = {k: v for k, v in res.items() if "value" in v}
      best = max(ok, key=lambda k: ok[k]["value"])
      line["cpu_baseline"] = dict(value=ok[best]["value"], unit="Mpx/s", cores=ncores, kind="port",
                                  sample=f"{ok[best]['images_per_step']} of the {B} images per step x 2 steps, {best}; " +
                                         ", ".join(f"{k}: {v['value']:.2f} Mpx/s" for k, v in ok.items()) + " (TF-2.10 is not installable offline)")
      try:
        z0, q0 = sets[0]
        g = model.decompress(z0[:1] if hyper else None, q0[:1], (H, W), return_float=True)
        g["index_rounding"] = model.index_rounding
        line["cpu_baseline"]["parity_sample"] = parity_sample(args.config, g, 1)
      except Exception as e:
        line["cpu_baseline"]["parity_sample"] = dict(error=f"{type(e).__name__}: {e}")
    print(json.dumps(line), file=json_out, flush=True)
  sampler.stop()
  if group is not None:
    group.barrier()
    group.close()
  return 0


if __name__ == "__main__":
  sys.exit(main())
