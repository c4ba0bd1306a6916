"""not-gpu: the CPU arms of bench.py (oracle/ref_configs.py, oracle/torch_ref.py), the file rendezvous of the NCCL group,
and the reference arm's promise not to load the product library."""
import json
import multiprocessing as mp
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ntc_oracle as O
from oracle import ref_configs as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["jpegl", "two_layer_syn", "two_layer_syn2", "two_layer_syn2:24", "two_layer_syn2:48", "mbt2018", "bls2017"])
def test_reference_side_shapes_are_the_products(name):
  """oracle/ref_configs.py derives variable and latent shapes from the reference's config files on its own; they must be
  exactly what the product asks for (so the two bench arms decode the same weights and symbols)."""
  from shallow_ntc_b200 import build_config
  m = build_config(name, prior=True)
  assert R.variable_shapes(name, prior=True) == {k: tuple(v) for k, v in m.variable_shapes().items()}
  for B, H, W in ((1, 512, 768), (3, 100, 150), (2, 1200, 1200), (1, 2160, 3840)):
    assert R.latent_shapes(name, B, H, W) == m.latent_shapes(B, H, W)
  cfg = R.get_config(name)
  assert cfg["synthesis"] == dict(m._transform_config["synthesis"])


def test_standalone_synthetic_module_gives_the_same_data():
  from shallow_ntc_b200 import synthetic, build_config
  cfg, wts, z, q = R.make_case("two_layer_syn", 2, 64, 128, "stress", first_index=7)
  m = build_config("two_layer_syn")
  w2 = synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls="TwoLayerResSynthesis")
  z2, q2 = synthetic.make_latents(*m.latent_shapes(2, 64, 128), first_index=7)
  assert all(np.array_equal(wts[k], w2[k]) for k in w2) and np.array_equal(z, z2) and np.array_equal(q, q2)


@pytest.mark.parametrize("name,H,W", [("jpegl", 64, 128), ("two_layer_syn", 100, 150), ("two_layer_syn2:24", 64, 64), ("mbt2018", 64, 64), ("bls2017", 48, 80)])
def test_torch_cpu_statement_matches_the_oracle(name, H, W):
  """The oneDNN stand-in is a third, independent formulation (conv_transpose2d + crop): float32 against float64 T0."""
  from oracle import torch_ref as T
  cfg, wts, z, q = R.make_case(name, 2, H, W, "stress")
  syn = cfg["synthesis"]
  kw = {k: v for k, v in syn.items() if k != "cls"}
  ref = O.mshyper_decode(wts, syn["cls"], z, q, H, W, kw) if cfg["hyperprior"] else O.factorized_decode(wts, syn["cls"], q, H, W, kw)
  got = T.TorchDecoder(cfg, wts)(z, q, H, W)
  assert got["image"].shape == ref["recon_u8"].shape
  assert np.abs(got["recon"] - ref["recon"]).max() < 2e-4
  d = np.abs(got["image"].astype(int) - ref["recon_u8"].astype(int))
  assert d.max() <= 1 and (d > 0).mean() < 5e-3
  if cfg["hyperprior"]:
    far = ref["idx_dist"] > 2e-4 * np.maximum(1, ref["i_c"])
    assert np.array_equal(got["idx"][far], ref["idx"][far])


def _rendezvous_worker(rank, path, q):
  from shallow_ntc_b200.parallel import exchange_bytes
  payload = bytes(range(128)) if rank == 0 else None
  q.put((rank, exchange_bytes(path, rank, payload, 128, timeout_s=30.0)))


def test_file_rendezvous_hands_the_id_to_every_rank(tmp_path):
  """NcclGroup's id exchange (rank 0 publishes atomically, the others poll) with 3 processes, readers started first."""
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  path = str(tmp_path / "sntc_nccl_test.id")
  procs = [ctx.Process(target=_rendezvous_worker, args=(r, path, q)) for r in (2, 1, 0)]
  for p in procs:
    p.start()
  got = dict(q.get(timeout=60) for _ in range(3))
  for p in procs:
    p.join(timeout=30)
    assert p.exitcode == 0
  assert all(got[r] == bytes(range(128)) for r in range(3))
  from shallow_ntc_b200.parallel import exchange_bytes
  with pytest.raises(TimeoutError):
    exchange_bytes(str(tmp_path / "never"), 1, None, 128, timeout_s=0.05)


def test_reference_arm_runs_without_the_product_library():
  """bench.py --impl reference: a JSON line with the contract's keys, produced by a process that never mapped libsntc.so."""
  code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--batch', '1', '--height', '64', '--width', '64'];\n"
          "try:\n  runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit:\n  pass\n"
          "maps = open('/proc/self/maps').read()\n"
          "print('LIBSNTC_MAPPED' if 'libsntc' in maps else 'LIBSNTC_ABSENT', file=sys.stderr)\n"
          "print('PKG_IMPORTED' if 'shallow_ntc_b200' in sys.modules else 'PKG_ABSENT', file=sys.stderr)\n")
  r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
  assert r.returncode == 0, r.stderr[-2000:]
  assert "LIBSNTC_ABSENT" in r.stderr and "PKG_ABSENT" in r.stderr, r.stderr[-500:]
  line = json.loads(r.stdout.strip().splitlines()[-1])
  assert line["impl"] == "reference" and line["unit"] == "Mpx/s" and line["higher_is_better"] is True
  assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == os.cpu_count()
  assert line["e2e"] == dict(value=line["value"], unit="Mpx/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
  assert set(line["config"]["all"]) == {"numpy_T1", "torch_onednn"}
