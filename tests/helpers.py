"""Shared helpers for the parity tests: run a config through libsntc and through the oracle."""
import numpy as np

from shallow_ntc_b200 import build_config, synthetic
from oracle import ntc_oracle as O

# tolerances (BASELINE.json north_star): float reconstruction 1e-3 max-abs in pixel space [-0.5, 0.5]
# (same scale as [0,1]); PSNR delta 0.01 dB; uint8 within 1 LSB; scale-table rows exact outside the
# tie margin |i_c - (k + .5)| < IDX_MARGIN * max(1, i_c)
RECON_TOL = 1e-3
PSNR_TOL = 0.01
IDX_MARGIN = {"fp32": 5e-5, "tc": 2e-4}   # relative to max(1, i_c): d i_c = i_c * d raw_sigma
MU_TOL = 1e-4
U8_FLIP_FRAC = {"fp32": 1e-3, "tc": 5e-3}


def syn_kwargs(model):
  cfg = model._transform_config["synthesis"]
  return cfg["cls"], {k: v for k, v in cfg.items() if k != "cls"}


def hyper_kwargs(model):
  cfg = model._transform_config.get("hyper_synthesis")
  if cfg is None:
    return "HyperSynthesis", {}
  return cfg["cls"], {k: v for k, v in cfg.items() if k not in ("cls", "bottleneck_size")}


def make_case(name, B, H, W, kind="stress", precision="fp32", ctx=None, first_index=0, index_rounding=None, **model_kw):
  if index_rounding is not None:
    model_kw["index_rounding"] = index_rounding
  model = build_config(name, precision=precision, ctx=ctx, **model_kw)
  cls, _ = syn_kwargs(model)
  wts = synthetic.make_weights(model.variable_shapes(), kind, synthesis_cls=cls)
  model.load_weights(wts)
  zs, ys = model.latent_shapes(B, H, W)
  z, q = synthetic.make_latents(zs, ys, first_index=first_index)
  return model, wts, z, q


def oracle_decode(model, wts, z, q, H, W, original=None, dtype=np.float64, gemm_form=False, index_rounding=None):
  cls, kw = syn_kwargs(model)
  index_rounding = index_rounding or model.index_rounding
  if model.hyperprior:
    hcls, hkw = hyper_kwargs(model)
    return O.mshyper_decode(wts, cls, z, q, H, W, kw, hcls, hkw, original, dtype, gemm_form, index_rounding)
  return O.factorized_decode(wts, cls, q, H, W, kw, original, dtype, gemm_form)


def check_against_oracle(got, ref, hyper=True, recon_tol=RECON_TOL, precision="fp32"):
  """The correctness gates of SURVEY 8(d).  Returns a dict of measured deviations."""
  rep = {}
  err = np.abs(got["float"].astype(np.float64) - ref["recon"])
  rep["recon_max_abs"] = float(err.max())
  assert rep["recon_max_abs"] < recon_tol, rep
  d = np.abs(got["image"].astype(np.int16) - ref["recon_u8"].astype(np.int16))
  rep["u8_max_diff"] = int(d.max())
  rep["u8_frac_diff"] = float((d > 0).mean())
  # a pixel flips by one LSB iff its float value sits within the float error of a rounding boundary:
  # expected fraction ~ 2 * 255 * mean|err| (fp32 path ~1e-4, split-fp16 tensor path a few 1e-3 on the deep decoders)
  assert rep["u8_max_diff"] <= 1 and rep["u8_frac_diff"] < U8_FLIP_FRAC[precision], rep
  if hyper:
    # y_hat = q + mu is a single fp32 add: mu_gpu is recovered exactly as y_hat - q when |q| small
    mu_err = np.abs(got["y_hat"].astype(np.float64) - ref["y_hat"].astype(np.float64))
    rep["yhat_max_abs"] = float(mu_err.max())
    assert rep["yhat_max_abs"] < MU_TOL * np.maximum(1.0, np.abs(ref["y_hat"]).max() / 64), rep
    margin = IDX_MARGIN[precision] * np.maximum(1.0, ref["i_c"])
    far = ref["idx_dist"] > margin
    rep["idx_in_margin"] = int((~far).sum())
    rep["idx_mismatch_outside_margin"] = int((got["idx"][far] != ref["idx"][far]).sum())
    rep["idx_mismatch_in_margin"] = int((got["idx"][~far] != ref["idx"][~far]).sum())
    assert rep["idx_mismatch_outside_margin"] == 0, rep
    assert rep["idx_in_margin"] < 0.02 * far.size, rep
    assert np.abs(got["idx"].astype(int) - ref["idx"].astype(int)).max() <= 1, "a boundary flip moves the row by at most one"
  return rep
