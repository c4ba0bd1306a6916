"""-m gpu: the parity holes the round-1 review listed -- every registered decoder-side transform against the float64
oracle on the GPU, both index rules, the tensor-core versions of the fp32-only identities, full-size images of the
side configurations, and proof of WHICH kernel family served each tensor-core model (sntc_launch_counts)."""
import contextlib

import numpy as np
import pytest

from shallow_ntc_b200 import Model, FactorizedModel, synthetic
from helpers import make_case, oracle_decode, check_against_oracle, syn_kwargs, PSNR_TOL

pytestmark = pytest.mark.gpu

ELIC = dict(cls="ElicAnalysis", channels=(192, 192, 192, 320))


@contextlib.contextmanager
def launches(ctx):
  """dict that receives the per-family kernel launch counts of the block (Context.launch_counts deltas)."""
  before, out = ctx.launch_counts, {}
  yield out
  after = ctx.launch_counts
  out.update({k: after[k] - before[k] for k in after})


def case_from_config(cfg, B, H, W, kind, precision, ctx, hyperprior=True, out_gain=None, **model_kw):
  cls = FactorizedModel if not hyperprior else Model
  model = cls({k: dict(v) for k, v in cfg.items()}, precision=precision, ctx=ctx, **model_kw)
  wts = synthetic.make_weights(model.variable_shapes(), kind, synthesis_cls=cfg["synthesis"]["cls"], out_gain=out_gain)
  model.load_weights(wts)
  zs, ys = model.latent_shapes(B, H, W)
  z, q = synthetic.make_latents(zs, ys)
  return model, wts, z, q


def run_and_check(model, wts, z, q, H, W, precision, ctx, expect_f32_bands=False):
  ref = oracle_decode(model, wts, z, q, H, W)
  orig = synthetic.make_original(ref["recon_u8"])
  with launches(ctx) as n:
    got = model.decompress(z, q, (H, W), return_float=True, return_yhat=True, original=orig)
  rep = check_against_oracle(got, ref, hyper=model.hyperprior, precision="fp32" if precision == "fp32" else "tc")
  from oracle import ntc_oracle as O
  _, ref_psnr = O.mse_psnr(orig, ref["recon_u8"])
  assert np.all(np.abs(got["psnr"] - ref_psnr) < PSNR_TOL), (got["psnr"], ref_psnr)
  if precision != "fp32":
    assert n["band_tc"] > 0, n
    if not expect_f32_bands:
      assert n["band_f32"] == 0 and n["final_f32"] == 0, f"a tensor-core model silently ran FFMA kernels: {n}"
  else:
    assert n["band_tc"] == 0 and n["tail_mma"] == 0 and n["tail_tc"] == 0, n
  return got, ref, rep, n


# --------------------------------------------------------------------------------------------------
# transforms that had no GPU test (common/transforms.py:195-206, 250-262, 291-293, 364-377)

@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("kind", ["init", "stress"])
def test_jpeg_like_hyper_synthesis(gpu_ctx, kind, precision):
  """JPEGLikeHyperSynthesis(bottleneck_size, kernel_size=6): ONE ConvT(6, 4) in the hyper position, its epilogue carries the
  whole entropy-model glue (split / exp / clamp / row, q + mu)."""
  cfg = dict(analysis=ELIC, synthesis=dict(cls="JPEGLikeSynthesis", kernel_size=18, strides=16),
             hyper_synthesis=dict(cls="JPEGLikeHyperSynthesis", bottleneck_size=320, kernel_size=6))
  model, wts, z, q = case_from_config(cfg, 2, 100, 150, kind, precision, gpu_ctx)
  assert model.downsample_factor == 64 and z.shape == (2, 2, 3, 320) and q.shape == (2, 8, 12, 320)
  print(run_and_check(model, wts, z, q, 100, 150, precision, gpu_ctx)[2:])


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("kind", ["init", "stress"])
def test_hyper_synthesis_small(gpu_ctx, kind, precision):
  """HyperSynthesisSmall: tfc SignalConv2D 5x5 up2 + relu -> 3x3 up1 ('same_zeros' alignment p = (k-1)//2 and the [kh,kw,Cin,Cout]
  kernel layout in the hyper position); upsample 2, so the model's downsample factor is 32."""
  cfg = dict(analysis=ELIC, synthesis=dict(cls="TwoLayerResSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5),
                                           activation_type="igdn", res_type="conv"),
             hyper_synthesis=dict(cls="HyperSynthesisSmall", bottleneck_size=320))
  model, wts, z, q = case_from_config(cfg, 2, 96, 160, kind, precision, gpu_ctx)
  assert model.downsample_factor == 32 and z.shape == (2, 3, 5, 320) and q.shape == (2, 6, 10, 320)
  print(run_and_check(model, wts, z, q, 96, 160, precision, gpu_ctx)[2:])


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("activation_type", ["leaky_relu", "igdn", "relu"])
def test_cnn_synthesis(gpu_ctx, activation_type, precision):
  """CNNSynthesis: 4 x Keras ConvT k5 s2; with 'igdn' ONE GDN1 object (one beta / gamma) is shared by layers 0-2
  (common/transforms.py:199-204)."""
  cfg = dict(analysis=dict(cls="CNNAnalysis", channels_base=192, output_channels=320),
             synthesis=dict(cls="CNNSynthesis", channels_base=192, output_channels=3, activation_type=activation_type))
  # three IGDN1 stages grow the signal ~40x more than leaky_relu: scale the last layer so the image stays in range
  model, wts, z, q = case_from_config(cfg, 1, 64, 128, "stress", precision, gpu_ctx, out_gain=0.005 if activation_type == "igdn" else None)
  if activation_type == "igdn":
    assert sum(k.endswith(".gamma") for k in wts) == 1 and wts["synthesis.activation.gamma"].shape == (192, 192)
  print(run_and_check(model, wts, z, q, 64, 128, precision, gpu_ctx)[2:])


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("use_bias", [True, False])
def test_jpeg_like_synthesis_use_offset(gpu_ctx, use_bias, precision):
  """JPEGLikeSynthesis(use_offset=True) appends a constant-1 channel (Cin = 321, transforms.py:291-293).  321 channels are
  not TMA-addressable: with precision='tc' this ONE layer runs on the FFMA band GEMM -- loudly (band_f32 > 0), the
  hyper-synthesis stays on tcgen05."""
  cfg = dict(analysis=ELIC, synthesis=dict(cls="JPEGLikeSynthesis", kernel_size=18, strides=16, use_offset=True, use_bias=use_bias))
  model, wts, z, q = case_from_config(cfg, 2, 128, 192, "stress", precision, gpu_ctx)
  assert wts["synthesis.conv.kernel"].shape == (18, 18, 3, 321) and ("synthesis.conv.bias" in wts) == use_bias
  got, ref, rep, n = run_and_check(model, wts, z, q, 128, 192, precision, gpu_ctx, expect_f32_bands=True)
  if precision == "tc":
    assert n["band_f32"] == 9 and n["band_tc"] == 3, n     # the 9 bands of ConvT(18, 16) / the 3 hyper-synthesis layers
  print(rep, n)


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("cls,act", [("TwoLayerSynthesis", "relu"), ("TwoLayerSynthesis", "gdn"), ("TwoLayerSynthesis", "leaky_relu"),
                                     ("TwoLayerResSynthesis", "relu"), ("TwoLayerResSynthesis", "gdn")])
def test_two_layer_activation_variants(gpu_ctx, cls, act, precision):
  """activation_type other than the shipped 'igdn' (get_activation_op, transforms.py:66-78): relu / leaky_relu are fused
  into the conv, (non-inverse) GDN1 divides by the norm pool."""
  syn = dict(cls=cls, channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5), activation_type=act)
  if cls == "TwoLayerResSynthesis":
    syn["res_type"] = "conv"
  cfg = dict(analysis=ELIC, synthesis=syn)
  model, wts, z, q = case_from_config(cfg, 2, 100, 150, "stress", precision, gpu_ctx)
  print(run_and_check(model, wts, z, q, 100, 150, precision, gpu_ctx)[2:])


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("act,B,H,W", [("igdn", 2, 100, 150), ("relu", 1, 128, 192), ("igdn", 1, 512, 768)])
def test_two_layer_res_synthesis_d2s(gpu_ctx, act, B, H, W, precision):
  """TwoLayerResSynthesis(res_type="d2s") (common/transforms.py:339-348): the residual is depth_to_space(2) -> Conv2D 1x1 (192)
  + leaky_relu -> depth_to_space(2) -> Conv2D 1x1 (48) + leaky_relu -> depth_to_space(2).  On the device each
  depth_to_space + 1x1 conv pair is ONE k = s = 2 transposed conv (band GEMM, tcgen05 under 'tc'), the last depth_to_space is
  the addressing of the residual in the activation stage."""
  syn = dict(cls="TwoLayerResSynthesis", channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5), activation_type=act, res_type="d2s")
  model, wts, z, q = case_from_config(dict(analysis=ELIC, synthesis=syn), B, H, W, "stress", precision, gpu_ctx)
  assert wts["synthesis.res.conv_0.kernel"].shape == (1, 1, 80, 192) and wts["synthesis.res.conv_1.kernel"].shape == (1, 1, 48, 48)
  if H * W <= 128 * 192:
    got, ref, rep, n = run_and_check(model, wts, z, q, H, W, precision, gpu_ctx)
    print(rep, n)
  else:   # full size: tensor path against the fp32 CUDA-core path of the same library (the float64 oracle takes minutes here)
    got = model.decompress(z, q, (H, W), return_float=True)
    if precision == "tc":
      base, _, _, _ = case_from_config(dict(analysis=ELIC, synthesis=syn), B, H, W, "stress", "fp32", gpu_ctx)
      f32 = base.decompress(z, q, (H, W), return_float=True)
      assert np.abs(f32["float"] - got["float"]).max() < 1e-4
  fast = model.decompress(z, q, (H, W))
  assert np.array_equal(fast["image"], got["image"])


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("gdn_form", ["gdn1", "classic"])
def test_mbt2018_gdn_forms(gpu_ctx, gdn_form, precision):
  """MBT2018Synthesis builds tfc.GDN(inverse=True) with the tfc 2.x defaults (alpha = epsilon = 1: IGDN1, oracle A5);
  the alpha=2 / epsilon=.5 form stays available as gdn_form='classic'."""
  cfg = dict(analysis=dict(cls="MBT2018Analysis", channels_base=192, output_channels=320),
             synthesis=dict(cls="MBT2018Synthesis", channels_base=192, output_channels=3, gdn_form=gdn_form))
  model, wts, z, q = case_from_config(cfg, 1, 64, 128, "stress", precision, gpu_ctx, out_gain=0.07 if gdn_form == "classic" else None)
  print(run_and_check(model, wts, z, q, 64, 128, precision, gpu_ctx)[2:])


def test_hidden_width_64_on_the_cuda_core_path(gpu_ctx):
  """GDN1 over 64 channels needs 49 920 B of dynamic shared memory in act_res_kernel (> the 48 KB default)."""
  for name in ("two_layer_syn2:64", "two_layer_syn:64"):
    model, wts, z, q = make_case(name, 1, 64, 64, "stress", "fp32", gpu_ctx)
    ref = oracle_decode(model, wts, z, q, 64, 64)
    got = model.decompress(z, q, (64, 64), return_float=True, return_yhat=True)
    print(name, check_against_oracle(got, ref, precision="fp32"))


# --------------------------------------------------------------------------------------------------
# scale-table row rule (A6): both selectable rules on both precisions

@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("rule", ["trunc", "rint"])
def test_index_rounding_rules(gpu_ctx, rule, precision):
  model, wts, z, q = make_case("two_layer_syn", 2, 128, 192, "stress", precision, gpu_ctx, index_rounding=rule)
  assert model.index_rounding == rule
  ref = oracle_decode(model, wts, z, q, 128, 192)
  got = model.decompress(z, q, (128, 192), return_float=True, return_yhat=True)
  rep = check_against_oracle(got, ref, precision=precision)
  # the two rules really differ: rows of the other rule disagree on about half of the unclamped elements
  other = oracle_decode(model, wts, z, q, 128, 192, index_rounding="rint" if rule == "trunc" else "trunc")
  assert (other["idx"] != got["idx"]).mean() > 0.05
  # two-phase decode uses the same rule
  assert np.array_equal(model.decode_hyper(z), got["idx"])
  print(rule, precision, rep)


def test_default_rule_is_truncation(gpu_ctx):
  from shallow_ntc_b200.models import DEFAULT_INDEX_ROUNDING
  from shallow_ntc_b200 import build_config
  assert DEFAULT_INDEX_ROUNDING == "trunc" and build_config("jpegl").index_rounding == "trunc"


# --------------------------------------------------------------------------------------------------
# tensor-core versions of the fp32-only identities

def test_tc_yhat_is_bit_exact_single_add(gpu_ctx):
  """y_hat = fl32(q + mu_gpu) in the hyper-head epilogue of the tcgen05 kernel, for every symbol dtype."""
  model, wts, z, q = make_case("two_layer_syn", 2, 128, 192, "stress", "tc", gpu_ctx)
  mu = model.decompress(z, np.zeros_like(q), (128, 192), return_yhat=True)["y_hat"]
  for dt in (np.float32, np.int16, np.int8):
    yh = model.decompress(z, q.astype(dt), (128, 192), return_yhat=True)["y_hat"]
    assert np.array_equal(yh, (q + mu).astype(np.float32)), dt
  # the standalone transform call returns the same mu (same kernel, fp32 destination)
  hs = model.hyper_synthesis(z)
  assert np.array_equal(hs[..., :320], mu)


@pytest.mark.parametrize("name", ["two_layer_syn", "jpegl"])
def test_tc_single_image_equals_its_slot_in_the_batch_of_24(gpu_ctx, name):
  """A real codec needs encoder and decoder to produce IDENTICAL rows: B = 1 takes the narrow n-tiling (<= 64 columns per
  work item), B = 24 the wide one (tc_use_narrow); idx, image, y_hat and the rate must be bit-identical either way, at the
  full 512 x 768 size.  Also: decode determinism, and int8 symbols == float32 symbols."""
  B, H, W = 24, 512, 768
  model, wts, z, q = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  full = model.decompress(z, q, (H, W), return_yhat=True, return_bits=True)
  again = model.decompress(z, q.astype(np.int8), (H, W), return_yhat=True, return_bits=True)
  for k in ("image", "idx", "y_hat", "bits_y"):
    assert np.array_equal(full[k], again[k]), k
  for b in (0, 7, 23):
    one = model.decompress(z[b:b + 1], q[b:b + 1], (H, W), return_yhat=True, return_bits=True)
    for k in ("image", "idx", "y_hat"):
      assert np.array_equal(one[k], full[k][b:b + 1]), (k, b)
    assert np.allclose(one["bits_y"], full["bits_y"][b:b + 1], rtol=1e-12)
  pair = model.decompress(z[3:5], q[3:5], (H, W))
  assert np.array_equal(pair["idx"], full["idx"][3:5]) and np.array_equal(pair["image"], full["image"][3:5])


def test_tc_cta_group_1_and_2_are_bit_identical(gpu_ctx, monkeypatch):
  """cta_group::2 (M = 256, the default) and cta_group::1 (M = 128) accumulate every output element over the same K order."""
  outs = []
  for cg in ("2", "1"):
    monkeypatch.setenv("SNTC_TC_CTA_GROUP", cg)
    model, wts, z, q = make_case("two_layer_syn", 3, 200, 300, "stress", "tc", gpu_ctx)
    outs.append(model.decompress(z, q, (200, 300), return_yhat=True, return_float=True))
  for k in ("image", "idx", "y_hat", "float"):
    assert np.array_equal(outs[0][k], outs[1][k]), k


def test_integer_inputs_skip_the_lo_pass_bit_identically(gpu_ctx):
  """z_hat is integer-valued, so its fp16 lo plane is zero and hyper layer 0 skips the a_lo * w_hi pass (flag written by
  split_planes_kernel).  A non-integer z takes the 3-pass product; both agree with the oracle, and the integer case is
  bit-identical to the result with the skip defeated (one element perturbed far below fp16 resolution elsewhere)."""
  from oracle import ntc_oracle as O
  model, wts, z, q = make_case("two_layer_syn", 1, 64, 128, "stress", "tc", gpu_ctx)
  a = model.hyper_synthesis(z)
  z2 = z.copy()
  z2[0, 0, 0, 0] += 2.0 ** -14            # not fp16-exact next to an integer: lo plane non-zero, 3 passes everywhere
  b = model.hyper_synthesis(z2)
  ref2 = O.hyper_synthesis(wts, z2.astype(np.float64))
  assert np.abs(b - ref2).max() < 2e-4 and np.abs(a - O.hyper_synthesis(wts, z)).max() < 2e-4
  far = np.ones(a.shape[1:3], bool)
  far[:12, :12] = False                    # receptive field of z[0,0] through k5s2, k5s2, k3s1
  assert np.array_equal(a[0][far], b[0][far])


# --------------------------------------------------------------------------------------------------
# full-size images of the side configurations against the float64 oracle (one image each)

def _full_size(gpu_ctx, name, H, W, precision="tc"):
  model, wts, z, q = make_case(name, 1, H, W, "stress", precision, gpu_ctx)
  ref = oracle_decode(model, wts, z, q, H, W)
  with launches(gpu_ctx) as n:
    got = model.decompress(z, q, (H, W), return_float=True, return_yhat=True)
  rep = check_against_oracle(got, ref, hyper=model.hyperprior, precision=precision)
  assert n["band_tc"] > 0 and n["band_f32"] == 0 and n["final_f32"] == 0, n
  fast = model.decompress(z, q, (H, W))
  assert np.array_equal(fast["image"], got["image"])
  print(name, H, W, rep, n)


@pytest.mark.parametrize("c1", [12, 24, 48])
def test_full_size_two_layer_syn2_1200(gpu_ctx, c1):
  """BASELINE configs[2]: two_layer_syn2 (no residual), one 1200 x 1200 Tecnick-shaped image (padded to 1216), C1 = 12 / 24 / 48."""
  _full_size(gpu_ctx, f"two_layer_syn2:{c1}", 1200, 1200)


def test_full_size_mbt2018_512x768(gpu_ctx):
  """BASELINE configs[3]: mbt2018 mean-scale hyperprior with the 4-layer 192-channel IGDN synthesis, one 512 x 768 image."""
  _full_size(gpu_ctx, "mbt2018", 512, 768)


def test_full_size_bls2017_4k(gpu_ctx):
  """BASELINE configs[4]: factorized bls2017 (256 filters), one 3840 x 2160 frame."""
  _full_size(gpu_ctx, "bls2017", 2160, 3840)


def test_factorized_device_resident_yhat(gpu_ctx):
  """FactorizedModel.decompress(return_yhat=True) with device-resident float32 symbols: y_hat = q must be WRITTEN to the
  caller's device buffer (it used to alias the input and leave the output untouched)."""
  model, wts, z, q = make_case("bls2017", 2, 48, 80, "stress", "tc", gpu_ctx)
  dq = gpu_ctx.to_device(q)
  yh = gpu_ctx.to_device(np.full(q.shape, -777.0, np.float32))
  out = model.decompress(dq, (48, 80), return_yhat=True, out=dict(y_hat=yh))
  assert np.array_equal(out["y_hat"].to_host(), q)
  host = model.decompress(q, (48, 80), return_yhat=True)
  assert np.array_equal(host["y_hat"], q) and np.array_equal(host["image"], out["image"].to_host())


# --------------------------------------------------------------------------------------------------
# opt-in 2-pass synthesis (SNTC_PRECISION_TC_F16X3_SYN2): what it keeps exact and what it costs

def test_two_pass_synthesis_mode_keeps_the_entropy_side_exact(gpu_ctx):
  H, W = 512, 768
  m3, wts, z, q = make_case("two_layer_syn", 2, H, W, "stress", "tc", gpu_ctx)
  m2, _, _, _ = make_case("two_layer_syn", 2, H, W, "stress", "tc_syn2", gpu_ctx)
  a = m3.decompress(z, q, (H, W), return_float=True, return_yhat=True, return_bits=True)
  b = m2.decompress(z, q, (H, W), return_float=True, return_yhat=True, return_bits=True)
  for k in ("idx", "y_hat", "bits_y"):
    assert np.array_equal(a[k], b[k]), k
  err = np.abs(a["float"] - b["float"]).max()
  d = np.abs(a["image"].astype(int) - b["image"].astype(int))
  ref = oracle_decode(m3, wts, z[:1], q[:1], H, W)
  err_oracle = np.abs(b["float"][:1].astype(np.float64) - ref["recon"]).max()
  print(f"2-pass synthesis vs 3-pass: max-abs {err:.3e}, vs oracle {err_oracle:.3e}, u8 moved by one LSB on {(d > 0).mean():.4%}, max {d.max()}")
  # Measured on B200 (24 x 512x768 shape, stress weights): 4.5e-4 against the oracle, 1.3 % of the uint8 samples move by one
  # LSB.  Inside the 1e-3 tolerance, but only 2.2x -- not the >= 3x margin the default path keeps, and far above its 0.5 %
  # flip gate: this is why the mode is opt-in and the 3-pass product stays the default.
  assert err_oracle < 1e-3 and d.max() <= 1 and (d > 0).mean() < 0.05


# --------------------------------------------------------------------------------------------------
# intra-frame sharding: latent-row bands with recomputed halos (BASELINE configs[4])

@pytest.mark.parametrize("precision", ["tc", "fp32"])
@pytest.mark.parametrize("name,B,H,W,n_bands", [("bls2017", 1, 2160, 3840, 8), ("bls2017", 2, 200, 112, 3), ("two_layer_syn", 2, 512, 768, 2),
                                                ("two_layer_syn2:24", 1, 1200, 1200, 4), ("jpegl", 1, 300, 200, 4), ("mbt2018", 1, 256, 192, 2)])
def test_tiled_decode_equals_whole_frame_bit_for_bit(gpu_ctx, name, B, H, W, n_bands, precision):
  """Each band decoded from its own sub-tensors (band rows + the halo derived from the scatter formula, tiling.py) gives,
  on its rows, EXACTLY the bytes / rows / floats of the whole-frame decode: same K order per output element, zero padding
  only where the frame itself ends.  This is what lets rank r of N decode band r of a 4K frame with no communication."""
  if precision == "fp32" and H * W > 1200 * 1200:
    pytest.skip("the CUDA-core path is the reference GPU implementation: checked at the smaller sizes")
  model, wts, z, q = make_case(name, B, H, W, "stress", precision, gpu_ctx)
  kw = dict(return_float=True)
  if model.hyperprior:
    kw["return_yhat"] = True
  whole = model.decompress(z, q, (H, W), **kw) if model.hyperprior else model.decompress(q, (H, W), **kw)
  bands = model.band_plan((H, W), n_bands)
  assert len(bands) == n_bands and bands[0].rows[0] == 0 and bands[-1].rows[1] == H
  for b in bands:
    zb, qb = model.band_inputs(z, q, b)
    part = model.decompress_band(zb, qb, (H, W), b, **kw)
    r0, r1 = b.rows
    assert np.array_equal(part["image"], whole["image"][:, r0:r1]), (name, b.index, "image")
    assert np.array_equal(part["float"], whole["float"][:, r0:r1]), (name, b.index, "float")
    if model.hyperprior:
      c0, c1 = b.y_core
      assert np.array_equal(part["idx"], whole["idx"][:, c0:c1]), (name, b.index, "idx")
      assert np.array_equal(part["y_hat"], whole["y_hat"][:, c0:c1]), (name, b.index, "y_hat")
  stitched = model.decompress_tiled(z, q, (H, W), n_bands)
  assert np.array_equal(stitched["image"], whole["image"])
  # device-resident band inputs (what a rank of the multi-GPU run holds) give the same rows
  b = bands[-1]
  zb, qb = model.band_inputs(z, q, b)
  dev = model.decompress_band(gpu_ctx.to_device(zb) if zb is not None else None, gpu_ctx.to_device(qb), (H, W), b)
  assert np.array_equal(dev["image"], whole["image"][:, b.rows[0]:b.rows[1]])


# --------------------------------------------------------------------------------------------------
# the ONE collective of the path: NCCL called by libsntc (no PyTorch); single-rank communicator here, N ranks in bench.py

def test_nccl_group_of_one_rank(gpu_ctx, tmp_path):
  from shallow_ntc_b200 import parallel
  g = parallel.NcclGroup(gpu_ctx, 0, 1, str(tmp_path / "id"))
  sums = np.array([31.5, 48.25, 1.0e6, 2.0e3, 24.0])
  assert np.array_equal(g.allreduce_metrics(sums), sums) and np.array_equal(parallel.reduce_metric_sums(g, sums), sums)
  v = np.arange(100, dtype=np.float64) - 50.5
  assert np.array_equal(g.allreduce_sum(v), v) and np.array_equal(parallel.max_over_ranks(g, v), v)
  g.barrier()
  g.close()


# --------------------------------------------------------------------------------------------------
# small-batch decodes replayed as CUDA graphs (+ programmatic dependent launch): same bytes as the eager launches

@pytest.mark.parametrize("name", ["jpegl", "two_layer_syn", "two_layer_syn2:24", "bls2017"])
def test_cuda_graph_replay_of_small_batches_is_bit_identical(gpu_ctx, name, monkeypatch):
  H, W = 200, 300
  model, wts, z, q = make_case(name, 2, H, W, "stress", "tc", gpu_ctx)
  hyper = model.hyperprior
  args = (lambda zz, qq: (zz, qq, (H, W))) if hyper else (lambda zz, qq: (qq, (H, W)))
  host = model.decompress(*args(z, q), return_yhat=hyper)                       # host tensors: always eager
  z1 = z[:1] if hyper else None
  other = model.decompress(*args(z1, q[:1]))                                     # a second geometry in between
  dz, dq = (gpu_ctx.to_device(z) if hyper else None), gpu_ctx.to_device(q.astype(np.int16))
  out = dict(image=gpu_ctx.to_device(np.full((2, H, W, 3), 0x5A, np.uint8)))
  if hyper:
    out["idx"] = gpu_ctx.empty(q.shape, np.uint8)
    out["y_hat"] = gpu_ctx.empty(q.shape, np.float32)
  n0 = gpu_ctx.launch_counts["total"]
  per_call = []
  for it in range(6):                                                            # eager, capture, replay x4
    out["image"].fill_bytes(0x5A)
    gpu_ctx.sync()
    before = gpu_ctx.launch_counts["total"]
    got = model.decompress(*args(dz, dq), return_yhat=hyper, out=out)
    per_call.append(gpu_ctx.launch_counts["total"] - before)
    assert np.array_equal(got["image"].to_host(), host["image"]), it
    if hyper:
      assert np.array_equal(got["idx"].to_host(), host["idx"]) and np.array_equal(got["y_hat"].to_host(), host["y_hat"]), it
    if it == 3:                                                                  # another geometry on the same model re-uploads band tables
      again = model.decompress(*args(z1, q[:1]))
      assert np.array_equal(again["image"], other["image"])
  assert len(set(per_call)) == 1 and per_call[0] >= 2, per_call                 # replays account for the kernels they launch
  # graphs off: same bytes
  monkeypatch.setenv("SNTC_GRAPH", "0")
  got = model.decompress(*args(dz, dq), out=out)
  assert np.array_equal(got["image"].to_host(), host["image"])


@pytest.mark.parametrize("name,B,H,W", [("mbt2018", 3, 100, 150), ("two_layer_syn", 5, 97, 149), ("bls2017", 2, 90, 130)])
def test_repeated_decodes_and_shards_are_bit_identical_on_the_tensor_path(gpu_ctx, name, B, H, W):
  """Race / ordering check of the tensor-core kernels with shared-memory hand-offs between warps (col2im overlap-add epilogue with
  its named barriers, window-GEMM tail conversion stage, TMEM double buffering): 8 repeats of one decode are bit-identical, float
  outputs included, and a shard equals its slice of the batch (different tiles / work-item assignment, same sums)."""
  model, wts, z, q = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  first = model.decompress(z, q, (H, W), return_float=True)
  for _ in range(7):
    again = model.decompress(z, q, (H, W), return_float=True)
    assert np.array_equal(first["image"], again["image"]) and np.array_equal(first["float"], again["float"])
    if model.hyperprior:
      assert np.array_equal(first["idx"], again["idx"])
  zz = z[1:2] if z is not None else None
  part = model.decompress(zz, q[1:2], (H, W), return_float=True)
  assert np.array_equal(part["image"], first["image"][1:2]) and np.array_equal(part["float"], first["float"][1:2])


def test_vjp_is_deterministic_and_batch_independent(gpu_ctx):
  """Same for the decoder backward (s2d + tcgen05 backward layers + adjoint kernels): repeats are bit-identical and the gradient of
  image b does not depend on its batch neighbours."""
  from shallow_ntc_b200 import build_config
  model = build_config("two_layer_syn", precision="tc", ctx=gpu_ctx, vjp=True)
  wts = synthetic.make_weights(model.variable_shapes(), "stress", synthesis_cls="TwoLayerResSynthesis")
  model.load_weights(wts)
  zs, ys = model.latent_shapes(3, 64, 128)
  rng = np.random.default_rng(8)
  y = rng.standard_normal(ys).astype(np.float32)
  g = rng.standard_normal((3, 64, 128, 3)).astype(np.float32)
  a = model.synthesis_vjp(y, g)
  for _ in range(4):
    assert np.array_equal(a, model.synthesis_vjp(y, g))
  assert np.array_equal(a[1:2], model.synthesis_vjp(y[1:2], g[1:2]))
