"""not-gpu: pins the oracle (oracle/ntc_oracle.py) against (a) the reference's own structural
known-answers (results/*.csv, notebooks/get_flops.ipynb) and (b) the torch-definition fixtures in
tests/golden/ (see make_golden.py for why TensorFlow itself cannot be the source)."""
import os
import numpy as np
import pytest

from oracle import ntc_oracle as O
from shallow_ntc_b200 import build_config, synthetic

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "torch_definitions.npz"))


@pytest.mark.parametrize("k,s", [(13, 8), (5, 2), (18, 16), (3, 1), (6, 4), (16, 16), (4, 2), (7, 3)])
@pytest.mark.parametrize("gemm_form", [False, True])
def test_keras_conv2d_transpose_matches_gradient_definition(k, s, gemm_form):
  x, w, b, y = (GOLD[f"keras_k{k}s{s}_{n}"] for n in "xwby")
  got = O.keras_conv2d_transpose(x, w, b, s, np.float64, gemm_form)
  assert got.shape == y.shape == (2, 3 * s, 4 * s, 4)
  assert np.abs(got - y).max() < 1e-12


@pytest.mark.parametrize("k,s", [(5, 2), (9, 4), (3, 1)])
def test_tfc_signal_conv_up_matches_upsample_then_convolve(k, s):
  x, w, b, y = (GOLD[f"tfc_k{k}s{s}_{n}"] for n in "xwby")
  got = O.tfc_signal_conv2d_up(x, w, b, s)
  assert got.shape == y.shape and np.abs(got - y).max() < 1e-12


def _wts(prefix):
  return {k[len(prefix):]: GOLD[k] for k in GOLD.files if k.startswith(prefix)}


def test_two_layer_res_synthesis_golden():
  got = O.two_layer_res_synthesis(_wts("tlr_w:"), GOLD["tlr_yhat"])
  assert np.abs(got - GOLD["tlr_out"]).max() < 1e-11


def test_d2s_residual_matches_an_independent_torch_statement():
  """TwoLayerResSynthesis(res_type="d2s") (common/transforms.py:339-348) against torch: F.pixel_shuffle orders the channels
  (c, dy, dx) where tf.nn.depth_to_space orders them (dy, dx, c), hence the channel permutation; 1x1 convs as F.conv2d."""
  import torch
  import torch.nn.functional as F
  from shallow_ntc_b200 import transforms as T
  rng = np.random.default_rng(11)
  shapes = T.TwoLayerResSynthesis(res_type="d2s").variable_shapes(320)
  wts = synthetic.make_weights(shapes, "stress", synthesis_cls="TwoLayerResSynthesis")
  y = rng.standard_normal((2, 3, 5, 320)).astype(np.float32)
  got = O.d2s_residual(wts, y)
  assert got.shape == (2, 24, 40, 12)

  def tf_d2s(x):     # NCHW tensor holding TF's channel order (dy, dx, c) -> pixel_shuffle's (c, dy, dx)
    B, C4, h, w = x.shape
    c = C4 // 4
    return F.pixel_shuffle(x.reshape(B, 2, 2, c, h, w).permute(0, 3, 1, 2, 4, 5).reshape(B, C4, h, w), 2)
  x = torch.from_numpy(y.astype(np.float64)).permute(0, 3, 1, 2)
  for i in range(2):
    k = torch.from_numpy(wts[f"synthesis.res.conv_{i}.kernel"].astype(np.float64))[0, 0].T[:, :, None, None]   # [Cout, Cin, 1, 1]
    x = F.leaky_relu(F.conv2d(tf_d2s(x), k, torch.from_numpy(wts[f"synthesis.res.conv_{i}.bias"].astype(np.float64))), 0.2)
  want = tf_d2s(x).permute(0, 2, 3, 1).numpy()
  assert np.abs(got - want).max() < 1e-12
  # and the whole transform: out_conv(act(base_conv(z)) + res(z)) with the res branch swapped in
  full = O.two_layer_res_synthesis(wts, y, res_type="d2s")
  base = O.apply_activation(O.keras_conv2d_transpose(y, wts["synthesis.base_conv.kernel"], wts["synthesis.base_conv.bias"], 8), "igdn", wts, "synthesis.activation")
  assert np.abs(full - O.keras_conv2d_transpose(base + want, wts["synthesis.out_conv.kernel"], wts["synthesis.out_conv.bias"], 2)).max() < 1e-12


def test_hyper_synthesis_golden():
  got = O.hyper_synthesis(_wts("hs_w:"), GOLD["hs_z"])
  assert got.shape == (1, 8, 12, 8) and np.abs(got - GOLD["hs_out"]).max() < 1e-11


def test_bls_and_mbt_style_golden():
  w = _wts("bls_w:")
  assert np.abs(O.bls2017_synthesis(w, GOLD["bls_yhat"]) - GOLD["bls_out"]).max() < 1e-10
  # prefix of MBT2018Synthesis: two SignalConv + classic IGDN stages
  x = GOLD["bls_yhat"]
  for i in range(2):
    x = O.tfc_signal_conv2d_up(x, w[f"synthesis.layer_{i}.kernel"], w[f"synthesis.layer_{i}.bias"], 2)
    x = O.gdn_classic(x, w[f"synthesis.igdn_{i}.beta"], w[f"synthesis.igdn_{i}.gamma"], True)
  assert np.abs(x - GOLD["mbt2_out"]).max() < 1e-10


# ---- the reference's own known answers -------------------------------------------------------
def test_parameter_counts_match_results_all_params_csv():
  """results/all_params.csv:3-5 and notebooks/get_flops.ipynb cells 21, 26, 30."""
  assert build_config("jpegl")._synthesis.count_params(320) == 311043
  assert build_config("two_layer_syn")._synthesis.count_params(320) == 1299003
  assert build_config("two_layer_syn2:24")._synthesis.count_params(320) == 1300347
  assert build_config("two_layer_syn")._hyper_synthesis.count_params(320) == 9166240
  from shallow_ntc_b200.transforms import CNNSynthesis, JPEGLikeSynthesis
  assert CNNSynthesis(192, activation_type="leaky_relu").count_params(320) == 3394179
  assert CNNSynthesis(192, activation_type="igdn").count_params(320) == 3431235
  assert JPEGLikeSynthesis(kernel_size=16, strides=16).count_params(320) == 245763
  # the oracle counts what it is given the same way
  m = build_config("two_layer_syn")
  w = synthetic.make_weights(m.variable_shapes())
  assert O.count_params(w, "synthesis") == 1299003 and O.count_params(w, "hyper_synthesis") == 9166240


def test_flops_per_pixel_match_results_csv():
  """results/flops_per_pixel.csv / all_fpp.csv: TF profiler FLOPs at 512x768 = 2*MAC (+ bias adds)."""
  px = 512 * 768
  hs = O.convt_macs(8, 12, 5, 2, 320, 320) + O.convt_macs(16, 24, 5, 2, 320, 480) + O.convt_macs(32, 48, 3, 1, 480, 640)
  assert abs(2 * hs / px - 30354.6875) / 30354.6875 < 2e-3          # g_h column
  jpegl = O.convt_macs(32, 48, 18, 16, 320, 3)
  assert abs(2 * jpegl / px - 2433.0) / 2433.0 < 2e-3               # JPEG-like g
  assert abs(2 * jpegl - 956694528) / 956694528 < 2e-3              # notebook cell 23
  gdn = lambda c: 256 * 384 * c * c                                  # the 1x1 conv inside GDN1 (transforms.py:47)
  tlr = 2 * O.convt_macs(32, 48, 13, 8, 320, 12) + O.convt_macs(256, 384, 5, 2, 12, 3) + gdn(12)
  assert abs(2 * tlr / px - 10677.0) / 10677.0 < 3e-3               # 2-layer g
  tl24 = O.convt_macs(32, 48, 13, 8, 320, 24) + O.convt_macs(256, 384, 5, 2, 24, 3) + gdn(24)
  assert abs(2 * tl24 - 4462610184) / 4462610184 < 3e-3             # notebook cell 29


def test_shapes_and_zero_input_property():
  """notebook cells 12-14, 26: y [1,32,48,320], z [1,8,12,320] for 512x768; vis_syn_filters cells 28-29:
  synthesis(zeros[1,1,1,320]) -> [1,16,16,3] == bias."""
  m = build_config("jpegl")
  assert m.latent_shapes(1, 512, 768) == ((1, 8, 12, 320), (1, 32, 48, 320))
  assert build_config("two_layer_syn2").latent_shapes(1, 1200, 1200) == ((1, 19, 19, 320), (1, 76, 76, 320))
  assert build_config("bls2017").latent_shapes(1, 2160, 3840) == (None, (1, 135, 240, 256))
  w = synthetic.make_weights(m.variable_shapes(), "stress")
  out = O.jpeg_like_synthesis(w, np.zeros((1, 1, 1, 320)), strides=16)
  assert out.shape == (1, 16, 16, 3) and np.allclose(out, w["synthesis.conv.bias"])


def test_linearity_and_single_pixel_support():
  """vis_syn_filters.ipynb cells 36-44."""
  m = build_config("jpegl")
  w = synthetic.make_weights(m.variable_shapes(), "init")
  e = np.zeros((1, 2, 2, 320)); e[0, 0, 0, 5] = 1
  g0 = O.jpeg_like_synthesis(w, np.zeros_like(e)); g1 = O.jpeg_like_synthesis(w, e); g3 = O.jpeg_like_synthesis(w, 3 * e)
  assert np.abs((g3 - g0) - 3 * (g1 - g0)).max() < 1e-12
  nz = np.argwhere(np.abs(g1 - g0)[0].sum(-1) > 0)
  assert nz.max() <= 16     # 18x18 patch starting at -1: rows/cols 0..16 survive the SAME crop


def test_conv_transpose_alignment_matches_outputs_the_reference_recorded():
  """REFERENCE-HELD fixture (tests/golden/make_golden_notebook.py): the PNG outputs embedded in notebooks/vis_syn_filters.ipynb, real
  TF-2.10 runs of the trained jpegl model (JPEGLikeSynthesis(kernel_size=18, strides=16)).  Cell 44: one active latent pixel at (0, 0)
  of a 2 x 2 grid -> the response covers rows / columns 0..16 in all 100 recorded channels, on a constant background.  That is
  Conv2DTranspose(padding='SAME') with p = max(k - s, 0) // 2 = 1 (assumption A1 of the oracle); p = 0 would give 0..17, p = 2 0..15."""
  import os
  d = np.load(os.path.join(os.path.dirname(__file__), "golden", "notebook_jpegl_responses.npz"))
  c44, c41 = d["cell44"].astype(int), d["cell41"].astype(int)
  assert c44.shape == (100, 32, 32, 3) and c41.shape == (100, 16, 16, 3)
  bg = c44[:, -1:, -1:, :]
  mask_ref = (c44 != bg).any(-1)                                   # [100, 32, 32]
  assert not mask_ref[:, 17:, :].any() and not mask_ref[:, :, 17:].any()          # nothing beyond row / column 16
  assert mask_ref[:, 16, :17].any(-1).all() and mask_ref[:, :17, 16].any(-1).all()   # and row / column 16 IS touched, in every channel
  assert len(np.unique(bg.reshape(-1, 3), axis=0)) == 1             # the background is one pixel value (floats_to_pixels of the bias)
  # the 1 x 1 grid of cell 41 (response minus g0) shows the SAME kernel taps as the top-left 16 x 16 of cell 44: the two differ by a
  # per-channel constant (255 * bias) up to the rounding of each, wherever neither saturates
  diff = c44[:, :16, :16] - c41
  free = (c44[:, :16, :16] > 0) & (c44[:, :16, :16] < 255) & (c41 > 0) & (c41 < 255)
  for ch in range(3):
    v = np.unique(diff[..., ch][free[..., ch]])
    assert len(v) == 2 and v[1] - v[0] == 1, v
  # the oracle, same experiment, dense random kernel: identical support; the neighbouring paddings do not reproduce it
  m = build_config("jpegl")
  w = synthetic.make_weights(m.variable_shapes(), "stress")
  rng = np.random.default_rng(0)
  w["synthesis.conv.kernel"] = (rng.standard_normal(w["synthesis.conv.kernel"].shape) + 3.0).astype(np.float32)   # no zero taps
  e = np.zeros((1, 2, 2, 320)); e[0, 0, 0, 7] = 30.0
  bias = np.asarray(w["synthesis.conv.bias"], np.float64)
  mask = (O.jpeg_like_synthesis(w, e, strides=16)[0] != bias).any(-1)
  assert np.array_equal(mask, mask_ref.any(0)) and all(np.array_equal(mask, mr) or not (mr & ~mask).any() for mr in mask_ref)
  e1 = np.zeros((1, 1, 1, 320)); e1[0, 0, 0, 7] = 30.0
  assert np.array_equal(O.jpeg_like_synthesis(w, e, strides=16)[0, :16, :16], O.jpeg_like_synthesis(w, e1, strides=16)[0])   # same taps
  for p_wrong in (0, 2):
    out = O.conv_transpose_scatter(e, w["synthesis.conv.kernel"], w["synthesis.conv.bias"], 16, p_wrong)
    assert not np.array_equal((out[0] != bias).any(-1), mask_ref.any(0))


# ---- tiers, glue, epilogue ---------------------------------------------------------------------
def test_t1_float32_gemm_form_agrees_with_t0():
  m = build_config("two_layer_syn")
  cfg = m._transform_config["synthesis"]
  kw = {k: v for k, v in cfg.items() if k != "cls"}
  w = synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls=cfg["cls"])
  zs, ys = m.latent_shapes(1, 64, 128)
  z, q = synthetic.make_latents(zs, ys)
  t0 = O.mshyper_decode(w, cfg["cls"], z, q, 64, 128, kw)
  t1 = O.mshyper_decode(w, cfg["cls"], z, q, 64, 128, kw, dtype=np.float32, gemm_form=True)
  assert np.abs(t1["recon"] - t0["recon"]).max() < 1e-5 * max(1.0, np.abs(t0["recon"]).max())
  far = t0["idx_dist"] > 5e-5 * np.maximum(1, t0["i_c"])
  assert np.array_equal(t1["idx"][far], t0["idx"][far])
  assert t0["idx"].min() == 0 and t0["idx"].max() == 63, "stress weights must exercise both clamps"
  assert len(np.unique(t0["idx"])) == 64, "and every scale-table row"


def test_scale_index_rules():
  raw = np.log(np.array([1e-9, 0.49, 0.51, 1.49, 2.51, 62.4, 62.6, 63.5, 1e6]))
  i_c, idx, dist = O.scale_indexes(raw, "rint")
  assert idx.tolist() == [0, 0, 1, 1, 3, 62, 63, 63, 63]         # clamp to [0, 63], round to nearest
  assert np.allclose(dist[:3], [0.5 - 1e-9, 0.01, 0.01], atol=1e-6) and abs(dist[7] - 1.0) < 1e-9 and dist[-1] > 1e5   # above the clamp: distance to 62.5
  assert np.rint(np.array([0.5, 1.5, 2.5, 62.5])).tolist() == [0, 2, 2, 62]   # A3: ties to even
  _, idx_t, dist_t = O.scale_indexes(raw)                          # the default (A6): tf.cast(indexes, tf.int32) truncates
  assert idx_t.tolist() == [0, 0, 0, 1, 2, 62, 62, 63, 63]
  # distance to the nearest boundary: integers 1 .. 63 (values below 1 and above the clamp have ONE neighbour boundary)
  assert np.allclose(dist_t[:8], [1.0, 0.51, 0.49, 0.49, 0.49, 0.4, 0.4, 0.5], atol=1e-8) and dist_t[-1] > 1e5
  assert abs(O.scale_fn(0) - 0.11) < 1e-12 and abs(O.scale_fn(63) - 256.0) < 1e-9


def test_pixel_epilogue_and_metrics():
  x = np.array([[-0.6, -0.5, -0.5 + 0.4 / 255, -0.5 + 1.6 / 255, 0.0, 0.5, 0.7]], dtype=np.float32).reshape(1, 1, 7, 1)
  u8 = O.floats_to_pixels(x).ravel().tolist()
  assert u8 == [0, 0, 0, 2, 128, 255, 255]                          # saturate; the one exact tie 127.5 -> 128 (even)
  a = np.zeros((2, 4, 4, 3), np.uint8); b = a.copy(); b[0] += 10
  mse, psnr = O.mse_psnr(a, b)
  assert np.allclose(mse, [100.0, 0.0]) and abs(psnr[0] - 10 * np.log10(255 ** 2 / 100)) < 1e-9 and np.isinf(psnr[1])
  assert np.array_equal(O.dequantize(np.float32([3, -2]), np.float32([0.25, 0.5])), np.float32([3.25, -1.5]))
  assert np.array_equal(O.quantize_latent(np.array([0.5, 1.5, 2.5, -0.5])), [0, 2, 2, -0])


def test_synthetic_inputs_are_shard_invariant():
  m = build_config("two_layer_syn")
  zs, ys = m.latent_shapes(4, 64, 64)
  z, q = synthetic.make_latents(zs, ys)
  z2, q2 = synthetic.make_latents((2,) + zs[1:], (2,) + ys[1:], first_index=2)
  assert np.array_equal(z[2:], z2) and np.array_equal(q[2:], q2)
  assert np.array_equal(q, np.rint(q)) and np.abs(q).max() <= 127
  assert [synthetic.shard_range(24, r, 8) for r in range(8)] == [(3 * r, 3 * r + 3) for r in range(8)]
  assert [synthetic.shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]


# --------------------------------------------------------------------------------------------------
# rate term: known answers of the two entropy models

@pytest.mark.parametrize("i_c", [0.0, 3.7, 20.0, 41.5, 63.0])
def test_noisy_normal_is_a_probability_mass_function(i_c):
  """sum over all integers q of P(q | sigma) = 1 (the bins tile the real line), to float64 accuracy."""
  sig = float(O.scale_fn(i_c))
  K = int(12 * sig) + 40
  q = np.arange(-K, K + 1, dtype=np.float64)
  p = 2.0 ** (-O.noisy_normal_bits(q, np.full_like(q, i_c)))
  assert abs(p.sum() - 1.0) < 1e-12
  assert np.allclose(p, p[::-1], rtol=1e-12)          # symmetric in q


def test_noisy_normal_bits_matches_direct_difference_and_stays_finite_in_the_tails():
  from scipy.special import ndtr
  rng = np.random.default_rng(0)
  i_c = rng.uniform(0, 63, size=2000)
  sig = O.scale_fn(i_c)
  q = np.rint(rng.normal(0, 1, size=2000) * sig)
  direct = -np.log2(ndtr((q + .5) / sig) - ndtr((q - .5) / sig))
  ok = np.isfinite(direct) & (direct < 40)
  assert np.allclose(O.noisy_normal_bits(q, i_c)[ok], direct[ok], rtol=1e-9)
  far = O.noisy_normal_bits(np.array([127.0, -127.0, 50.0]), np.array([0.0, 0.0, 1.0]))   # sigma ~ 0.11: z ~ 1150
  assert np.all(np.isfinite(far)) and far[0] == far[1] and far[0] > 9e5
  # leading term of the tail: -log2 Phi(-a) ~ a^2 / (2 ln 2)
  a = (127 - .5) / 0.11
  assert abs(far[0] / (a * a / (2 * np.log(2))) - 1) < 1e-4


@pytest.mark.parametrize("kind", ["init", "stress"])
def test_deep_factorized_prior_is_a_monotone_cdf_and_a_pmf(kind):
  m = build_config("two_layer_syn", prior=True)
  wts = synthetic.make_weights(m.variable_shapes(), kind, synthesis_cls="TwoLayerResSynthesis")
  C = 320
  z = np.arange(-400, 401, dtype=np.float64)[:, None] * np.ones((1, C))
  lg = O.deep_factorized_logits(z, wts)
  assert lg.shape == z.shape and np.all(np.diff(lg, axis=0) > 0)      # softplus matrices + |tanh factor| < 1 -> increasing
  p = 2.0 ** (-O.deep_factorized_bits(z, wts))
  assert np.all(np.abs(p.sum(0) - 1.0) < 1e-6)                        # mass outside +-400 is negligible (init_scale 10)
  # parameter count of tfc.NoisyDeepFactorized(batch_shape=(320,)), num_filters (3,3,3): 3+3+3 + 9+3+3 + 9+3+3 + 3+1 = 43 per channel
  assert O.count_params(wts, "prior") == 43 * C


def test_rate_bits_shapes_and_additivity():
  m = build_config("two_layer_syn", prior=True)
  wts = synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls="TwoLayerResSynthesis")
  zs, ys = m.latent_shapes(3, 64, 64)
  z, q = synthetic.make_latents(zs, ys)
  out = O.mshyper_decode(wts, "TwoLayerResSynthesis", z, q, 64, 64, dict(channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5)))
  assert out["bits_y"].shape == (3,) and out["bits_z"].shape == (3,)
  one = O.mshyper_decode(wts, "TwoLayerResSynthesis", z[1:2], q[1:2], 64, 64, dict(channels=(12, 3), strides=(8, 2), kernel_sizes=(13, 5)))
  assert np.allclose(one["bits_y"], out["bits_y"][1:2], rtol=1e-12) and np.allclose(one["bits_z"], out["bits_z"][1:2], rtol=1e-12)


def test_msssim_oracle_matches_the_torch_statement_of_tf_image_ssim_multiscale():
  """A9: oracle.msssim (separable 1-D window, numpy) against an independent torch statement of tf.image.ssim /
  ssim_multiscale (full 2-D softmax window as a grouped conv, replicate-pad + avg_pool2d), tests/golden/make_golden_msssim.py:
  odd sizes (end padding at three scales), the single-scale branch (< 160 px), identical images -> 1, and the size rule."""
  from shallow_ntc_b200.eval_lib import msssim_defined
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", "msssim_torch.npz"))
  for name in ("multi_odd", "small", "multi_even"):
    val, db = O.msssim(g[name + "_a"], g[name + "_b"])
    assert np.abs(val - g[name + "_val"]).max() < 1e-12, name
    assert np.allclose(db, -10 * np.log10(1 - val))
  a = g["multi_even_a"]
  assert np.allclose(O.msssim(a, a)[0], 1.0)
  assert msssim_defined(512, 768) and msssim_defined(96, 120) and msssim_defined(161, 176) and not msssim_defined(100, 200) and not msssim_defined(8, 100)
  with pytest.raises(ValueError):
    O.msssim(np.zeros((1, 100, 200, 3), np.uint8), np.zeros((1, 100, 200, 3), np.uint8))
