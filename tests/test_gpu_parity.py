"""-m gpu: libsntc (through the C ABI / ctypes) against the float64 oracle on the same seeded inputs."""
import numpy as np
import pytest

from shallow_ntc_b200 import synthetic
from helpers import make_case, oracle_decode, check_against_oracle, PSNR_TOL

pytestmark = pytest.mark.gpu

SMALL = [  # (config, B, H, W)
  ("jpegl", 2, 128, 192),
  ("two_layer_syn", 2, 128, 192),
  ("two_layer_syn2", 1, 100, 150),       # C1 = 12, needs reflect-pad geometry (Hp=128, Wp=192) + crop
  ("two_layer_syn2:24", 1, 128, 128),
  ("two_layer_syn2:48", 1, 64, 128),
  ("mbt2018", 1, 64, 128),
  ("bls2017", 2, 48, 80),
]


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("kind", ["init", "stress"])
@pytest.mark.parametrize("name,B,H,W", SMALL)
def test_decode_matches_oracle(gpu_ctx, name, B, H, W, kind, precision):
  """precision 'fp32' = CUDA-core FFMA kernels; 'tc' = tcgen05 split-fp16 (3-pass) kernels."""
  model, wts, z, q = make_case(name, B, H, W, kind, precision, gpu_ctx)
  ref = oracle_decode(model, wts, z, q, H, W)
  orig = synthetic.make_original(ref["recon_u8"])
  ref_mse, ref_psnr = __import__("oracle.ntc_oracle", fromlist=["x"]).mse_psnr(orig, ref["recon_u8"])
  n0 = gpu_ctx.launch_counts
  got = model.decompress(z, q, (H, W), return_float=True, return_yhat=True, original=orig)
  n = {k: v - n0[k] for k, v in gpu_ctx.launch_counts.items()}
  rep = check_against_oracle(got, ref, hyper=model.hyperprior, precision=precision)
  assert np.all(np.abs(got["psnr"] - ref_psnr) < PSNR_TOL), (got["psnr"], ref_psnr)
  # which kernels served the call: a 'tc' model must not fall back to FFMA silently, an 'fp32' model never touches tcgen05
  if precision == "tc":
    assert n["band_tc"] >= 1 and n["band_f32"] == 0 and n["final_f32"] == 0, n
  else:
    assert n["band_tc"] == 0 and n["tail_mma"] == 0 and n["tail_tc"] == 0 and n["band_f32"] >= 1, n
  print(name, kind, precision, rep, n)


def test_yhat_is_bit_exact_single_add(gpu_ctx):
  """y_hat = q + mu must be exactly fl32(q + mu_gpu): decode twice with q and with q = 0 (-> mu)."""
  model, wts, z, q = make_case("two_layer_syn", 1, 64, 64, "stress", "fp32", gpu_ctx)
  mu = model.decompress(z, np.zeros_like(q), (64, 64), return_yhat=True)["y_hat"]
  yh = model.decompress(z, q, (64, 64), return_yhat=True)["y_hat"]
  assert np.array_equal(yh, (q + mu).astype(np.float32))


def test_hyper_synthesis_and_synthesis_transform_calls(gpu_ctx):
  """The Keras-layer call convention: hyper_synthesis(z_hat) and synthesis(y_hat) on their own."""
  from oracle import ntc_oracle as O
  model, wts, z, q = make_case("two_layer_syn", 1, 64, 128, "stress", "fp32", gpu_ctx)
  hs = model.hyper_synthesis(z)
  ref = O.hyper_synthesis(wts, z)
  assert hs.shape == ref.shape and np.abs(hs - ref).max() < 1e-4
  y_hat = (q + hs[..., :320]).astype(np.float32)
  x = model.synthesis(y_hat)
  refx = O.two_layer_res_synthesis(wts, y_hat.astype(np.float64))
  assert x.shape == refx.shape == (1, 64, 128, 3) and np.abs(x - refx).max() < 1e-3


def test_int8_int16_latents_match_float(gpu_ctx):
  model, wts, z, q = make_case("two_layer_syn", 2, 64, 64, "stress", "fp32", gpu_ctx)
  a = model.decompress(z, q, (64, 64))
  b = model.decompress(z, q.astype(np.int16), (64, 64))
  c = model.decompress(z, q.astype(np.int8), (64, 64))
  for k in ("image", "idx"):
    assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k])


def test_device_resident_zero_copy_matches_host(gpu_ctx):
  model, wts, z, q = make_case("two_layer_syn", 2, 64, 128, "stress", "fp32", gpu_ctx)
  host = model.decompress(z, q, (64, 128))
  dz, dq = gpu_ctx.to_device(z), gpu_ctx.to_device(q)
  dev = model.decompress(dz, dq, (64, 128))
  assert np.array_equal(dev["image"].to_host(), host["image"])
  assert np.array_equal(dev["idx"].to_host(), host["idx"])


def test_batch_independence_and_determinism(gpu_ctx):
  """Images are independent units: decoding a shard equals the slice of the full batch."""
  model, wts, z, q = make_case("two_layer_syn", 4, 64, 64, "stress", "fp32", gpu_ctx)
  full = model.decompress(z, q, (64, 64))
  again = model.decompress(z, q, (64, 64))
  assert np.array_equal(full["image"], again["image"]) and np.array_equal(full["idx"], again["idx"])
  part = model.decompress(z[1:3], q[1:3], (64, 64))
  assert np.array_equal(part["image"], full["image"][1:3]) and np.array_equal(part["idx"], full["idx"][1:3])


def test_zero_latents_give_bias_image(gpu_ctx):
  """vis_syn_filters.ipynb cells 28-29: synthesis(zeros) equals the bias broadcast (JPEG-like)."""
  model, wts, z, q = make_case("jpegl", 1, 64, 64, "stress", "fp32", gpu_ctx)
  x = model.synthesis(np.zeros((1, 4, 4, 320), np.float32))
  assert np.allclose(x, wts["synthesis.conv.bias"].reshape(1, 1, 1, 3), atol=1e-7)


def test_jpegl_linearity(gpu_ctx):
  """vis_syn_filters.ipynb cells 36-41: g(k e_i) - g(0) = k (g(e_i) - g(0)) for the one-layer synthesis."""
  model, wts, z, q = make_case("jpegl", 1, 64, 64, "init", "fp32", gpu_ctx)
  e = np.zeros((1, 4, 4, 320), np.float32)
  e[0, 1, 2, 17] = 1.0
  g0 = model.synthesis(np.zeros_like(e))
  g1 = model.synthesis(e)
  g5 = model.synthesis(5 * e)
  assert np.abs((g5 - g0) - 5 * (g1 - g0)).max() < 1e-5
  # support of one latent pixel: an 18x18 patch at (16*1 - 1, 16*2 - 1)
  nz = np.argwhere(np.abs(g1 - g0)[0].sum(-1) > 0)
  assert nz[:, 0].min() >= 15 and nz[:, 0].max() <= 32 and nz[:, 1].min() >= 31 and nz[:, 1].max() <= 48


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_single_pixel_response_matches_what_the_reference_recorded(gpu_ctx, precision):
  """REFERENCE-HELD fixture (tests/golden/notebook_jpegl_responses.npz, made from the PNG outputs embedded in
  notebooks/vis_syn_filters.ipynb cell 44: real TF-2.10 runs of the trained jpegl model): one active latent pixel at (0, 0) of a 2 x 2
  grid gives a response on rows / columns 0..16 of the 32 x 32 image and one constant pixel value elsewhere.  The CUDA path, same
  experiment through floats_to_pixels: identical support, constant background."""
  import os
  from shallow_ntc_b200 import synthetic
  ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "notebook_jpegl_responses.npz"))["cell44"].astype(int)
  mask_ref = (ref != ref[:, -1:, -1:, :]).any(-1).any(0)
  model, wts, z, q = make_case("jpegl", 1, 32, 32, "stress", precision, gpu_ctx)
  rng = np.random.default_rng(0)
  wts = dict(wts)
  wts["synthesis.conv.kernel"] = (0.002 * (rng.standard_normal(wts["synthesis.conv.kernel"].shape) + 3.0)).astype(np.float32)   # no zero taps
  model.load_weights(wts)
  e = np.zeros((1, 2, 2, 320), np.float32)
  e[0, 0, 0, 7] = 30.0
  x = model.synthesis(e)[0]
  px = np.clip(np.rint((x + 0.5) * 255.0), 0, 255).astype(int)                      # data_lib.floats_to_pixels
  bg = px[-1, -1]
  assert np.array_equal((px != bg).any(-1), mask_ref)
  assert np.array_equal(bg, np.clip(np.rint((wts["synthesis.conv.bias"] + 0.5) * 255.0), 0, 255).astype(int))


def test_argument_errors_are_loud(gpu_ctx):
  from shallow_ntc_b200 import SntcError
  model, wts, z, q = make_case("two_layer_syn", 1, 64, 64, "init", "fp32", gpu_ctx)
  with pytest.raises(SntcError):
    model.decompress(z, q[..., :100].copy(), (64, 64))          # wrong channel count
  with pytest.raises(SntcError):
    model.decompress(z, q, (65, 64))                              # image larger than the latent grid
  with pytest.raises(SntcError):
    model.decompress(z[:, :0], q, (64, 64))                       # z / y geometry mismatch
  empty = model.decompress(z[:0], q[:0], (64, 64))                # empty batch is a no-op
  assert empty["image"].shape == (0, 64, 64, 3)


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_full_size_config2_properties(gpu_ctx, precision):
  """BASELINE config 2 (two_layer_syn, 24 x 512x768): oracle on 2 of the 24 images + size-independent
  properties (shard independence, exact integer SSD / PSNR against a host recomputation)."""
  B, H, W = 24, 512, 768
  model, wts, z, q = make_case("two_layer_syn", B, H, W, "stress", precision, gpu_ctx)
  got = model.decompress(z, q, (H, W), return_float=True, return_yhat=True)
  for b in (0, 23):
    ref = oracle_decode(model, wts, z[b:b + 1], q[b:b + 1], H, W)
    sub = {k: v[b:b + 1] for k, v in got.items()}
    print(b, check_against_oracle(sub, ref, precision=precision))
  shard = model.decompress(z[5:9], q[5:9], (H, W))
  assert np.array_equal(shard["image"], got["image"][5:9]) and np.array_equal(shard["idx"], got["idx"][5:9])
  orig = synthetic.make_original(got["image"])
  m = model.decompress(z, q, (H, W), original=orig)
  ssd = ((m["image"].astype(np.int64) - orig.astype(np.int64)) ** 2).reshape(B, -1).sum(1)
  assert np.array_equal(m["ssd"].astype(np.int64), ssd)
  assert np.array_equal(m["image"], got["image"])


@pytest.mark.parametrize("B,H,W", [(2, 128, 192), (1, 97, 149), (3, 100, 150), (1, 512, 768)])
def test_jpegl_byte_run_epilogue_writes_the_scalar_paths_bytes(gpu_ctx, monkeypatch, B, H, W):
  """JPEG-like synthesis on the tensor cores: when the uint8 image is the only destination the epilogue writes byte runs
  (32-bit stores re-aligned with funnel shifts, tc_epi_rgb_chunk); asking for the float image too takes the scalar
  epilogue.  Same bytes for crops to odd sizes (runs cut at W*3 not a multiple of 4, rows cut at H), for use_offset-free
  and ragged tiles, and with the run path disabled (SNTC_TC_RGB_RUNS=0)."""
  model, wts, z, q = make_case("jpegl", B, H, W, "stress", "tc", gpu_ctx)
  scalar = model.decompress(z, q, (H, W), return_float=True, return_yhat=True)
  if H * W <= 128 * 192:
    print(check_against_oracle(scalar, oracle_decode(model, wts, z, q, H, W), precision="tc"))
  # device-resident canvas pre-filled with a sentinel (in place, no staging): every byte must be written, none outside
  canvas = gpu_ctx.to_device(np.full((B + 1, H, W, 3), 0xA5, np.uint8))
  out_img = gpu_ctx.to_device(np.full((B, H, W, 3), 0xA5, np.uint8))
  fast = model.decompress(gpu_ctx.to_device(z), gpu_ctx.to_device(q), (H, W), out=dict(image=out_img))
  assert np.array_equal(fast["image"].to_host(), scalar["image"])
  del canvas
  monkeypatch.setenv("SNTC_TC_RGB_RUNS", "0")
  base, _, _, _ = make_case("jpegl", B, H, W, "stress", "tc", gpu_ctx)
  assert np.array_equal(base.decompress(z, q, (H, W))["image"], scalar["image"])


@pytest.mark.parametrize("name", ["two_layer_syn", "jpegl", "two_layer_syn2:48"])
def test_extreme_image_sizes(gpu_ctx, name):
  """Smallest and awkward geometries (a single pixel, one latent cell, sizes just off the 64-px padding grid, a large
  image whose sides are not multiples of anything): the tensor-core path agrees with the fp32 CUDA-core path (each pinned
  to the oracle on the other shapes), its uint8-only fast path writes the same bytes, and rows are never off by more
  than one.  tools/edge_check.py runs the longer list (up to 2048 x 2048)."""
  tc, _, _, _ = make_case(name, 1, 64, 64, "stress", "tc", gpu_ctx)
  ff, _, _, _ = make_case(name, 1, 64, 64, "stress", "fp32", gpu_ctx)
  for B, H, W in ((1, 1, 1), (1, 16, 16), (3, 63, 65), (2, 257, 511), (1, 1201, 999)):
    zs, ys = tc.latent_shapes(B, H, W)
    z, q = synthetic.make_latents(zs, ys)
    a = tc.decompress(z, q, (H, W), return_float=True)
    b = ff.decompress(z, q, (H, W), return_float=True)
    assert np.abs(a["float"] - b["float"]).max() < 1e-4, (B, H, W)
    assert np.abs(a["image"].astype(int) - b["image"].astype(int)).max() <= 1
    assert np.abs(a["idx"].astype(int) - b["idx"].astype(int)).max() <= 1 and (a["idx"] != b["idx"]).mean() < 2e-2
    assert np.array_equal(tc.decompress(z, q, (H, W))["image"], a["image"])


def test_large_batch_equals_its_shards(gpu_ctx):
  """Four bench-sized shards in ONE call (96 x 512x768: 3.6 GB of activations in flight, > 2^31 bytes of t) give exactly
  the bytes of the four 24-image calls -- index arithmetic, work-item scheduling and the persistent kernels' tile loops
  do not depend on the batch a tile belongs to (weak scaling moves shards between GPUs, never changes them)."""
  B, H, W = 96, 512, 768
  model, wts, z, q = make_case("two_layer_syn", B, H, W, "stress", "tc", gpu_ctx)
  q8 = q.astype(np.int8)
  full = model.decompress(z, q8, (H, W))
  for s in range(4):
    part = model.decompress(z[24 * s:24 * s + 24], q8[24 * s:24 * s + 24], (H, W))
    assert np.array_equal(part["image"], full["image"][24 * s:24 * s + 24]) and np.array_equal(part["idx"], full["idx"][24 * s:24 * s + 24])


@pytest.mark.parametrize("name,B,H,W", [("two_layer_syn", 2, 128, 192), ("jpegl", 1, 64, 128)])
def test_tc_path_matches_fp32_path(gpu_ctx, name, B, H, W):
  """The two GPU implementations agree with each other far inside the tolerance."""
  m32, wts, z, q = make_case(name, B, H, W, "stress", "fp32", gpu_ctx)
  mtc, _, _, _ = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  a = m32.decompress(z, q, (H, W), return_float=True, return_yhat=True)
  b = mtc.decompress(z, q, (H, W), return_float=True, return_yhat=True)
  assert np.abs(a["float"] - b["float"]).max() < 1e-4
  assert np.abs(a["y_hat"] - b["y_hat"]).max() < 1e-3
  assert (a["idx"] != b["idx"]).mean() < 1e-3
  l0 = gpu_ctx.launch_count
  mtc.decompress(z, q, (H, W))
  assert gpu_ctx.launch_count > l0, "the tensor-core decode must launch libsntc kernels"


def test_tc_hyper_transform_call(gpu_ctx):
  from oracle import ntc_oracle as O
  model, wts, z, q = make_case("two_layer_syn", 1, 64, 128, "stress", "tc", gpu_ctx)
  hs = model.hyper_synthesis(z)
  ref = O.hyper_synthesis(wts, z)
  assert np.abs(hs - ref).max() < 2e-4


@pytest.mark.parametrize("q_dtype", [np.float32, np.int16, np.int8])
@pytest.mark.parametrize("packed", [False, True])
def test_streaming_pipeline_matches_direct_decode(gpu_ctx, q_dtype, packed):
  """DecodePipeline (H2D / decode / D2H on three streams, double-buffered) returns exactly what the synchronous call
  returns, for every batch, in order -- with caller-owned pinned arrays (one copy per array) and with the pipeline's own
  packed page-locked slots (one copy per direction), with and without the idx download."""
  from shallow_ntc_b200 import DecodePipeline
  B, H, W, N = 2, 64, 128, 5
  model, wts, z, q = make_case("two_layer_syn", B * N, H, W, "stress", "tc", gpu_ctx)
  refs = [model.decompress(z[i * B:(i + 1) * B], q[i * B:(i + 1) * B], (H, W)) for i in range(N)]
  zs, ys = model.latent_shapes(B, H, W)
  for return_idx in (True, False):
    pipe = DecodePipeline(model, B, (H, W), q_dtype=q_dtype, depth=2, return_idx=return_idx, host_slots=N, write_combined=packed)
    if packed:
      slots = [pipe.host_slot(i) for i in range(N)]
      for i, h in enumerate(slots):
        h["z"][...] = z[i * B:(i + 1) * B]
        h["q"][...] = q[i * B:(i + 1) * B].astype(q_dtype)
        h["image"][...] = 0xA5
        h["idx"][...] = 0xA5
      tickets = [pipe.submit(h["z"], h["q"], h["image"], h["idx"]) for h in slots]
      outs = slots
    else:
      outs = [dict(image=gpu_ctx.pinned_empty((B, H, W, 3), np.uint8), idx=gpu_ctx.pinned_empty(ys, np.uint8)) for _ in range(N)]
      for o in outs:
        o["idx"][...] = 0xA5
      pins = [(gpu_ctx.pinned_like(z[i * B:(i + 1) * B]), gpu_ctx.pinned_like(q[i * B:(i + 1) * B].astype(q_dtype))) for i in range(N)]
      tickets = [pipe.submit(pz, pq, o["image"], o["idx"]) for (pz, pq), o in zip(pins, outs)]
    pipe.wait(tickets[-1])
    pipe.drain()
    for i in range(N):
      assert np.array_equal(outs[i]["image"], refs[i]["image"]), (i, return_idx)
      if return_idx:
        assert np.array_equal(outs[i]["idx"], refs[i]["idx"]), i
      else:
        assert np.all(outs[i]["idx"] == 0xA5), "idx must not be downloaded unless asked for"
    # the rows stay available on the device either way
    last = pipe.slots[(N - 1) % 2]["idx"].to_host()
    assert np.array_equal(last, refs[N - 1]["idx"])


def test_streaming_pipeline_factorized_model(gpu_ctx):
  from shallow_ntc_b200 import DecodePipeline
  B, H, W = 2, 48, 80
  model, wts, z, q = make_case("bls2017", B * 3, H, W, "stress", "tc", gpu_ctx)
  pipe = DecodePipeline(model, B, (H, W), q_dtype=np.int8, depth=2)
  for i in range(3):
    h = pipe.host_slot(i)
    h["q"][...] = q[i * B:(i + 1) * B].astype(np.int8)
    t = pipe.submit(None, h["q"], h["image"])
    pipe.wait(t)
    ref = model.decompress(q[i * B:(i + 1) * B], (H, W))
    assert np.array_equal(h["image"], ref["image"])
  pipe.drain()


@pytest.mark.parametrize("name,B,H,W", [("two_layer_syn", 2, 128, 192), ("two_layer_syn2", 1, 100, 150), ("two_layer_syn2:24", 1, 128, 128)])
def test_tensor_core_tail_kernel_matches_oracle(gpu_ctx, monkeypatch, name, B, H, W):
  """The tcgen05 tail (halo-reuse A patch, un-swizzled descriptors, [w_hi | w_lo] in N; automatic for C1 > 16, forced
  here with SNTC_TC_TAIL=1 for C1 = 12 too) meets the same gates as the other tails (SNTC_TC_TAIL=0: FFMA / warp-MMA)."""
  monkeypatch.setenv("SNTC_TC_TAIL", "1")
  model, wts, z, q = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  ref = oracle_decode(model, wts, z, q, H, W)
  got = model.decompress(z, q, (H, W), return_float=True, return_yhat=True)
  print(name, check_against_oracle(got, ref, precision="tc"))
  monkeypatch.setenv("SNTC_TC_TAIL", "0")
  base, _, _, _ = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  ref2 = base.decompress(z, q, (H, W), return_float=True)
  assert np.abs(ref2["float"] - got["float"]).max() < 1e-4


@pytest.mark.parametrize("name,B,H,W", [("two_layer_syn", 2, 128, 192), ("two_layer_syn2", 1, 100, 150), ("two_layer_syn2", 2, 97, 149),
                                        ("two_layer_syn", 1, 512, 768)])
def test_warp_mma_tail_kernel(gpu_ctx, monkeypatch, name, B, H, W):
  """The warp-MMA tail (sntc_kernels_tail_mma.cuh, default for C1 = 12 on the tensor-core path): same gates against the
  oracle as every other kernel; its uint8-only fast path writes exactly the bytes of its generic path (odd widths,
  crops and ragged tiles included); and it agrees with the FFMA tail (SNTC_TAIL_MMA=0) far inside the tolerance."""
  monkeypatch.setenv("SNTC_TAIL_TZ", "0")     # the window-GEMM tcgen05 tail is the default for C1 = 12 since round 2
  model, wts, z, q = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  n0 = gpu_ctx.launch_counts
  got = model.decompress(z, q, (H, W), return_float=True, return_yhat=True)
  assert gpu_ctx.launch_counts["tail_mma"] == n0["tail_mma"] + 1
  if H * W <= 128 * 192:
    ref = oracle_decode(model, wts, z, q, H, W)
    print(name, check_against_oracle(got, ref, precision="tc"))
  fast = model.decompress(z, q, (H, W))
  assert np.array_equal(fast["image"], got["image"])
  monkeypatch.setenv("SNTC_TAIL_MMA", "0")
  base, _, _, _ = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  ffma = base.decompress(z, q, (H, W), return_float=True)
  assert np.abs(ffma["float"] - got["float"]).max() < 2e-6
  d = np.abs(ffma["image"].astype(int) - got["image"].astype(int))
  assert d.max() <= 1 and (d > 0).mean() < 1e-3


@pytest.mark.parametrize("name,B,H,W", [("two_layer_syn", 2, 128, 192), ("two_layer_syn2", 1, 100, 150), ("two_layer_syn2", 2, 97, 149),
                                        ("two_layer_syn", 3, 72, 520), ("two_layer_syn", 1, 512, 768), ("two_layer_syn2", 1, 1200, 1200)])
def test_window_gemm_tail_kernel(gpu_ctx, monkeypatch, name, B, H, W):
  """The window-GEMM tcgen05 tail (sntc_kernels_tail_tz.cuh, default for C1 = 12 on the tensor-core path): one GEMM row =
  4 t-pixels, banded weights, [w_hi | w_lo] in N.  Same gates against the oracle as every other kernel; its 8-byte-store
  fast path writes exactly the bytes of its generic path (odd widths, crops, ragged tiles, several tiles per row); it
  agrees with the warp-MMA tail (SNTC_TAIL_TZ=0) to fp32 rounding; and the launch counters prove which kernel ran."""
  model, wts, z, q = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  n0 = gpu_ctx.launch_counts
  got = model.decompress(z, q, (H, W), return_float=True, return_yhat=True)
  n = {k: v - n0[k] for k, v in gpu_ctx.launch_counts.items()}
  assert n["tail_tc"] == 1 and n["tail_mma"] == 0 and n["final_f32"] == 0 and n["band_f32"] == 0, n
  if H * W <= 128 * 192:
    ref = oracle_decode(model, wts, z, q, H, W)
    print(name, check_against_oracle(got, ref, precision="tc"))
  fast = model.decompress(z, q, (H, W))
  assert np.array_equal(fast["image"], got["image"])
  monkeypatch.setenv("SNTC_TAIL_TZ", "0")
  base, _, _, _ = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  mma = base.decompress(z, q, (H, W), return_float=True)
  assert np.array_equal(mma["idx"], got["idx"])
  assert np.abs(mma["float"] - got["float"]).max() < 2e-6
  d = np.abs(mma["image"].astype(int) - got["image"].astype(int))
  assert d.max() <= 1 and (d > 0).mean() < 1e-3


@pytest.mark.parametrize("name,B,H,W,label", [("two_layer_syn2:48", 2, 97, 149, "synthesis.conv1+activation"),
                                              ("two_layer_syn:48", 1, 128, 192, "synthesis.base_conv+activation"),
                                              ("two_layer_syn:24", 2, 100, 150, "synthesis.base_conv+activation"),
                                              ("two_layer_syn2:24", 1, 97, 149, "synthesis.conv1+activation")])
def test_wide_hidden_layers_take_the_fused_tensor_core_path(gpu_ctx, name, B, H, W, label):
  """two_layer_syn2's sweep widths (mshyper/configs/two_layer_syn2.py:87-89: hidden_channels 24, 48) and the residual
  variant at those widths: IGDN1 (+ residual) runs in the layer-1 GEMM epilogue (C1 = 48: gamma / beta in shared
  memory) and the tail conv on tcgen05 (automatic for C1 > 16) -- checked against the oracle with crop + ragged tiles,
  and against the fp32 CUDA-core path; the layer profile proves the fused kernels are the ones that ran."""
  model, wts, z, q = make_case(name, B, H, W, "stress", "tc", gpu_ctx)
  ref = oracle_decode(model, wts, z, q, H, W)
  model._ensure_native()
  model.profile_layers(True)
  got = model.decompress(z, q, (H, W), return_float=True, return_yhat=True)
  model.profile_layers(False)
  labels = set(model.layer_profile())
  assert label in labels and not any("activation+res" in l or l.endswith(".activation") for l in labels), labels
  print(name, check_against_oracle(got, ref, precision="tc"))
  base, _, _, _ = make_case(name, B, H, W, "stress", "fp32", gpu_ctx)
  f32 = base.decompress(z, q, (H, W), return_float=True)
  assert np.abs(f32["float"] - got["float"]).max() < 1e-4
  fast = model.decompress(z, q, (H, W))                     # uint8-only outputs: same bytes
  assert np.array_equal(fast["image"], got["image"])


# --------------------------------------------------------------------------------------------------
# rate term

def _rate_case(name, B, H, W, precision, ctx):
  from shallow_ntc_b200 import build_config
  model = build_config(name, precision=precision, ctx=ctx, prior=True)
  cls = model._transform_config["synthesis"]["cls"]
  wts = synthetic.make_weights(model.variable_shapes(), "stress", synthesis_cls=cls)
  model.load_weights(wts)
  zs, ys = model.latent_shapes(B, H, W)
  z, q = synthetic.make_latents(zs, ys)
  return model, wts, z, q


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("name,B,H,W", [("two_layer_syn", 3, 128, 192), ("jpegl", 2, 64, 128), ("two_layer_syn", 5, 64, 96)])   # last: m-tiles span images
def test_rate_term_matches_oracle(gpu_ctx, name, B, H, W, precision):
  """bits_y / bits_z of sntc_decode_rd against the float64 oracle.  bits_y inherits the accuracy of raw sigma through
  sigma = SCALE_FN(clamp(exp(raw))): d(bits)/bits ~ 2 * 0.123 * i_c * d(raw) on tail symbols, hence the looser bound on
  the split-fp16 path; the rate arithmetic itself is checked at 2e-5 against the GPU's own raw sigma."""
  from oracle import ntc_oracle as O
  model, wts, z, q = _rate_case(name, B, H, W, precision, gpu_ctx)
  got = model.decompress(z, q, (H, W), return_bits=True)
  ref = oracle_decode(model, wts, z, q, H, W)
  tol_y = 3e-4 if precision == "fp32" else 3e-3
  assert np.all(np.abs(got["bits_y"] / ref["bits_y"] - 1) < tol_y), (got["bits_y"], ref["bits_y"])
  assert np.all(np.abs(got["bits_z"] / ref["bits_z"] - 1) < 2e-5), (got["bits_z"], ref["bits_z"])
  assert np.allclose(got["bpp"], (got["bits_y"] + got["bits_z"]) / (H * W))
  # the rate arithmetic in isolation: oracle bits from the GPU's own hyper-synthesis output
  hs = model.hyper_synthesis(z)
  by, _ = O.rate_bits(wts, hs[..., hs.shape[-1] // 2:].astype(np.float64), q)
  assert np.all(np.abs(got["bits_y"] / by - 1) < 2e-5), (got["bits_y"], by)
  # same image, same bits: deterministic partial sums, independent of the batch it is decoded in
  again = model.decompress(z, q, (H, W), return_bits=True)
  assert np.array_equal(again["bits_y"], got["bits_y"]) and np.array_equal(again["bits_z"], got["bits_z"])
  one = model.decompress(z[1:2], q[1:2], (H, W), return_bits=True)
  assert np.allclose(one["bits_y"], got["bits_y"][1:2], rtol=1e-12) and np.array_equal(one["bits_z"], got["bits_z"][1:2])
  # decode outputs are unchanged by asking for the rate
  plain = model.decompress(z, q, (H, W))
  assert np.array_equal(plain["image"], got["image"]) and np.array_equal(plain["idx"], got["idx"])


def test_rate_without_prior_reports_zero_bits_z_and_int8_symbols_agree(gpu_ctx):
  model, wts, z, q = make_case("two_layer_syn", 2, 64, 128, "stress", "tc", gpu_ctx)
  a = model.decompress(z, q, (64, 128), return_bits=True)
  b = model.decompress(z, q.astype(np.int8), (64, 128), return_bits=True)
  assert np.all(a["bits_z"] == 0) and np.array_equal(a["bits_y"], b["bits_y"]) and np.all(a["bits_y"] > 0)


def test_evaluate_loop_records_match_oracle(gpu_ctx, tmp_path):
  """Row f4: the per-image records of the evaluate loop (bpp, psnr, mse, rd_loss, instance_id) against the oracle,
  batched decode (batch 2 over 5 images, ragged last batch) == image-by-image."""
  import json
  from shallow_ntc_b200 import eval_lib
  H, W = 64, 128
  model, wts, z, q = _rate_case("two_layer_syn", 5, H, W, "tc", gpu_ctx)
  ref = oracle_decode(model, wts, z, q, H, W)
  orig = synthetic.make_original(ref["recon_u8"])
  recs = list(model.evaluate(z, q, orig, batch_size=2, rd_lambda=0.08, extra=dict(rd_lambda=0.08)))
  assert [r["instance_id"] for r in recs] == [0, 1, 2, 3, 4]
  ref_mse, ref_psnr = __import__("oracle.ntc_oracle", fromlist=["x"]).mse_psnr(orig, ref["recon_u8"])
  for i, r in enumerate(recs):
    assert abs(r["psnr"] - ref_psnr[i]) < PSNR_TOL
    assert abs(r["bpp"] / ((ref["bits_y"][i] + ref["bits_z"][i]) / (H * W)) - 1) < 3e-3
    assert abs(r["rd_loss"] - (r["bpp"] + 0.08 * r["mse"])) < 1e-12
  # msssim of the record: single-scale branch at 64 x 128 (< 160 px), against the oracle on the GPU's own uint8 images
  got_img = model.decompress(z, q, (H, W))["image"]
  ref_ms, ref_db = __import__("oracle.ntc_oracle", fromlist=["x"]).msssim(orig, got_img)
  assert all(abs(r["msssim"] - ref_ms[i]) < 2e-6 and abs(r["msssim_db"] - ref_db[i]) < 1e-3 for i, r in enumerate(recs))
  single = list(model.evaluate(z, q, orig, batch_size=1))
  assert all(abs(a["bpp"] - b["bpp"]) < 1e-12 * max(1.0, a["bpp"]) and a["mse"] == b["mse"] for a, b in zip(recs, single))
  path = eval_lib.dump_json(recs, tmp_path / "results.json")
  assert len(json.load(open(path))) == 5


# --------------------------------------------------------------------------------------------------
# MS-SSIM (the validation metric of the evaluate loop, mshyper/models.py:321-332)

def test_msssim_kernel_matches_oracle_and_golden(gpu_ctx):
  """sntc_image_msssim (fp32 separable window out of shared memory, deterministic double partial sums) against the float64
  oracle and the torch-generated fixture: multi-scale with odd sizes, single-scale branch, host and device inputs, identical
  images, bit-identical repeat, and the loud failure for sizes TensorFlow rejects."""
  import os
  from oracle import ntc_oracle as O
  from shallow_ntc_b200 import SntcError
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", "msssim_torch.npz"))
  for name in ("multi_odd", "small", "multi_even"):
    a, b = np.ascontiguousarray(g[name + "_a"]), np.ascontiguousarray(g[name + "_b"])
    val, db = gpu_ctx.msssim(a, b)
    ref, ref_db = O.msssim(a, b)
    assert np.abs(val - g[name + "_val"]).max() < 2e-6 and np.abs(val - ref).max() < 2e-6, (name, val, ref)
    assert np.abs(db - ref_db).max() < 1e-2
    dev, _ = gpu_ctx.msssim(gpu_ctx.to_device(a), gpu_ctx.to_device(b))
    assert np.array_equal(dev, val)
  a = np.ascontiguousarray(g["multi_even_a"])
  assert np.abs(gpu_ctx.msssim(a, a)[0] - 1.0).max() < 1e-6
  with pytest.raises(SntcError):
    gpu_ctx.msssim(np.zeros((1, 100, 200, 3), np.uint8), np.zeros((1, 100, 200, 3), np.uint8))
  with pytest.raises(SntcError):
    gpu_ctx.msssim(np.zeros((1, 64, 64, 3), np.uint8), np.zeros((1, 64, 65, 3), np.uint8))
  assert gpu_ctx.msssim(np.zeros((0, 64, 64, 3), np.uint8), np.zeros((0, 64, 64, 3), np.uint8))[0].shape == (0,)


def test_msssim_full_size_batch(gpu_ctx):
  """BASELINE config 2 shape (24 x 512x768): decoded images against a noisy original; oracle on 2 of the 24 images,
  batch independence (a slice gives the same numbers) on the rest."""
  from oracle import ntc_oracle as O
  B, H, W = 24, 512, 768
  model, wts, z, q = make_case("two_layer_syn", B, H, W, "stress", "tc", gpu_ctx)
  img = model.decompress(z, q, (H, W))["image"]
  orig = synthetic.make_original(img)
  val, _ = gpu_ctx.msssim(orig, img)
  ref, _ = O.msssim(orig[[0, 23]], img[[0, 23]])
  assert np.abs(val[[0, 23]] - ref).max() < 2e-6, (val[[0, 23]], ref)
  part, _ = gpu_ctx.msssim(np.ascontiguousarray(orig[5:9]), np.ascontiguousarray(img[5:9]))
  assert np.array_equal(part, val[5:9])
  assert np.all((val > 0) & (val < 1))
