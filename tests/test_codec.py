"""Host entropy coder + container.  CPU tests: tables against the published tfc construction and the
oracle's pmfs, range-coder round trips (escapes, ragged / empty input), corruption is detected.  GPU tests: the
two-phase decode equals the fused one bit for bit, and compress -> bytes -> decompress reproduces symbols and pixels."""
import numpy as np
import pytest

from oracle import ntc_oracle as O
from shallow_ntc_b200 import EntropyCoder, build_config, synthetic, codec, SntcError


@pytest.fixture(scope="module")
def coder_and_weights():
  m = build_config("two_layer_syn", prior=True)
  w = synthetic.make_weights(m.variable_shapes(), "stress", synthesis_cls="TwoLayerResSynthesis")
  return EntropyCoder(prior_weights=w), w


def test_scale_tables_follow_the_tfc_construction(coder_and_weights):
  from scipy.stats import norm
  k, _ = coder_and_weights
  for row in (0, 1, 17, 40, 63):
    sig = float(O.scale_fn(row))
    off, cdf = k.table(0, row)
    nsym = len(cdf) - 2
    # support: [floor(sigma * Phi^-1(tail_mass / 2)), ceil(sigma * Phi^-1(1 - tail_mass / 2))], tail_mass = 2^-8
    assert off == int(np.floor(sig * norm.ppf(2.0 ** -9))) and off + nsym - 1 == int(np.ceil(sig * norm.ppf(1 - 2.0 ** -9)))
    f = np.diff(cdf.astype(np.int64))
    assert cdf[0] == 0 and cdf[-1] == 4096 and np.all(f >= 1)            # range_coder_precision = 12, no zero frequency
    x = np.arange(off, off + nsym, dtype=np.float64)
    p = 2.0 ** (-O.noisy_normal_bits(x, np.full_like(x, float(row))))
    forced = int((p * 4096 < 1).sum()) + 1                                # entries (and the overflow symbol) lifted to frequency 1
    assert np.abs(f[:-1] / 4096.0 - p).max() < (1.6 + forced) / 4096     # quantised pmf == NoisyNormal pmf up to rounding
    assert abs(f[-1] / 4096.0 - max(1 - p.sum(), 0)) < 2.0 / 4096        # overflow symbol carries the tail mass


def test_prior_tables_match_the_oracle_pmf(coder_and_weights):
  k, w = coder_and_weights
  for ch in (0, 5, 319):
    off, cdf = k.table(1, ch)
    nsym = len(cdf) - 2
    z = np.zeros((nsym, 320)); z[:, ch] = np.arange(off, off + nsym)
    p = 2.0 ** (-O.deep_factorized_bits(z, w)[:, ch])
    f = np.diff(cdf.astype(np.int64))
    forced = int((p * 4096 < 1).sum()) + 1
    assert cdf[-1] == 4096 and np.all(f >= 1) and np.abs(f[:-1] / 4096.0 - p).max() < (1.6 + forced) / 4096


def test_range_coder_round_trip_and_efficiency(coder_and_weights):
  k, _ = coder_and_weights
  rng = np.random.default_rng(3)
  idx = rng.integers(0, 64, size=200_000).astype(np.uint8)
  q = np.rint(rng.normal(size=idx.size) * O.scale_fn(idx.astype(np.float64))).astype(np.int32)
  q[::997] = rng.integers(-30000, 30000, size=q[::997].size)             # escapes through the overflow symbol
  data = k.encode_y(q, idx)
  assert np.array_equal(k.decode_y(data, idx), q)
  # coded size == cross-entropy under the quantised tables (+ escape bits), to 0.1 %
  ideal = 0.0
  for row in range(64):
    off, cdf = k.table(0, row)
    f = np.diff(cdf.astype(np.float64)) / 4096.0
    s = q[idx == row] - off
    ins = (s >= 0) & (s < len(f) - 1)
    ideal += -np.log2(f[s[ins]]).sum() + (~ins).sum() * -np.log2(f[-1])
    v = np.where(s[~ins] < 0, 2 * (-s[~ins]) - 1, 2 * (s[~ins] - (len(f) - 1)) + 2)
    ideal += (2 * np.floor(np.log2(v)) + 1).sum()                         # Elias-gamma
  assert abs(8 * len(data) / ideal - 1) < 1e-3
  assert 8 * len(data) / idx.size < 0.6 * 32                              # it does compress


def test_ragged_empty_and_corrupt_input(coder_and_weights):
  k, _ = coder_and_weights
  assert k.decode_y(k.encode_y(np.zeros(0), np.zeros(0, np.uint8)), np.zeros(0, np.uint8)).size == 0
  one = k.encode_y(np.array([-7]), np.array([3], np.uint8))
  assert k.decode_y(one, np.array([3], np.uint8))[0] == -7
  idx = np.full(5000, 30, np.uint8)
  q = np.arange(5000) % 41 - 20
  data = k.encode_y(q, idx)
  with pytest.raises(SntcError, match="truncated"):
    k.decode_y(data[:len(data) // 2], idx)
  z = np.clip(np.rint(np.random.default_rng(0).normal(0, 1.5, size=(3, 4, 320))), -16, 16)
  assert np.array_equal(k.decode_z(k.encode_z(z), z.shape), z)
  blob = codec.pack([(b"ab", b"cde"), (b"", b"f")], 2, 10, 12, (2, 1, 1, 320), (2, 4, 4, 320))
  strings, (B, H, W), zs, ys = codec.unpack(blob)
  assert strings == [(b"ab", b"cde"), (b"", b"f")] and (B, H, W) == (2, 10, 12) and zs == (2, 1, 1, 320) and ys == (2, 4, 4, 320)
  with pytest.raises(ValueError):
    codec.unpack(blob[:-1])
  with pytest.raises(ValueError):
    codec.unpack(b"XXXX" + blob[4:])


# --------------------------------------------------------------------------------------------------
def test_threaded_decode_equals_serial(coder_and_weights):
  """codec.decompress(threads=N) decodes one image per host thread against shared read-only tables: same symbols."""
  k, _ = coder_and_weights
  rng = np.random.default_rng(3)
  idx = rng.integers(0, 64, (6, 8, 8, 320)).astype(np.uint8)
  q = np.rint(rng.normal(size=idx.shape) * O.scale_fn(idx.astype(np.float64))).astype(np.int32)
  strings = [k.encode_y(q[b], idx[b]) for b in range(6)]
  serial = codec._map(lambda bs: k.decode_y(bs[1], idx[bs[0]]), list(enumerate(strings)), 1)
  threaded = codec._map(lambda bs: k.decode_y(bs[1], idx[bs[0]]), list(enumerate(strings)), 6)
  assert all(np.array_equal(a, b) and np.array_equal(a, q[i]) for i, (a, b) in enumerate(zip(serial, threaded)))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("q_dtype", [np.float32, np.int8])
def test_two_phase_decode_equals_fused_decode(gpu_ctx, precision, q_dtype):
  from helpers import make_case
  model, wts, z, q = make_case("two_layer_syn", 3, 100, 150, "stress", precision, gpu_ctx)
  fused = model.decompress(z, q.astype(q_dtype), (100, 150), return_yhat=True, return_float=True)
  idx = model.decode_hyper(z)
  two = model.decode_latents(q.astype(q_dtype), (100, 150), return_yhat=True, return_float=True)
  assert np.array_equal(idx, fused["idx"])
  assert np.array_equal(two["y_hat"], fused["y_hat"]) and np.array_equal(two["image"], fused["image"])
  assert np.array_equal(two["float"], fused["float"])
  with pytest.raises(SntcError):
    model.decode_latents(q[:2].astype(q_dtype), (100, 150))      # batch does not match the pending phase-1 call


@pytest.mark.gpu
def test_compress_decompress_round_trip(gpu_ctx):
  """symbols -> container bytes -> symbols -> pixels: exact symbols, pixels identical to the direct decode, and on
  latents that follow the model (q ~ NoisyNormal(SCALE_FN(idx))) the coded size is within 3 % of the rate estimate."""
  from shallow_ntc_b200 import build_config
  B, H, W = 2, 128, 192
  model = build_config("two_layer_syn", precision="tc", ctx=gpu_ctx, prior=True)
  wts = synthetic.make_weights(model.variable_shapes(), "stress", synthesis_cls="TwoLayerResSynthesis")
  model.load_weights(wts)
  coder = EntropyCoder(prior_weights=wts)
  zs, ys = model.latent_shapes(B, H, W)
  z, q = synthetic.make_latents(zs, ys)
  blob = codec.compress(model, coder, z, q, (H, W))
  timing = {}
  out = codec.decompress(model, coder, blob, timing=timing)
  direct = model.decompress(z, q, (H, W))
  assert np.array_equal(out["z_hat"], z) and np.array_equal(out["q_y"], q)
  assert np.array_equal(out["image"], direct["image"]) and np.array_equal(out["idx"], direct["idx"])
  assert timing["range_decode_s"] > 0 and timing["gpu_s"] > 0
  # model-matched symbols: the bitstream length agrees with bits_y + bits_z
  idx = model.decode_hyper(z)
  rng = np.random.default_rng(5)
  qm = np.clip(np.rint(rng.normal(size=idx.shape) * O.scale_fn(idx.astype(np.float64))), -2000, 2000).astype(np.float32)
  blob = codec.compress(model, coder, z, qm, (H, W))
  est = model.decompress(z, qm, (H, W), return_bits=True)
  strings, *_ = codec.unpack(blob)
  for b in range(B):
    assert abs(8 * len(strings[b][1]) / est["bits_y"][b] - 1) < 0.03, (8 * len(strings[b][1]), est["bits_y"][b])
    assert abs(8 * len(strings[b][0]) / est["bits_z"][b] - 1) < 0.05, (8 * len(strings[b][0]), est["bits_z"][b])
  back = codec.decompress(model, coder, blob, threads=2)
  assert np.array_equal(back["q_y"], qm)
