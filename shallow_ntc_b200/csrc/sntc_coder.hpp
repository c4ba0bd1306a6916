// Host-side entropy coder of the decode pipeline: quantised CDF tables for the two entropy models
// of mshyper/models.py:27-34, 135, 246-251 and a byte-oriented range coder, so that the symbols the GPU path starts
// from can really come out of a bitstream.  The reference never enables compression (compression=False on every
// call), so there is no reference wire format to match; the table construction follows the published
// tensorflow-compression 2.10 procedure (ContinuousEntropyModelBase._build_tables + pmf_to_quantized_cdf), restated:
//   * row i of the indexed model: NoisyNormal(0, SCALE_FN(i)); support [minima, maxima] =
//     [floor(sigma * Phi^-1(tail_mass / 2)), ceil(sigma * Phi^-1(1 - tail_mass / 2))]; pmf(x) = Phi((x+.5)/sigma) -
//     Phi((x-.5)/sigma); one extra "overflow" symbol with the remaining mass; values outside the support are sent as
//     overflow + an Elias-gamma code (bits at probability 1/2);
//   * pmf -> integer frequencies summing to 2^precision, every frequency >= 1, the excess / deficit removed where it
//     costs the fewest bits (the penalty heuristic of pmf_to_quantized_cdf);
//   * per-channel tables of the hyper-latent from the DeepFactorized prior, support found by scanning outward.
// The range coder is the classic carry-propagating byte coder (64-bit low, 32-bit range, cache + carry run).
// This is an I/O stage on the host: it is timed separately from the GPU hot path (BASELINE north_star).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <string>
#include <vector>

namespace sntc {

struct CdfTable {
  int32_t offset = 0;              // value of symbol 0 (minima)
  int32_t nsym = 0;                // regular symbols; symbol nsym = overflow
  std::vector<uint32_t> cdf;       // nsym + 2 entries, cdf[0] = 0, cdf[nsym + 1] = 1 << precision
};

// pmf (last entry = overflow mass) -> cumulative integer frequencies with total 1 << precision, each frequency >= 1
inline std::vector<uint32_t> pmf_to_quantized_cdf(const std::vector<double>& pmf, int precision) {
  const int n = (int)pmf.size();
  const int64_t total = (int64_t)1 << precision;
  std::vector<int64_t> f(n);
  int64_t sum = 0;
  for (int i = 0; i < n; ++i) { f[i] = std::max<int64_t>(1, (int64_t)std::llrint(pmf[i] * (double)total)); sum += f[i]; }
  // penalty of moving entry i by one count, in bits weighted by its mass
  auto lower_pen = [&](int i) { return pmf[i] * (std::log2((double)f[i]) - std::log2((double)f[i] - 1.0)); };
  auto upper_pen = [&](int i) { return pmf[i] * (std::log2((double)f[i]) - std::log2((double)f[i] + 1.0)); };
  typedef std::pair<double, int> Item;   // (penalty, index): smallest penalty first, ties by index
  if (sum > total) {
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> pq;
    for (int i = 0; i < n; ++i) if (f[i] > 1) pq.push({lower_pen(i), i});
    while (sum > total && !pq.empty()) {
      const int i = pq.top().second; pq.pop();
      --f[i]; --sum;
      if (f[i] > 1) pq.push({lower_pen(i), i});
    }
  } else if (sum < total) {
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> pq;
    for (int i = 0; i < n; ++i) pq.push({upper_pen(i), i});
    while (sum < total) {
      const int i = pq.top().second; pq.pop();
      ++f[i]; ++sum;
      pq.push({upper_pen(i), i});
    }
  }
  std::vector<uint32_t> cdf(n + 1, 0);
  for (int i = 0; i < n; ++i) cdf[i + 1] = cdf[i] + (uint32_t)f[i];
  return cdf;
}

inline double norm_cdf(double x) { return 0.5 * std::erfc(-x * 0.70710678118654752440); }
// Phi^-1 by bisection on erfc (called 64 times at construction)
inline double norm_ppf(double p) {
  double lo = -40.0, hi = 40.0;
  for (int it = 0; it < 200; ++it) { const double mid = 0.5 * (lo + hi); if (norm_cdf(mid) < p) lo = mid; else hi = mid; }
  return 0.5 * (lo + hi);
}

inline CdfTable build_normal_table(double sigma, double tail_mass, int precision) {
  CdfTable t;
  const int32_t minima = (int32_t)std::floor(sigma * norm_ppf(0.5 * tail_mass));
  const int32_t maxima = (int32_t)std::ceil(sigma * norm_ppf(1.0 - 0.5 * tail_mass));
  t.offset = minima; t.nsym = maxima - minima + 1;
  std::vector<double> pmf(t.nsym + 1);
  double s = 0;
  for (int i = 0; i < t.nsym; ++i) {
    const double x = (double)(minima + i), ax = std::fabs(x);
    pmf[i] = norm_cdf(-(ax - 0.5) / sigma) - norm_cdf(-(ax + 0.5) / sigma);   // upper-tail form: no cancellation
    s += pmf[i];
  }
  pmf[t.nsym] = std::max(1.0 - s, 0.0);
  t.cdf = pmf_to_quantized_cdf(pmf, precision);
  return t;
}

// DeepFactorized(num_filters = (3,3,3)) of one channel from the RAW tfc variables (double precision)
struct DeepFactorizedChannel {
  double m0[3], b0[3], f0[3], m1[9], b1[3], f1[3], m2[9], b2[3], f2[3], m3[3], b3;
  double logits(double x) const {
    double h[3], g[3];
    for (int i = 0; i < 3; ++i) { const double v = m0[i] * x + b0[i]; h[i] = v + f0[i] * std::tanh(v); }
    for (int i = 0; i < 3; ++i) { double v = b1[i]; for (int j = 0; j < 3; ++j) v += m1[i * 3 + j] * h[j]; g[i] = v + f1[i] * std::tanh(v); }
    for (int i = 0; i < 3; ++i) h[i] = g[i];
    for (int i = 0; i < 3; ++i) { double v = b2[i]; for (int j = 0; j < 3; ++j) v += m2[i * 3 + j] * h[j]; g[i] = v + f2[i] * std::tanh(v); }
    double v = b3;
    for (int j = 0; j < 3; ++j) v += m3[j] * g[j];
    return v;
  }
  static double sigmoid(double x) { return x >= 0 ? 1.0 / (1.0 + std::exp(-x)) : std::exp(x) / (1.0 + std::exp(x)); }
  double prob(double z) const {   // c(z + .5) - c(z - .5) on the side of the median where both sigmoids are small
    const double lo = logits(z - 0.5), up = logits(z + 0.5);
    const double sg = (lo + up) > 0 ? -1.0 : 1.0;
    return std::fabs(sigmoid(sg * up) - sigmoid(sg * lo));
  }
};

inline CdfTable build_prior_table(const DeepFactorizedChannel& ch, double tail_mass, int precision) {
  CdfTable t;
  const int LIM = 1 << 14;
  int lo = 0, hi = 0;
  while (lo > -LIM && DeepFactorizedChannel::sigmoid(ch.logits(lo - 0.5)) > 0.5 * tail_mass) --lo;
  while (hi < LIM && DeepFactorizedChannel::sigmoid(-ch.logits(hi + 0.5)) > 0.5 * tail_mass) ++hi;
  t.offset = lo; t.nsym = hi - lo + 1;
  std::vector<double> pmf(t.nsym + 1);
  double s = 0;
  for (int i = 0; i < t.nsym; ++i) { pmf[i] = ch.prob((double)(lo + i)); s += pmf[i]; }
  pmf[t.nsym] = std::max(1.0 - s, 0.0);
  t.cdf = pmf_to_quantized_cdf(pmf, precision);
  return t;
}

// ------------------------------------------------------------------------------------------------
class RangeEncoder {
 public:
  void encode(uint32_t lo, uint32_t hi, int precision) {   // [lo, hi) out of 1 << precision
    const uint32_t r = range_ >> precision;
    low_ += (uint64_t)r * lo;
    range_ = r * (hi - lo);
    while (range_ < (1u << 24)) { shift_low(); range_ <<= 8; }
  }
  std::vector<uint8_t> finish() {
    for (int i = 0; i < 5; ++i) shift_low();
    return std::move(out_);
  }
 private:
  void shift_low() {   // the canonical carry-propagating byte emitter (one leading zero byte, skipped by the decoder)
    if ((uint32_t)low_ < 0xFF000000u || (low_ >> 32) != 0) {
      const uint8_t carry = (uint8_t)(low_ >> 32);
      uint8_t temp = cache_;
      do { out_.push_back((uint8_t)(temp + carry)); temp = 0xFF; } while (--cache_size_ != 0);
      cache_ = (uint8_t)(low_ >> 24);
    }
    ++cache_size_;
    low_ = (low_ & 0x00FFFFFFu) << 8;
  }
  uint64_t low_ = 0; uint32_t range_ = 0xFFFFFFFFu;
  uint8_t cache_ = 0; uint64_t cache_size_ = 1;
  std::vector<uint8_t> out_;
};

class RangeDecoder {
 public:
  RangeDecoder(const uint8_t* p, size_t n) : p_(p), end_(p + n) {
    for (int i = 0; i < 5; ++i) code_ = (code_ << 8) | next();   // the first byte is the encoder's leading zero
  }
  // target in [0, 1 << precision): the caller finds the symbol, then calls consume with its interval
  uint32_t peek(int precision) {
    r_ = range_ >> precision;
    const uint32_t v = code_ / r_;
    const uint32_t top = (1u << precision) - 1;
    return v < top ? v : top;
  }
  void consume(uint32_t lo, uint32_t hi) {
    code_ -= r_ * lo;
    range_ = r_ * (hi - lo);
    while (range_ < (1u << 24)) { code_ = (code_ << 8) | next(); range_ <<= 8; }
  }
  bool overrun() const { return overrun_ > 8; }
 private:
  uint32_t next() { if (p_ < end_) return *p_++; ++overrun_; return 0; }
  const uint8_t* p_; const uint8_t* end_;
  uint32_t code_ = 0, range_ = 0xFFFFFFFFu, r_ = 0; int overrun_ = 0;
};

inline void encode_symbol(RangeEncoder& enc, const CdfTable& t, int32_t value, int precision) {
  int32_t s = value - t.offset;
  if (s >= 0 && s < t.nsym) { enc.encode(t.cdf[s], t.cdf[s + 1], precision); return; }
  enc.encode(t.cdf[t.nsym], t.cdf[t.nsym + 1], precision);          // overflow symbol
  // Elias-gamma of v >= 1: below the support -> odd, above -> even
  uint32_t v = s < 0 ? (uint32_t)(2 * (int64_t)(-s) - 1) : (uint32_t)(2 * (int64_t)(s - t.nsym) + 2);
  int nbits = 0;
  while ((v >> nbits) > 1) ++nbits;                                  // floor(log2 v)
  for (int i = 0; i < nbits; ++i) enc.encode(0, 1, 1);
  enc.encode(1, 2, 1);
  for (int i = nbits - 1; i >= 0; --i) { const uint32_t b = (v >> i) & 1u; enc.encode(b, b + 1, 1); }
}

inline int32_t decode_symbol(RangeDecoder& dec, const CdfTable& t, int precision) {
  const uint32_t target = dec.peek(precision);
  // largest s with cdf[s] <= target
  int lo = 0, hi = t.nsym + 1;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (t.cdf[mid] <= target) lo = mid; else hi = mid; }
  dec.consume(t.cdf[lo], t.cdf[lo + 1]);
  if (lo < t.nsym) return t.offset + lo;
  int nbits = 0;
  for (;;) { const uint32_t b = dec.peek(1); dec.consume(b, b + 1); if (b) break; if (++nbits > 40) return 0; }
  uint32_t v = 1;
  for (int i = 0; i < nbits; ++i) { const uint32_t b = dec.peek(1); dec.consume(b, b + 1); v = (v << 1) | b; }
  if (v & 1u) return t.offset - (int32_t)((v + 1) / 2);
  return t.offset + t.nsym + (int32_t)((v - 2) / 2);
}

struct Coder {
  int precision = 12;
  double tail_mass = 1.0 / 256;
  std::vector<CdfTable> scale_rows;    // indexed by the scale-table row (uint8 idx)
  std::vector<CdfTable> prior_rows;    // indexed by the hyper-latent channel
};

}  // namespace sntc
