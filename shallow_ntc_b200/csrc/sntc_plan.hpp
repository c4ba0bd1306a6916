// Host-side layer plans for the decode path: transform classes -> layer lists -> band decomposition.
//
// Band decomposition of a transposed convolution (the "no overlap-add" form used by every kernel):
//   ConvT(k, s, p):  out[o] += in[n] * W[a],  o = n*s + a - p.
//   Write u = o + p = s*m + phi  (cell m, phase phi in [0,s)).  Then a = phi + s*j, n = m - j for
//   j = 0 .. T(phi)-1 with T(phi) = #{a in [0,k): a == phi mod s}.  Contiguous phases with equal T
//   form a *band*; a 2-D band (by, bx) is a dense GEMM
//       C[cell m, (phi_y, phi_x, co)] = sum_{jy, jx, ci} in[m - j, ci] * W[phi + s*j, co, ci]
//   with N = nphi_y*nphi_x*Cout columns and K = Ty*Tx*Cin.  Every output element is produced by
//   exactly one band exactly once (no atomics, no col2im), and no MAC is wasted.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>

#include "../../include/sntc.h"

namespace sntc {

enum KernelLayout { LAYOUT_KERAS_OI = 0 /* [kh,kw,Cout,Cin] */, LAYOUT_TFC_IO = 1 /* [kh,kw,Cin,Cout] */,
                    // tf.nn.depth_to_space(x, 2) followed by a Keras Conv2D 1x1 (variable [1,1,Cin/4,Cout]) IS a transposed conv with
                    // k = s = 2, p = 0 whose tap (dy, dx) only sees input channels [(2 dy + dx) Cin/4, (2 dy + dx + 1) Cin/4):
                    //   out[2y+dy, 2x+dx, co] = sum_c in[y, x, (2 dy + dx) Cin/4 + c] * K[0, 0, c, co]      (NHWC / DCR order)
                    LAYOUT_D2S_1X1 = 2,
                    // the input-gradient of another layer written as a stride-1 layer of this plan (make_backward_conv)
                    LAYOUT_BACKWARD = 3,
                    // the per-input-pixel contraction of a final layer in col2im form (make_col2im_conv)
                    LAYOUT_COL2IM = 4 };
enum GdnKind { GDN_NONE = 0, GDN_1 = 1 /* beta + |x| gamma */, GDN_CLASSIC = 2 /* sqrt(beta + x^2 gamma) */ };

struct Band1D {
  int phi0, nphi, T;   // phases [phi0, phi0+nphi), taps j in [0,T)
  int mlo;             // first cell: outputs of phase phi sit at o = s*(mlo + i) + phi - p, i in [0, n_in)
};

inline int ceil_div_floor(int a, int b) {  // ceil(a/b) for b>0, any sign of a
  return a >= 0 ? (a + b - 1) / b : -((-a) / b);
}
inline int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// Phases are grouped into a band when they share the tap count T AND the first cell mlo; every phase
// owns exactly n_in outputs, so every band is a GEMM over exactly h*w cells per image (no ragged tiles).
inline std::vector<Band1D> bands_1d(int k, int s, int p) {
  std::vector<Band1D> out;
  for (int phi = 0; phi < s; ++phi) {
    int T = phi < k ? (k - phi + s - 1) / s : 0;
    int mlo = std::max(0, ceil_div_floor(p - phi, s));
    if (!out.empty() && out.back().T == T && out.back().mlo == mlo) out.back().nphi++;
    else out.push_back({phi, 1, T, mlo});
  }
  return out;
}

struct Band {
  int phy0, nphy, Ty, phx0, nphx, Tx;
  int N, Npad, K;        // GEMM dims: N = nphy*nphx*cout (padded to 4), K = Ty*Tx*cin
  size_t w_off;          // float offset of this band's [K][Npad] matrix in the packed weight buffer
  // cell ranges for an input of n_in rows/cols are computed per call (depend on h, w)
};

// cell range [mlo, mlo+cnt) of a 1-D band for n_in input samples: always exactly n_in cells
inline void cell_range(const Band1D& b, int s, int p, int n_in, int* mlo, int* cnt) {
  (void)s; (void)p;
  *mlo = b.mlo;
  *cnt = n_in;
}

struct ConvSource { std::string kernel, bias; int cout; };

struct ConvLayer {
  int k = 0, s = 1, p = 0, cin = 0, cout = 0;
  int cin_pad = 0;        // cin rounded up to a multiple of 4 (ones channel of use_offset included)
  int layout = LAYOUT_KERAS_OI;
  bool has_bias = true;
  bool append_ones = false;  // JPEGLikeSynthesis(use_offset=True): input gets a constant-1 channel
  // Merged form (final layers with tiny Cout on the tensor-core path): ONE band holds all s*s output residues
  // r = o mod s of a cell (o = s*i + r), with a uniform tap window d = i - n in [dlo, dlo + T): a = r + p + s*d
  // (taps outside [0,k) are zero weights).  N = s*s*cout instead of a handful of 3-column bands.
  bool merged = false; int dlo = 0;
  // col2im form of a final layer with tiny Cout and a wide input (deep decoders: ConvT(5, 2, 192 -> 3), ConvT(9, 4, 256 -> 3)) on the
  // tensor-core path: ONE 1x1 GEMM per input pixel, P[n, (co, a_y, a_x)] = sum_ci in[n, ci] W[a, co, ci] (c2i[0], K = Cin: the input
  // is read once instead of once per tap), and an overlap-add epilogue out[o] = sum_{n in 3x3} P[n, a = o + p - s n] over
  // overlapping tiles (sntc_kernels_tc.cuh, TC_EPI_COL2IM).  c2i_kp = columns reserved per output channel (k*k rounded up to 32).
  bool col2im = false; int c2i_kp = 0; std::vector<ConvLayer> c2i;
  int out_crop = 0;          // the output grid is h*s - out_crop rows / columns (backward layers: the input carries T-1 extra rows)
  std::vector<ConvLayer> bwd_src;   // LAYOUT_BACKWARD: copy of the forward layer whose input-gradient this layer computes
  int act = SNTC_ACT_NONE;   // only NONE / RELU / LEAKY_RELU are fused into the conv
  std::vector<ConvSource> sources;  // concatenated along Cout (base || res)
  std::vector<Band1D> by, bx;
  std::vector<Band> bands;
  size_t w_floats = 0;
  float* d_w = nullptr;      // packed band matrices
  float* d_bias = nullptr;   // [cout] (zeros when !has_bias)
  std::vector<float> h_bias; // host copy
  // rgb-cell packing for final layers with tiny Cout: [k*k][cin_pad][4]
  float* d_w_rgb = nullptr;
  std::vector<float> h_w_tail;   // [k*k][cin][3] + bias[3]: by-value kernel parameter of the constant-bank tail kernel
  // tensor-core packing (fp16 hi/lo), see sntc_kernels_tc.cuh
  void* d_w_hi = nullptr; void* d_w_lo = nullptr; float w_scale = 1.f; int cin_tc = 0; int n_tc_rows = 0;
  std::vector<size_t> tc_band_row0;
};

struct GdnLayer {
  int C = 0; int kind = GDN_NONE; bool inverse = true;
  std::string beta, gamma;
  float* d_beta = nullptr; float* d_gamma = nullptr;  // gamma [in][out] (padded to Npad columns)
  std::vector<float> h_beta, h_gamma;                 // host copies ([C], [C][C]) of small layers (kernel parameters)
  int Npad = 0;
};

enum OpType {
  OP_CONVT = 0,     // band-GEMM transposed conv (+bias, +relu/leaky)
  OP_GDN = 1,       // GDN over all channels of the current tensor
  OP_ACT_RES = 2,   // current tensor is [.., 2*C] = base || res: out[.., C] = act(base) + res
  OP_CONVT_RGB = 3, // final conv to <=4 channels, fused crop + uint8 epilogue
  OP_STASH = 4,     // TwoLayerResSynthesis(res_type="d2s"): the current tensor is the residual branch before its last depth_to_space;
                    // keep it aside and go back to the transform's input
  OP_ACT_RES_D2S = 5  // out[b,y,x,c] = act(cur[b,y,x,:])[c] + stash[b, y/2, x/2, (2 (y%2) + x%2) C + c]
};

struct Op { int type; int conv = -1; int gdn = -1; int act = SNTC_ACT_NONE; };

struct Transform {
  int kind = SNTC_T_NONE;
  int in_channels = 0, out_channels = 0, upsample = 1;
  std::vector<ConvLayer> convs;
  std::vector<GdnLayer> gdns;
  std::vector<Op> ops;
};

struct VarSpec { std::string name; std::vector<int64_t> shape; };

inline int keras_pad(int k, int s) { return std::max(k - s, 0) / 2; }
inline int tfc_pad(int k) { return (k - 1) / 2; }

inline void finish_conv(ConvLayer& c) {
  c.cin_pad = ((c.cin + (c.append_ones ? 1 : 0)) + 3) / 4 * 4;
  c.by = bands_1d(c.k, c.s, c.p);
  c.bx = bands_1d(c.k, c.s, c.p);
  c.bands.clear();
  size_t off = 0;
  for (auto& y : c.by) for (auto& x : c.bx) {
    Band b{};
    b.phy0 = y.phi0; b.nphy = y.nphi; b.Ty = y.T;
    b.phx0 = x.phi0; b.nphx = x.nphi; b.Tx = x.T;
    b.N = y.nphi * x.nphi * c.cout;
    b.Npad = (b.N + 3) / 4 * 4;
    b.K = y.T * x.T * c.cin_pad;
    b.w_off = off;
    off += (size_t)b.K * b.Npad;
    c.bands.push_back(b);
  }
  c.w_floats = off;
}

// Kernel index of (residue / phase f, tap j) of a band along one axis; < 0 or >= k means "no such tap" (zero weight).
inline int band_tap_index(const ConvLayer& c, int phi0, int f, int j) {
  return c.merged ? f + c.p + c.s * (c.dlo + j) : phi0 + f + c.s * j;
}

// Re-plans `c` in the merged form (see ConvLayer::merged).
inline void finish_conv_merged(ConvLayer& c) {
  c.cin_pad = ((c.cin + (c.append_ones ? 1 : 0)) + 3) / 4 * 4;
  int dlo = 1 << 30, dhi = -(1 << 30);
  for (int r = 0; r < c.s; ++r)
    for (int d = -c.k; d <= c.k; ++d) {
      const int a = r + c.p + c.s * d;
      if (a >= 0 && a < c.k) { dlo = std::min(dlo, d); dhi = std::max(dhi, d); }
    }
  c.merged = true; c.dlo = dlo;
  const int T = dhi - dlo + 1;
  c.by = {Band1D{0, c.s, T, -dlo}};
  c.bx = c.by;
  Band b{};
  b.phy0 = 0; b.nphy = c.s; b.Ty = T; b.phx0 = 0; b.nphx = c.s; b.Tx = T;
  b.N = c.s * c.s * c.cout; b.Npad = (b.N + 3) / 4 * 4; b.K = T * T * c.cin_pad; b.w_off = 0;
  c.bands = {b};
  c.w_floats = (size_t)b.K * b.Npad;
}

inline ConvLayer make_conv(const std::string& prefix, const std::string& name, int k, int s, int cin, int cout,
                           int layout, bool bias, int act) {
  ConvLayer c;
  c.k = k; c.s = s; c.cin = cin; c.cout = cout; c.layout = layout; c.has_bias = bias; c.act = act;
  c.p = layout == LAYOUT_KERAS_OI ? keras_pad(k, s) : tfc_pad(k);
  c.sources.push_back({prefix + "." + name + ".kernel", bias ? prefix + "." + name + ".bias" : std::string(), cout});
  finish_conv(c);
  return c;
}

// Input-gradient of the transposed conv `f` as a FORWARD layer of this plan (decoder backward, mshyper/models.py:401-408).
//   gin[n, ci] = sum_{a, co} g[s n + a - p, co] W[a, co, ci].   Write a = s d + r (r in [0,s), d in [0,T), T = ceil(k/s)) and
//   G'[m, (ry, rx, co)] = g[s m + r - p, co]  (zero outside g; m in [0, h + T - 1): a shifted space-to-depth of the gradient).
//   Then gin[n] = sum_{d} G'[n + d] Wb[d],  Wb[d][(r, co), ci] = W[s d + r, co, ci] (zero where s d + r >= k):
//   a stride-1 layer with T x T taps, Cin' = s*s*Cout, Cout' = Cin, p' = T - 1 in the ConvT form (tap a' = T - 1 - d),
//   run on the [h + T - 1, w + T - 1] grid of G' with the output cropped to [h, w].
inline ConvLayer make_backward_conv(const ConvLayer& f) {
  ConvLayer c;
  const int T = (f.k + f.s - 1) / f.s;
  c.k = T; c.s = 1; c.p = T - 1; c.cin = f.s * f.s * f.cout; c.cout = f.cin;
  c.layout = LAYOUT_BACKWARD; c.has_bias = false; c.act = SNTC_ACT_NONE; c.out_crop = T - 1;
  for (auto& s : f.sources) c.sources.push_back({s.kernel, std::string(), 0});   // names only (weight range of the fp16 packing)
  c.sources[0].cout = c.cout;
  c.bwd_src.push_back(f);
  c.bwd_src[0].bwd_src.clear();
  finish_conv(c);
  return c;
}

// Final layers the col2im form serves: every output depends on the 3x3 neighbourhood of its cell at most.
inline bool col2im_eligible(const ConvLayer& c) {
  if (c.append_ones || c.cout > 4 || c.cout < 1 || (c.s != 2 && c.s != 4) || c.act != SNTC_ACT_NONE) return false;
  if (c.cin % 8 != 0 || c.cin < 64) return false;
  const int hi = (c.s - 1 + c.p) / c.s;                               // n_max - i
  const int lo = -ceil_div_floor(c.p - c.k + 1, c.s);                 // i - n_min
  // all output channels must fit ONE n-tile of <= 96 columns (k5: 3 x 32): with one n-tile per channel (k9 s4: 3 x 96) the input is
  // read and staged once per channel and the merged form wins (bls2017 4K final layer: 1.04 vs 0.93 ms)
  return hi <= 1 && lo <= 1 && c.cout * ((c.k * c.k + 31) / 32 * 32) <= 96;
}

inline ConvLayer make_col2im_conv(const ConvLayer& f) {
  ConvLayer c;
  const int kp = (f.k * f.k + 31) / 32 * 32;
  c.k = 1; c.s = 1; c.p = 0; c.cin = f.cin; c.cout = f.cout * kp;
  c.layout = LAYOUT_COL2IM; c.has_bias = false; c.act = SNTC_ACT_NONE;
  for (auto& s : f.sources) c.sources.push_back({s.kernel, std::string(), 0});
  c.sources[0].cout = c.cout;
  c.bwd_src.push_back(f);
  c.bwd_src[0].bwd_src.clear(); c.bwd_src[0].c2i.clear();
  c.c2i_kp = kp;
  finish_conv(c);
  return c;
}

inline int conv_act_of(int act) {
  return (act == SNTC_ACT_RELU || act == SNTC_ACT_LEAKY_RELU) ? act : SNTC_ACT_NONE;
}

// Builds the layer list of one registry class (common/transforms.py) from its constructor kwargs.
inline Transform build_transform(const sntc_transform_desc& d, const std::string& prefix) {
  Transform t;
  t.kind = d.kind;
  t.in_channels = d.in_channels;
  auto add_conv = [&](ConvLayer c) { t.convs.push_back(std::move(c)); Op o; o.type = OP_CONVT; o.conv = (int)t.convs.size() - 1; t.ops.push_back(o); };
  auto add_gdn = [&](const std::string& name, int C, int kind, bool inverse) {
    GdnLayer g; g.C = C; g.kind = kind; g.inverse = inverse;
    g.beta = prefix + "." + name + ".beta"; g.gamma = prefix + "." + name + ".gamma";
    t.gdns.push_back(g); Op o; o.type = OP_GDN; o.gdn = (int)t.gdns.size() - 1; t.ops.push_back(o);
  };
  auto add_act = [&](int act, const std::string& name, int C) {   // activation applied after a conv that did not fuse it
    if (act == SNTC_ACT_IGDN1) add_gdn(name, C, GDN_1, true);
    else if (act == SNTC_ACT_GDN1) add_gdn(name, C, GDN_1, false);
  };
  if (d.in_channels <= 0) throw std::invalid_argument("transform in_channels must be > 0");
  switch (d.kind) {
    case SNTC_T_HYPER_SYNTHESIS: {  // transforms.py:222-232
      int C = d.channels[0];
      if (C <= 0) throw std::invalid_argument("HyperSynthesis: bottleneck_size must be > 0");
      int act = d.activation;
      if (act != SNTC_ACT_RELU && act != SNTC_ACT_LEAKY_RELU && act != SNTC_ACT_NONE)
        throw std::invalid_argument("HyperSynthesis: activation_type must be relu / leaky_relu / None");
      add_conv(make_conv(prefix, "layer_0", 5, 2, d.in_channels, C, LAYOUT_KERAS_OI, true, act));
      add_conv(make_conv(prefix, "layer_1", 5, 2, C, (int)(C * 1.5), LAYOUT_KERAS_OI, true, act));
      add_conv(make_conv(prefix, "layer_2", 3, 1, (int)(C * 1.5), C * 2, LAYOUT_KERAS_OI, true, SNTC_ACT_NONE));
      t.out_channels = C * 2; t.upsample = 4;
      break;
    }
    case SNTC_T_JPEG_LIKE_HYPER: {  // transforms.py:364-377
      int C = d.channels[0]; int k = d.kernel_sizes[0] > 0 ? d.kernel_sizes[0] : 6;
      add_conv(make_conv(prefix, "conv", k, 4, d.in_channels, C * 2, LAYOUT_KERAS_OI, true, SNTC_ACT_NONE));
      t.out_channels = C * 2; t.upsample = 4;
      break;
    }
    case SNTC_T_HYPER_SMALL: {  // transforms.py:250-262
      int C = d.channels[0];
      add_conv(make_conv(prefix, "layer_0", 5, 2, d.in_channels, (int)(C * 1.5), LAYOUT_TFC_IO, true, SNTC_ACT_RELU));
      add_conv(make_conv(prefix, "layer_1", 3, 1, (int)(C * 1.5), C * 2, LAYOUT_TFC_IO, true, SNTC_ACT_NONE));
      t.out_channels = C * 2; t.upsample = 2;
      break;
    }
    case SNTC_T_JPEG_LIKE_SYNTHESIS: {  // transforms.py:265-295
      int cout = d.channels[0] > 0 ? d.channels[0] : 3;
      int k = d.kernel_sizes[0] > 0 ? d.kernel_sizes[0] : 16;
      int s = d.strides[0] > 0 ? d.strides[0] : 16;
      ConvLayer c = make_conv(prefix, "conv", k, s, d.in_channels, cout, LAYOUT_KERAS_OI, d.use_bias != 0, SNTC_ACT_NONE);
      if (d.use_offset) { c.append_ones = true; finish_conv(c); }
      add_conv(std::move(c));
      t.out_channels = cout; t.upsample = s;
      break;
    }
    case SNTC_T_TWO_LAYER: {  // transforms.py:298-317
      int C1 = d.channels[0], Co = d.channels[1];
      int k1 = d.kernel_sizes[0], k2 = d.kernel_sizes[1], s1 = d.strides[0], s2 = d.strides[1];
      add_conv(make_conv(prefix, "conv1", k1, s1, d.in_channels, C1, LAYOUT_KERAS_OI, true, conv_act_of(d.activation)));
      add_act(d.activation, "activation", C1);
      add_conv(make_conv(prefix, "conv2", k2, s2, C1, Co, LAYOUT_KERAS_OI, true, SNTC_ACT_NONE));
      t.out_channels = Co; t.upsample = s1 * s2;
      break;
    }
    case SNTC_T_TWO_LAYER_RES: {  // transforms.py:320-361, res_type="conv"
      int C1 = d.channels[0], Co = d.channels[1];
      int k1 = d.kernel_sizes[0], k2 = d.kernel_sizes[1], s1 = d.strides[0], s2 = d.strides[1];
      ConvLayer c = make_conv(prefix, "base_conv", k1, s1, d.in_channels, C1, LAYOUT_KERAS_OI, true, SNTC_ACT_NONE);
      c.sources.push_back({prefix + ".res.kernel", prefix + ".res.bias", C1});   // base || res along Cout
      c.cout = 2 * C1;
      finish_conv(c);
      add_conv(std::move(c));
      {
        Op o; o.type = OP_ACT_RES; o.act = d.activation;
        if (d.activation == SNTC_ACT_IGDN1 || d.activation == SNTC_ACT_GDN1) {
          GdnLayer g; g.C = C1; g.kind = GDN_1; g.inverse = d.activation == SNTC_ACT_IGDN1;
          g.beta = prefix + ".activation.beta"; g.gamma = prefix + ".activation.gamma";
          t.gdns.push_back(g); o.gdn = (int)t.gdns.size() - 1;
        }
        t.ops.push_back(o);
      }
      add_conv(make_conv(prefix, "out_conv", k2, s2, C1, Co, LAYOUT_KERAS_OI, true, SNTC_ACT_NONE));
      t.out_channels = Co; t.upsample = s1 * s2;
      break;
    }
    case SNTC_T_TWO_LAYER_RES_D2S: {  // transforms.py:320-361, res_type="d2s" (:339-348)
      int C1 = d.channels[0], Co = d.channels[1];
      int k1 = d.kernel_sizes[0], k2 = d.kernel_sizes[1], s1 = d.strides[0], s2 = d.strides[1];
      if (s1 != 8) throw std::invalid_argument("TwoLayerResSynthesis(res_type='d2s'): three depth_to_space(2) steps need strides[0] == 8");
      if (d.in_channels % 16 != 0) throw std::invalid_argument("TwoLayerResSynthesis(res_type='d2s'): input channels must be a multiple of 16");
      auto d2s_conv = [&](const std::string& name, int cin, int cout) {
        ConvLayer c = make_conv(prefix, name, 2, 2, cin, cout, LAYOUT_D2S_1X1, true, SNTC_ACT_LEAKY_RELU);
        c.p = 0;
        finish_conv(c);
        return c;
      };
      add_conv(d2s_conv("res.conv_0", d.in_channels, 192));
      add_conv(d2s_conv("res.conv_1", 192, 4 * C1));
      { Op o; o.type = OP_STASH; t.ops.push_back(o); }
      add_conv(make_conv(prefix, "base_conv", k1, s1, d.in_channels, C1, LAYOUT_KERAS_OI, true, SNTC_ACT_NONE));
      {
        Op o; o.type = OP_ACT_RES_D2S; o.act = d.activation;
        if (d.activation == SNTC_ACT_IGDN1 || d.activation == SNTC_ACT_GDN1) {
          GdnLayer g; g.C = C1; g.kind = GDN_1; g.inverse = d.activation == SNTC_ACT_IGDN1;
          g.beta = prefix + ".activation.beta"; g.gamma = prefix + ".activation.gamma";
          t.gdns.push_back(g); o.gdn = (int)t.gdns.size() - 1;
        }
        t.ops.push_back(o);
      }
      add_conv(make_conv(prefix, "out_conv", k2, s2, C1, Co, LAYOUT_KERAS_OI, true, SNTC_ACT_NONE));
      t.out_channels = Co; t.upsample = s1 * s2;
      break;
    }
    case SNTC_T_MBT2018: {  // transforms.py:158-175
      int C = d.channels[0]; int Co = d.channels[1] > 0 ? d.channels[1] : C; int nl = d.n_layers > 0 ? d.n_layers : 4;
      int cin = d.in_channels; int up = 1;
      for (int i = 0; i < nl; ++i) {
        bool last = i + 1 == nl;
        add_conv(make_conv(prefix, "layer_" + std::to_string(i), 5, 2, cin, last ? Co : C, LAYOUT_TFC_IO, true, SNTC_ACT_NONE));
        // tfc.GDN(inverse=True) with the tfc 2.x defaults alpha_parameter = epsilon_parameter = 1 is IGDN1 (see sntc.h,
        // SNTC_ACT_IGDN_CLASSIC); the (alpha=2, epsilon=.5) form only on request
        if (!last) add_gdn("igdn_" + std::to_string(i), C, d.activation == SNTC_ACT_IGDN_CLASSIC ? GDN_CLASSIC : GDN_1, true);
        cin = C; up *= 2;
      }
      t.out_channels = Co; t.upsample = up;
      break;
    }
    case SNTC_T_BLS2017: {  // transforms.py:115-134
      int C = d.channels[0];
      add_conv(make_conv(prefix, "layer_0", 5, 2, d.in_channels, C, LAYOUT_TFC_IO, true, SNTC_ACT_NONE));
      add_gdn("igdn_0", C, GDN_1, true);
      add_conv(make_conv(prefix, "layer_1", 5, 2, C, C, LAYOUT_TFC_IO, true, SNTC_ACT_NONE));
      add_gdn("igdn_1", C, GDN_1, true);
      add_conv(make_conv(prefix, "layer_2", 9, 4, C, 3, LAYOUT_TFC_IO, true, SNTC_ACT_NONE));
      t.out_channels = 3; t.upsample = 16;
      break;
    }
    case SNTC_T_CNN: {  // transforms.py:195-206 (one activation object shared by layers 0-2)
      int C = d.channels[0]; int Co = d.channels[1] > 0 ? d.channels[1] : 3;
      int cin = d.in_channels;
      for (int i = 0; i < 4; ++i) {
        bool last = i == 3;
        add_conv(make_conv(prefix, "layer_" + std::to_string(i), 5, 2, cin, last ? Co : C, LAYOUT_KERAS_OI, true,
                           last ? SNTC_ACT_NONE : conv_act_of(d.activation)));
        if (!last) add_act(d.activation, "activation", C);
        cin = C;
      }
      t.out_channels = Co; t.upsample = 16;
      break;
    }
    default:
      throw std::invalid_argument("unknown transform kind " + std::to_string(d.kind));
  }
  for (auto& c : t.convs)
    if (c.k <= 0 || c.s <= 0 || c.cin <= 0 || c.cout <= 0) throw std::invalid_argument("transform: non-positive layer dimension");
  return t;
}

inline std::vector<VarSpec> transform_variables(const Transform& t) {
  std::vector<VarSpec> v;
  auto seen = [&](const std::string& n) { for (auto& s : v) if (s.name == n) return true; return false; };
  for (auto& op : t.ops) {
    if (op.type == OP_CONVT || op.type == OP_CONVT_RGB) {
      const ConvLayer& c = t.convs[op.conv];
      int cin = c.cin + (c.append_ones ? 1 : 0);
      for (auto& s : c.sources) {
        if (c.layout == LAYOUT_KERAS_OI) v.push_back({s.kernel, {c.k, c.k, s.cout, cin}});
        else if (c.layout == LAYOUT_D2S_1X1) v.push_back({s.kernel, {1, 1, cin / 4, s.cout}});
        else v.push_back({s.kernel, {c.k, c.k, cin, s.cout}});
        if (!s.bias.empty()) v.push_back({s.bias, {s.cout}});
      }
    }
    if (op.gdn >= 0) {
      const GdnLayer& g = t.gdns[op.gdn];
      if (!seen(g.beta)) { v.push_back({g.beta, {g.C}}); v.push_back({g.gamma, {g.C, g.C}}); }
    }
  }
  return v;
}

using HostWeights = std::map<std::string, std::pair<std::vector<int64_t>, std::vector<float>>>;

// W[ay, ax, co, ci] of a conv layer with concatenated sources, from the reference's native layouts.
inline float conv_w(const ConvLayer& c, const HostWeights& hw, int ay, int ax, int co, int ci) {
  if (c.layout == LAYOUT_BACKWARD) {   // see make_backward_conv: tap a' = T-1-d, input channel (ry, rx, co_f), output channel ci_f
    const ConvLayer& f = c.bwd_src[0];
    const int dy = c.k - 1 - ay, dx = c.k - 1 - ax;
    const int r = ci / f.cout, cof = ci - r * f.cout;
    if (r >= f.s * f.s) return 0.f;
    const int fay = f.s * dy + r / f.s, fax = f.s * dx + r % f.s;
    if (fay >= f.k || fax >= f.k) return 0.f;
    return conv_w(f, hw, fay, fax, cof, co);
  }
  if (c.layout == LAYOUT_COL2IM) {     // see make_col2im_conv: column (g, t), t = a_y * k + a_x < k*k
    const ConvLayer& f = c.bwd_src[0];
    const int g = co / c.c2i_kp, tt = co - g * c.c2i_kp;
    if (tt >= f.k * f.k) return 0.f;
    return conv_w(f, hw, tt / f.k, tt % f.k, g, ci);
  }
  int base = 0;
  for (auto& s : c.sources) {
    if (co < base + s.cout) {
      const auto& arr = hw.at(s.kernel).second;
      int cl = co - base;
      int cin = c.cin + (c.append_ones ? 1 : 0);
      if (c.layout == LAYOUT_D2S_1X1) {   // Conv2D kernel [1,1,Cin/4,Cout] seen through the preceding depth_to_space(2)
        const int cq = cin / 4;
        return ci / cq == 2 * ay + ax ? arr[(size_t)(ci % cq) * s.cout + cl] : 0.f;
      }
      size_t idx = c.layout == LAYOUT_KERAS_OI
                     ? (((size_t)ay * c.k + ax) * s.cout + cl) * cin + ci
                     : (((size_t)ay * c.k + ax) * cin + ci) * s.cout + cl;
      return arr[idx];
    }
    base += s.cout;
  }
  return 0.f;
}

// Packed band matrices: for band b, row kk = (jy*Tx + jx)*cin_pad + ci, column n = (fy*nphx + fx)*cout + co.
inline std::vector<float> pack_band_weights(const ConvLayer& c, const HostWeights& hw) {
  std::vector<float> w(c.w_floats, 0.f);
  int cin = c.cin + (c.append_ones ? 1 : 0);
  for (auto& b : c.bands) {
    for (int jy = 0; jy < b.Ty; ++jy) for (int jx = 0; jx < b.Tx; ++jx)
      for (int fy = 0; fy < b.nphy; ++fy) for (int fx = 0; fx < b.nphx; ++fx) {
        int ay = b.phy0 + fy + c.s * jy, ax = b.phx0 + fx + c.s * jx;
        if (ay >= c.k || ax >= c.k) continue;
        for (int ci = 0; ci < cin; ++ci) {
          size_t row = (size_t)(jy * b.Tx + jx) * c.cin_pad + ci;
          float* dst = &w[b.w_off + row * b.Npad + (size_t)(fy * b.nphx + fx) * c.cout];
          for (int co = 0; co < c.cout; ++co) dst[co] = conv_w(c, hw, ay, ax, co, ci);
        }
      }
  }
  return w;
}

inline std::vector<float> pack_bias(const ConvLayer& c, const HostWeights& hw) {
  std::vector<float> b(c.cout, 0.f);
  int base = 0;
  for (auto& s : c.sources) {
    if (!s.bias.empty()) {
      const auto& arr = hw.at(s.bias).second;
      for (int i = 0; i < s.cout; ++i) b[base + i] = arr[i];
    }
    base += s.cout;
  }
  return b;
}

// rgb-cell packing: [ay*k+ax][ci (cin_pad)][4] with co padded to 4
inline std::vector<float> pack_rgb_weights(const ConvLayer& c, const HostWeights& hw) {
  std::vector<float> w((size_t)c.k * c.k * c.cin_pad * 4, 0.f);
  int cin = c.cin + (c.append_ones ? 1 : 0);
  for (int ay = 0; ay < c.k; ++ay) for (int ax = 0; ax < c.k; ++ax) for (int ci = 0; ci < cin; ++ci)
    for (int co = 0; co < c.cout && co < 4; ++co)
      w[(((size_t)ay * c.k + ax) * c.cin_pad + ci) * 4 + co] = conv_w(c, hw, ay, ax, co, ci);
  return w;
}

}  // namespace sntc
