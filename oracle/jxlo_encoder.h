// ORACLE — test infrastructure only (see jxlo_bits.h header). PARITY UNPINNED.
// A minimal JPEG XL encoder so that .jxl inputs exist at all offline (SURVEY.md §7.1 step 2,
// Appendix A.11): VarDCT (XYB, fixed or heuristic AC strategies, optional CfL / adaptive
// quant / gaborish / EPF signalling) and lossless Modular (YCgCo RCT + gradient predictor,
// fixed MA tree). It is the CPU restatement of what the engine's SaveImage path produces,
// i.e. of the libjxl work behind JxlEncoderAddImageFrame / JxlEncoderFlushInput
// (N/Encoder/JxlEncoder.cpp:128,367) at the settings EncoderWriteImage applies (:207-335).
#pragma once
#include "jxlo_image.h"

namespace jxlo {
inline int64_t* LastEncodeStrategyCells() { static thread_local int64_t cells[27]; return cells; }   // cells covered per AC strategy in the last EncodeImage call of this thread

struct EncodeParams {
  float distance = 1.0f; int effort = 7; bool lossless = false;
  int gab = -1, epf = -1;                 // -1: derive from effort/distance
  int varblocks = -1, cfl = -1, adaptive_quant = -1;
  int force_strategy = -1;                // tile the frame with this AC strategy where it fits
  bool use_prefix = false; bool container = true; int modular_group_shift = 1; uint32_t orientation = 1; std::string frame_name;
  bool skip_lf_smoothing = false; int threads = 1;
  int varblock_pattern = 0;               // 1: modulate varblock_scale per 64x64 tile by {0, 0.55, 1, 2.75} (a fixed four-way mix of block sizes)
  float varblock_scale = 1.0f;            // multiplies the smoothness thresholds of the varblock merge (bench: a frame with a known share of large blocks)
  int num_passes = 1, pass_shift = 1;     // 2 passes: pass 0 carries the quantised coefficients >> pass_shift, pass 1 the remainder (progressive files)
  // source description for non-8-bit sources (tests of 16-bit / float / HDR output paths)
  BitDepth bd; ColorEncoding ce; float intensity_target = 255.f; bool premultiplied = false; bool black_channel = false;
  // Layers (multi-frame stills): the pixels are one frame of a canvas_w x canvas_h image, placed at (crop_x0, crop_y0) and blended onto reference
  // slot blend_source with blend_mode (colour) / alpha_blend_mode (alpha channel). frame_only: emit the frame alone (header, TOC, sections), to be
  // appended to a codestream whose image header was written by an earlier call with the same source description.
  uint32_t canvas_w = 0, canvas_h = 0; int32_t crop_x0 = 0, crop_y0 = 0; uint32_t blend_mode = 0, alpha_blend_mode = 0, blend_source = 0; bool blend_clamp = false;
  bool is_last = true; uint32_t save_as_reference = 0; bool frame_only = false;
};
struct EncodeInput {
  uint32_t width = 0, height = 0; int num_color = 3; bool has_alpha = false;
  const uint8_t* u8 = nullptr;     // interleaved Gray/GrayA/RGB/RGBA (what AddFrame hands to libjxl, N/Encoder/JxlEncoder.cpp:91-144)
  const float* f32 = nullptr;      // alternative: interleaved float samples in the encoding described by EncodeParams::ce (nominal range 0..1)
  const uint8_t* exif = nullptr; size_t exif_size = 0; const uint8_t* xmp = nullptr; size_t xmp_size = 0; const uint8_t* icc = nullptr; size_t icc_size = 0;
};

// ------------------------------------------------------------------ fixed MA trees (BFS layout)
struct TreeBuilder {
  struct Tmp { int prop; int32_t split; int l, r; int pred; }; std::vector<Tmp> n;
  int Leaf(int pred) { n.push_back({-1, 0, -1, -1, pred}); return int(n.size()) - 1; }
  int Split(int prop, int32_t val, int gt, int le) { n.push_back({prop, val, gt, le, 0}); return int(n.size()) - 1; }
  // balanced split on a single property; every leaf uses `pred`
  int Range(int prop, const std::vector<int32_t>& thr, int lo, int hi, int pred) { if (lo >= hi) return Leaf(pred); int mid = (lo + hi) / 2; int gt = Range(prop, thr, mid + 1, hi, pred); int le = Range(prop, thr, lo, mid, pred); return Split(prop, thr[mid], gt, le); }
  int ResidualCtx(int pred) { static const std::vector<int32_t> thr = {-64, -24, -8, -3, -1, 0, 2, 7, 23, 63}; return Range(8, thr, 0, int(thr.size()), pred); }
  int Channels(int nch, int first, int pred) { if (nch == 1) return ResidualCtx(pred); int mid = first + nch / 2 - 1; int gt = Channels(nch - nch / 2, mid + 1, pred); int le = Channels(nch / 2, first, pred); return Split(0, mid, gt, le); }
  Tree Flatten(int root) {
    Tree tree; std::vector<int> queue{root}; size_t head = 0; int leaf = 0;
    while (head < queue.size()) { const Tmp t = n[queue[head++]]; TreeNode o;
      if (t.prop < 0) { o.property = -1; o.predictor = t.pred; o.leaf_id = leaf++; } else { o.property = t.prop; o.splitval = t.split; o.lchild = int(queue.size()); o.rchild = int(queue.size()) + 1; queue.push_back(t.l); queue.push_back(t.r); }
      tree.push_back(o); }
    return tree;
  }
};
inline Tree MakeVarDctTree(uint32_t nlf, int num_ec) {
  TreeBuilder b; int sharp = b.Leaf(0), hfmul = b.Leaf(0), strat = b.Leaf(0), cflc = b.Leaf(0);   // Zero predictor: constant maps decode as zero-entropy rows
  int blockinfo = b.Split(2, 0, hfmul, strat); int hfmeta = b.Split(0, 1, b.Split(0, 2, sharp, blockinfo), cflc);
  int groups = num_ec > 0 ? b.Channels(num_ec, 0, 5) : b.Leaf(5);
  int upper = b.Split(1, int32_t(3 * nlf + 17), groups, hfmeta);
  int lfc = b.Channels(3, 0, 5); int global = b.Leaf(5); int lower = b.Split(1, 0, lfc, global);
  return b.Flatten(b.Split(1, int32_t(2 * nlf), upper, lower));
}
inline Tree MakeLosslessTree(int nch) { TreeBuilder b; return b.Flatten(b.Channels(nch, 0, 5)); }

// ------------------------------------------------------------------ colour (forward)
inline float SrgbToLinear(float v) { float a = std::fabs(v); float r = a <= 0.04045f ? a / 12.92f : std::pow((a + 0.055f) / 1.055f, 2.4f); return v < 0 ? -r : r; }
inline float TfToLinear(float v, const ColorEncoding& ce, float intensity_target) {
  float a = std::fabs(v), r;
  if (ce.have_gamma) r = std::pow(a, 1.0f / (float(ce.gamma) * 1e-7f));
  else switch (ce.tf) {
    case kTfLinear: r = a; break; case kTfSRGB: return SrgbToLinear(v);
    case kTf709: r = a < 0.081f ? a / 4.5f : std::pow((a + 0.099f) / 1.099f, 1.0f / 0.45f); break;
    case kTfPQ: { const double m1 = 2610.0 / 16384, m2 = 2523.0 / 4096 * 128, c1 = 3424.0 / 4096, c2 = 2413.0 / 4096 * 32, c3 = 2392.0 / 4096 * 32;
      double p = std::pow(double(a), 1.0 / m2); double num = std::max(p - c1, 0.0), den = c2 - c3 * p; r = float(std::pow(num / den, 1.0 / m1) * 10000.0 / intensity_target); break; }
    case kTfDCI: r = std::pow(a, 2.6f); break;
    case kTfHLG: { const double ha = 0.17883277, hb = 0.28466892, hc = 0.55991073; r = a <= 0.5f ? float(double(a) * a / 3.0) : float((std::exp((a - hc) / ha) + hb) / 12.0); break; }
    default: throw Error("unsupported source transfer function");
  }
  return v < 0 ? -r : r;
}
static const float kOpsinM[9] = {0.30f, 0.622f, 0.078f, 0.23f, 0.692f, 0.078f, 0.24342268924547819f, 0.20476744424496821f, 0.55180986650955360f};
static const float kOpsinBias = 0.0037930732552754493f;
inline void LinearToXyb(const float* rgb, float* xyb) {
  float g[3]; for (int c = 0; c < 3; c++) { float m = kOpsinM[3 * c] * rgb[0] + kOpsinM[3 * c + 1] * rgb[1] + kOpsinM[3 * c + 2] * rgb[2] + kOpsinBias; g[c] = std::cbrt(std::max(m, 0.f)) - std::cbrt(kOpsinBias); }
  xyb[0] = 0.5f * (g[0] - g[1]); xyb[1] = 0.5f * (g[0] + g[1]); xyb[2] = g[2];
}

struct VarDctPlan {
  int xb, yb, xpad, ypad; std::vector<uint8_t> strategy, is_first; std::vector<int32_t> hf_mul; std::vector<uint8_t> sharp; std::vector<int8_t> ytox, ytob; int xt, yt;
  Quantizer q; float lf_dequant[3] = {1.0f / 4096, 1.0f / 512, 1.0f / 256};
};

inline void QuantizerFromDistance(float d, Quantizer* q, float* q_ac) {
  d = std::max(d, 0.01f); float qac = 0.79f / d; float eff = 0.3f * std::pow(d / 0.3f, 0.83f); eff = std::min(d, std::max(0.5f * d, eff)); float qdc = std::min(50.0f, 1.0959f / eff);
  float scale = 65536.0f * qac / 5.0f; scale = std::min(32768.0f, std::max(1.0f, scale)); int gs = int(scale); int sdc = int(qdc * 4096.0f * 1.6f); if (gs > sdc) gs = std::max(1, sdc);
  q->global_scale = uint32_t(gs); float v = qdc * (65536.0f / float(gs)) + 0.5f; q->quant_lf = uint32_t(std::max(1.0f, std::min(65536.0f, v))); *q_ac = qac;
}

inline int32_t QuantizeCoef(float v, float inv_step) { float q = v * inv_step; if (std::fabs(q) < 0.56f) return 0; return int32_t(std::lrintf(q)); }

// ------------------------------------------------------------------ encoder proper
struct SectionWriter {   // mirrors SectionReaders: one shared writer when the TOC has a single entry
  bool single; std::vector<BitWriter> w; BitWriter shared;
  SectionWriter(size_t n) : single(n == 1), w(n == 1 ? 0 : n) {}
  BitWriter& Get(size_t i) { return single ? shared : w[i]; }
};

inline std::vector<uint8_t> EncodeImage(const EncodeInput& in, const EncodeParams& p) {
  JXLO_CHECK(in.width > 0 && in.height > 0 && (in.u8 || in.f32) && (in.num_color == 1 || in.num_color == 3), "encoder input");
  const int xs = int(in.width), ys = int(in.height), C = in.num_color + (in.has_alpha ? 1 : 0) + (p.black_channel ? 1 : 0);
  ImageMetadata m; m.xsize = in.width; m.ysize = in.height; m.orientation = p.orientation; m.xyb_encoded = !p.lossless; m.bd = p.bd; m.ce = p.ce; m.tm.intensity_target = p.intensity_target;
  if (in.num_color == 1) m.ce.color_space = kCsGray;
  if (in.icc_size) { m.ce.want_icc = true; m.icc.assign(in.icc, in.icc + in.icc_size); }
  if (p.black_channel) { ExtraChannelInfo k; k.type = kEcBlack; k.bd = p.bd; m.ec.push_back(k); }
  if (in.has_alpha) { ExtraChannelInfo a; a.type = kEcAlpha; a.bd = p.bd; a.alpha_associated = p.premultiplied; m.ec.push_back(a); }
  const int num_ec = int(m.ec.size());
  auto sample = [&](int x, int y, int c) -> float { size_t i = (size_t(y) * xs + x) * C + c; return in.u8 ? float(in.u8[i]) * (1.0f / 255.0f) : in.f32[i]; };
  // integer view of a sample for Modular coding (lossless colour, and extra channels always)
  auto isample = [&](int x, int y, int c) -> int32_t {
    size_t i = (size_t(y) * xs + x) * C + c; if (in.u8) return in.u8[i];
    if (p.bd.float_sample) {
      if (p.bd.bits == 16 && p.bd.exp_bits == 5) return int32_t(BitWriter::FloatToHalf(in.f32[i]));   // binary16: the half's bit pattern is the Modular sample
      JXLO_CHECK(p.bd.bits == 32 && p.bd.exp_bits == 8, "float lossless source must be binary32 or binary16"); int32_t v; memcpy(&v, &in.f32[i], 4); return v; }
    return int32_t(std::lrintf(std::min(1.f, std::max(0.f, in.f32[i])) * float((1u << p.bd.bits) - 1)));
  };
  if (in.u8) JXLO_CHECK(!p.bd.float_sample && p.bd.bits == 8, "u8 input requires an 8-bit stream");

  FrameHeader fh; fh.encoding = p.lossless ? 1 : 0; fh.name = p.frame_name; fh.ec_upsampling.assign(num_ec, 1); fh.ec_blending.assign(num_ec, BlendingInfo());
  if (p.canvas_w || p.canvas_h) {   // one layer of a larger (or equal) canvas
    m.xsize = p.canvas_w; m.ysize = p.canvas_h; fh.have_crop = true; fh.x0 = p.crop_x0; fh.y0 = p.crop_y0; fh.width = in.width; fh.height = in.height;
  }
  fh.is_last = p.is_last; fh.save_as_reference = p.is_last ? 0 : p.save_as_reference;
  { const int alpha_ec = m.alpha_index(); fh.blending.mode = p.blend_mode; fh.blending.source = p.blend_source; fh.blending.clamp = p.blend_clamp; fh.blending.alpha_channel = alpha_ec >= 0 ? uint32_t(alpha_ec) : 0;
    JXLO_CHECK(!(p.blend_mode == 2 || p.blend_mode == 3) || alpha_ec >= 0, "alpha blending needs an alpha channel");
    for (int i = 0; i < num_ec; i++) { BlendingInfo& b = fh.ec_blending[i]; b.mode = i == alpha_ec ? p.alpha_blend_mode : p.blend_mode; b.source = p.blend_source; b.clamp = p.blend_clamp; b.alpha_channel = alpha_ec >= 0 ? uint32_t(alpha_ec) : 0; } }
  if (p.lossless) { fh.group_size_shift = uint32_t(p.modular_group_shift); fh.lf.gab = false; fh.lf.epf_iters = 0; }
  else {
    int gab = p.gab >= 0 ? p.gab : (p.effort >= 5 ? 1 : 0); int epf = p.epf;
    if (epf < 0) { epf = 0; if (p.effort >= 5) { const float thr[3] = {0.7f, 1.5f, 4.0f}; for (float t : thr) if (p.distance >= t) epf++; } }
    fh.lf.gab = gab != 0; fh.lf.epf_iters = uint32_t(epf); if (p.skip_lf_smoothing) fh.flags |= kFlagSkipAdaptiveLfSmoothing;
  }
  JXLO_CHECK(p.num_passes == 1 || p.num_passes == 2, "num_passes"); fh.passes.num_passes = uint32_t(p.num_passes); if (p.num_passes == 2) fh.passes.shift[0] = uint32_t(p.pass_shift);
  const uint32_t np = fh.passes.num_passes;
  DeriveFrameDims(fh, m);
  const uint32_t nlf = fh.num_lf_groups, ng = fh.num_groups; const size_t nsec = NumTocEntries(fh);
  SectionWriter sw(nsec);
  Tree tree = p.lossless ? MakeLosslessTree(in.num_color + num_ec) : MakeVarDctTree(nlf, num_ec);
  EncOptions mopt; mopt.cfg = HybridCfg{4, 1, 0}; mopt.use_prefix = p.use_prefix; mopt.max_clusters = 48;
  std::vector<Token> tree_tokens; TokenizeTree(tree, &tree_tokens); EncOptions topt; topt.cfg = HybridCfg{4, 1, 0}; topt.use_prefix = p.use_prefix;

  // ---- global modular image (colour for lossless, extra channels always)
  ModularImage gimg; gimg.bitdepth = int(p.bd.bits); GroupHeader gheader; gheader.use_global_tree = true;
  if (p.lossless) for (int c = 0; c < in.num_color; c++) { Channel ch(xs, ys); for (int y = 0; y < ys; y++) for (int x = 0; x < xs; x++) ch.row(y)[x] = isample(x, y, c); gimg.ch.push_back(std::move(ch)); }
  for (int e = 0; e < num_ec; e++) { Channel ch(xs, ys); int src = in.num_color + e; for (int y = 0; y < ys; y++) for (int x = 0; x < xs; x++) ch.row(y)[x] = isample(x, y, src); gimg.ch.push_back(std::move(ch)); }
  if (p.lossless && in.num_color == 3) {
    Transform t; t.id = 0; t.begin_c = 0; t.rct_type = 6; gheader.transforms.push_back(t);
    for (size_t i = 0; i < gimg.ch[0].d.size(); i++) { int32_t Y, Co, Cg; ForwardRCT_YCgCo(gimg.ch[0].d[i], gimg.ch[1].d[i], gimg.ch[2].d[i], &Y, &Co, &Cg); gimg.ch[0].d[i] = Y; gimg.ch[1].d[i] = Co; gimg.ch[2].d[i] = Cg; }
  }
  const int gd = int(fh.group_dim); size_t global_n = 0; for (; global_n < gimg.ch.size(); global_n++) if (gimg.ch[global_n].w > gd || gimg.ch[global_n].h > gd) break;
  std::vector<Token> global_tokens; if (global_n) ModularTokenize(gimg, 0, global_n, 0, tree, &global_tokens);
  // group-local modular streams (pass 0 carries everything: single pass)
  std::vector<std::vector<Token>> mg_tokens(ng); std::vector<uint8_t> mg_present(ng, 0);
  ParallelFor(ng, p.threads, [&](size_t g) {
    int gx = int(g % fh.xgroups), gy = int(g / fh.xgroups); ModularImage gi; gi.bitdepth = gimg.bitdepth;
    for (size_t c = global_n; c < gimg.ch.size(); c++) { const Channel& fc = gimg.ch[c]; int x0 = gx * gd, y0 = gy * gd; if (x0 >= fc.w || y0 >= fc.h) continue; int w = std::min(gd, fc.w - x0), h = std::min(gd, fc.h - y0);
      Channel ch(w, h); for (int y = 0; y < h; y++) memcpy(ch.row(y), fc.row(y0 + y) + x0, sizeof(int32_t) * size_t(w)); gi.ch.push_back(std::move(ch)); }
    if (gi.ch.empty()) return; mg_present[g] = 1; ModularTokenize(gi, 0, gi.ch.size(), StreamIdModularGroup(fh, np - 1, uint32_t(g)), tree, &mg_tokens[g]);   // shift-0 channels belong to the last pass
  });

  // ---- VarDCT analysis
  VarDctPlan pl; std::vector<std::vector<Token>> lf_tokens(nlf), hfmeta_tokens(nlf), ac_tokens(size_t(ng) * np); std::vector<uint32_t> hfmeta_nb(nlf, 0);
  if (!p.lossless) {
    pl.xb = int(fh.xblocks); pl.yb = int(fh.yblocks); pl.xpad = pl.xb * 8; pl.ypad = pl.yb * 8; pl.xt = (pl.xb + 7) / 8; pl.yt = (pl.yb + 7) / 8;
    Plane xyb[3]; for (auto& pln : xyb) pln = Plane(pl.xpad, pl.ypad);
    float to_srgb_lin[9]; { ColorEncoding src = m.ce; src.want_icc = false; float fwd[9]; LinearSrgbToTarget(src, fwd); double a[9], ai[9]; for (int i = 0; i < 9; i++) a[i] = fwd[i]; Inv3x3(a, ai); for (int i = 0; i < 9; i++) to_srgb_lin[i] = float(ai[i]); }
    bool src_srgb = m.ce.IsDefault() || (m.ce.tf == kTfSRGB && m.ce.primaries == kPrSRGB && m.ce.white_point == kWpD65 && !m.ce.have_gamma) || m.ce.want_icc;
    float lut[256]; for (int i = 0; i < 256; i++) lut[i] = SrgbToLinear(float(i) * (1.0f / 255.0f));
    float it_mul = p.intensity_target / 255.0f;
    ParallelFor(size_t(pl.ypad), p.threads, [&](size_t yy) {
      int y = std::min(int(yy), ys - 1);
      for (int xx = 0; xx < pl.xpad; xx++) { int x = std::min(xx, xs - 1); float rgb[3], lin[3], o[3];
        for (int c = 0; c < 3; c++) { int sc = in.num_color == 1 ? 0 : c; if (in.u8 && src_srgb) lin[c] = lut[in.u8[(size_t(y) * xs + x) * C + sc]]; else { rgb[c] = sample(x, y, sc); lin[c] = src_srgb ? SrgbToLinear(rgb[c]) : TfToLinear(rgb[c], m.ce, p.intensity_target); } }
        if (in.has_alpha && p.premultiplied) { /* colour is stored premultiplied as given */ }
        if (!src_srgb) { float t[3] = {lin[0], lin[1], lin[2]}; for (int c = 0; c < 3; c++) lin[c] = to_srgb_lin[3 * c] * t[0] + to_srgb_lin[3 * c + 1] * t[1] + to_srgb_lin[3 * c + 2] * t[2]; }
        for (float& v : lin) v *= it_mul;
        LinearToXyb(lin, o); for (int c = 0; c < 3; c++) xyb[c].row(int(yy))[xx] = o[c]; }
    });
    if (fh.lf.gab) {   // approximate inverse gaborish: two Van Cittert iterations against the decoder's kernel
      Plane orig[3] = {xyb[0], xyb[1], xyb[2]};
      for (int it = 0; it < 2; it++) { Plane blur[3] = {xyb[0], xyb[1], xyb[2]}; Gaborish(blur, pl.xpad, pl.ypad, fh.lf); for (int c = 0; c < 3; c++) for (size_t i = 0; i < xyb[c].d.size(); i++) xyb[c].d[i] += orig[c].d[i] - blur[c].d[i]; }
    }
    float q_ac; QuantizerFromDistance(p.distance, &pl.q, &q_ac); const float inv_gs = pl.q.InvGlobalScale();
    size_t ncell = size_t(pl.xb) * pl.yb; pl.strategy.assign(ncell, kDCT8); pl.is_first.assign(ncell, 1); pl.sharp.assign(ncell, fh.lf.epf_iters ? 4 : 0); pl.ytox.assign(size_t(pl.xt) * pl.yt, 0); pl.ytob.assign(size_t(pl.xt) * pl.yt, 0);
    int base_qf = std::max(1, std::min(255, int(std::lrintf(q_ac * 65536.0f / float(pl.q.global_scale))))); pl.hf_mul.assign(ncell, base_qf);
    // per-cell activity from the DCT8 of Y (drives varblock merging and adaptive quant)
    std::vector<float> act(ncell, 0.f);
    bool varblocks = p.varblocks >= 0 ? p.varblocks != 0 : p.effort >= 5; bool aq = p.adaptive_quant >= 0 ? p.adaptive_quant != 0 : p.effort >= 5; bool cfl = p.cfl >= 0 ? p.cfl != 0 : p.effort >= 5;
    if (varblocks || aq) ParallelFor(size_t(pl.yb), p.threads, [&](size_t by) { float co[64]; for (int bx = 0; bx < pl.xb; bx++) { TransformFromPixels(kDCT8, xyb[1].row(int(by) * 8) + bx * 8, size_t(pl.xpad), co); float s = 0; for (int k = 1; k < 64; k++) s += std::fabs(co[k]); act[by * pl.xb + bx] = s; } });
    auto place = [&](int by, int bx, int s) { int bw = kCoveredX[s], bh = kCoveredY[s]; for (int iy = 0; iy < bh; iy++) for (int ix = 0; ix < bw; ix++) { size_t o = size_t(by + iy) * pl.xb + bx + ix; pl.strategy[o] = uint8_t(s); pl.is_first[o] = 0; } pl.is_first[size_t(by) * pl.xb + bx] = 1; };
    if (p.force_strategy >= 0) {
      int s = p.force_strategy, bw = kCoveredX[s], bh = kCoveredY[s];
      for (int by = 0; by + bh <= pl.yb; by += bh) for (int bx = 0; bx + bw <= pl.xb; bx += bw) place(by, bx, s);
    } else if (varblocks) {
      float step_y = (1.0f / 560.0f) * inv_gs / float(base_qf);
      auto smooth = [&](int by, int bx, int n, float thr) { if (by + n > pl.yb || bx + n > pl.xb) return false; for (int iy = 0; iy < n; iy++) for (int ix = 0; ix < n; ix++) if (act[size_t(by + iy) * pl.xb + bx + ix] > thr * step_y) return false; return true; };
      for (int by = 0; by < pl.yb; by += 4) for (int bx = 0; bx < pl.xb; bx += 4) {
        static const float kPat[4] = {0.f, 0.55f, 1.0f, 2.75f}; const float vs = p.varblock_scale * (p.varblock_pattern ? kPat[((bx / 8) + 2 * (by / 8)) & 3] : 1.0f);
        if (smooth(by, bx, 4, 6.0f * vs)) { place(by, bx, kDCT32); continue; }
        for (int sy = 0; sy < 4; sy += 2) for (int sx = 0; sx < 4; sx += 2) {
          int y0 = by + sy, x0 = bx + sx; if (y0 >= pl.yb || x0 >= pl.xb) continue;
          if (smooth(y0, x0, 2, 14.0f * vs)) { place(y0, x0, kDCT16); continue; }
          if (x0 + 1 < pl.xb) { for (int r = 0; r < 2 && y0 + r < pl.yb; r++) if (act[size_t(y0 + r) * pl.xb + x0] < 24.0f * vs * step_y && act[size_t(y0 + r) * pl.xb + x0 + 1] < 24.0f * vs * step_y) place(y0 + r, x0, kDCT8x16); }
        }
      }
    }
    if (aq) {   // masking-style modulation, constant inside a varblock
      double mean = 0; for (float a : act) mean += std::log(a + 1e-4); mean = std::exp(mean / double(ncell));
      for (int by = 0; by < pl.yb; by++) for (int bx = 0; bx < pl.xb; bx++) { size_t o = size_t(by) * pl.xb + bx; if (!pl.is_first[o]) continue; int s = pl.strategy[o]; float a = 0; int n = 0;
        for (int iy = 0; iy < kCoveredY[s]; iy++) for (int ix = 0; ix < kCoveredX[s]; ix++) { a += act[o + size_t(iy) * pl.xb + ix]; n++; }
        float mulq = std::pow(float(mean) / (a / n + 1e-4f), 0.2f); mulq = std::min(1.35f, std::max(0.75f, mulq)); int qf = std::max(1, std::min(255, int(std::lrintf(base_qf * mulq))));
        for (int iy = 0; iy < kCoveredY[s]; iy++) for (int ix = 0; ix < kCoveredX[s]; ix++) pl.hf_mul[o + size_t(iy) * pl.xb + ix] = qf; }
    }
    { int64_t* st = LastEncodeStrategyCells(); for (int i = 0; i < 27; i++) st[i] = 0; for (size_t i = 0; i < ncell; i++) st[pl.strategy[i]]++; }
    // dequant tables + natural orders
    std::vector<std::vector<float>> dequant(kNumQuantTables); for (int t = 0; t < kNumQuantTables; t++) dequant[t] = ComputeDequantTable(t, LibraryEncoding(t));
    std::vector<std::vector<uint32_t>> natural(kNumOrders); for (int o = 0; o < kNumOrders; o++) { int s = kOrderStrategy[o]; natural[o] = NaturalOrder(std::min(kCoveredX[s], kCoveredY[s]), std::max(kCoveredX[s], kCoveredY[s])); }
    const float xm = std::pow(0.8f, float(fh.x_qm_scale) - 2.0f), bm = std::pow(0.8f, float(fh.b_qm_scale) - 2.0f); OpsinInverse opsin; const float* qbias = opsin.quant_bias;
    Plane lf[3]; for (auto& pln : lf) pln = Plane(pl.xb, pl.yb);
    BlockCtxMap bctx; const uint32_t nbctx = bctx.num_ctxs;
    ParallelFor(ng, p.threads, [&](size_t g) {
      int gx = int(g % fh.xgroups), gy = int(g / fh.xgroups), cx0 = gx * 32, cy0 = gy * 32, w = std::min(32, pl.xb - cx0), h = std::min(32, pl.yb - cy0);
      struct Blk { int by, bx, s; std::vector<float> co[3]; }; std::vector<Blk> blocks;
      for (int by = 0; by < h; by++) for (int bx = 0; bx < w; bx++) { size_t o = size_t(cy0 + by) * pl.xb + cx0 + bx; if (!pl.is_first[o]) continue; Blk b; b.by = by; b.bx = bx; b.s = pl.strategy[o]; size_t size = size_t(kCoveredX[b.s]) * kCoveredY[b.s] * 64;
        for (int c = 0; c < 3; c++) { b.co[c].resize(size); TransformFromPixels(b.s, xyb[c].row((cy0 + by) * 8) + (cx0 + bx) * 8, size_t(pl.xpad), b.co[c].data()); DCFromLowestFrequencies(b.s, b.co[c].data(), lf[c].row(cy0 + by) + cx0 + bx, size_t(pl.xb)); }
        blocks.push_back(std::move(b)); }
      if (cfl) {   // least-squares chroma-from-luma factor per 64x64 tile (blocks vote with their top-left tile)
        for (int ty = cy0 / 8; ty <= (cy0 + h - 1) / 8; ty++) for (int tx = cx0 / 8; tx <= (cx0 + w - 1) / 8; tx++) { double sxy = 0, sby = 0, syy = 0;
          for (const Blk& b : blocks) { if ((cy0 + b.by) / 8 != ty || (cx0 + b.bx) / 8 != tx) continue; size_t llf = size_t(kCoveredX[b.s]) * kCoveredY[b.s]; int SW = std::max(kCoveredX[b.s], kCoveredY[b.s]) * 8, lr = std::min(kCoveredX[b.s], kCoveredY[b.s]), lc = std::max(kCoveredX[b.s], kCoveredY[b.s]); (void)llf;
            for (size_t k = 0; k < b.co[1].size(); k++) { int r = int(k) / SW, cc = int(k) % SW; if (r < lr && cc < lc) continue; double yv = b.co[1][k]; sxy += yv * b.co[0][k]; sby += yv * b.co[2][k]; syy += yv * yv; } }
          int fx = 0, fb = 0; if (syy > 1e-12) { fx = int(std::lrint(84.0 * (sxy / syy))); fb = int(std::lrint(84.0 * (sby / syy - 1.0))); }
          pl.ytox[size_t(ty) * pl.xt + tx] = int8_t(std::max(-128, std::min(127, fx))); pl.ytob[size_t(ty) * pl.xt + tx] = int8_t(std::max(-128, std::min(127, fb))); }
      }
      ColorCorrelation cc; std::vector<uint16_t> nzs_all[2][3]; for (auto& a : nzs_all) for (auto& v : a) v.assign(32 * 32, 0);
      for (const Blk& b : blocks) {
        size_t o = size_t(cy0 + b.by) * pl.xb + cx0 + b.bx; int s = b.s, bw = kCoveredX[s], bh = kCoveredY[s]; uint32_t covered = uint32_t(bw * bh), log2c = uint32_t(FloorLog2(covered)), size = covered * 64; int ord = kStrategyOrder[s], t = kQuantTableOf[s];
        int SW = std::max(bw, bh) * 8, lr = std::min(bw, bh), lc = std::max(bw, bh);
        const float* dq = dequant[t].data(); float scale = inv_gs / float(pl.hf_mul[o]); size_t tile = size_t((cy0 + b.by) / 8) * pl.xt + (cx0 + b.bx) / 8; float kx = cc.YtoX(pl.ytox[tile]), kb = cc.YtoB(pl.ytob[tile]);
        std::vector<int32_t> q[3]; for (auto& v : q) v.assign(size, 0);
        for (uint32_t k = 0; k < size; k++) { int r = int(k) / SW, c2 = int(k) % SW; if (r < lr && c2 < lc) continue;
          float sy = dq[size + k] * scale; q[1][k] = QuantizeCoef(b.co[1][k], 1.0f / sy); float ydq = AdjustQuantBias(q[1][k], qbias[1], qbias[3]) * sy;
          q[0][k] = QuantizeCoef(b.co[0][k] - kx * ydq, 1.0f / (dq[k] * scale * xm)); q[2][k] = QuantizeCoef(b.co[2][k] - kb * ydq, 1.0f / (dq[2 * size + k] * scale * bm)); }
        std::vector<int32_t> qfull[3] = {q[0], q[1], q[2]};
        for (uint32_t pass = 0; pass < np; pass++) { std::vector<Token>& out = ac_tokens[size_t(pass) * ng + g]; std::vector<uint16_t>* nzs = nzs_all[pass];
        if (np == 2) for (int c = 0; c < 3; c++) for (uint32_t k = 0; k < size; k++) { int32_t hi = qfull[c][k] >> p.pass_shift; q[c][k] = pass == 0 ? hi : qfull[c][k] - (hi << p.pass_shift); }
        for (int ci = 0; ci < 3; ci++) { int c = ci == 0 ? 1 : ci == 1 ? 0 : 2; const std::vector<uint32_t>& order = natural[ord];
          uint32_t nz = 0; for (uint32_t k = covered; k < size; k++) nz += q[c][order[k]] != 0;
          uint32_t pred; { uint16_t* z = nzs[c].data(); int by = b.by, bx = b.bx; if (bx == 0) pred = by == 0 ? 32 : z[(by - 1) * 32 + bx]; else if (by == 0) pred = z[by * 32 + bx - 1]; else pred = (uint32_t(z[(by - 1) * 32 + bx]) + z[by * 32 + bx - 1] + 1) / 2; }
          uint32_t bc = bctx.Context(0, uint32_t(pl.hf_mul[o]), uint32_t(ord), uint32_t(c)); out.push_back({NonZeroCtxBucket(pred) * nbctx + bc, nz});
          { uint16_t v = uint16_t((nz + covered - 1) >> log2c); for (int iy = 0; iy < bh; iy++) for (int ix = 0; ix < bw; ix++) nzs[c][(b.by + iy) * 32 + b.bx + ix] = v; }
          uint32_t histo = nbctx * kNonZeroBuckets + kZeroDensityContextCount * bc, prev = nz > size / 16 ? 0 : 1;
          for (uint32_t k = covered; k < size && nz != 0; k++) { int32_t v = q[c][order[k]]; out.push_back({histo + ZeroDensityContext(nz, k, covered, log2c, prev), PackSigned(v)}); prev = v != 0; nz -= prev; } }
        }
      }
    });
    // LF quantisation + tokens, HF metadata tokens (per LF group)
    float lfinv = inv_gs / float(pl.q.quant_lf); ColorCorrelation cc;
    ParallelFor(nlf, p.threads, [&](size_t g) {
      int gx = int(g % fh.xlfgroups), gy = int(g / fh.xlfgroups), cx0 = gx * 256, cy0 = gy * 256, w = std::min(256, pl.xb - cx0), h = std::min(256, pl.yb - cy0);
      ModularImage img; img.bitdepth = 16; for (int c = 0; c < 3; c++) img.ch.push_back(Channel(w, h));
      for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { size_t o = size_t(cy0 + y) * pl.xb + cx0 + x;
        float fy = pl.lf_dequant[1] * lfinv; int32_t qy = int32_t(std::lrintf(lf[1].d[o] / fy)); float ydq = float(qy) * fy;
        int32_t qx = int32_t(std::lrintf((lf[0].d[o] - cc.YtoX(0) * ydq) / (pl.lf_dequant[0] * lfinv))); int32_t qb = int32_t(std::lrintf((lf[2].d[o] - cc.YtoB(0) * ydq) / (pl.lf_dequant[2] * lfinv)));
        img.ch[0].row(y)[x] = qy; img.ch[1].row(y)[x] = qx; img.ch[2].row(y)[x] = qb; }
      ModularTokenize(img, 0, 3, StreamIdLfCoeff(fh, uint32_t(g)), tree, &lf_tokens[g]);
      uint32_t nb = 0; for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) nb += pl.is_first[size_t(cy0 + y) * pl.xb + cx0 + x]; hfmeta_nb[g] = nb;
      int tw = (w + 7) / 8, th = (h + 7) / 8; ModularImage hm; hm.bitdepth = 8; hm.ch.push_back(Channel(tw, th, 3, 3)); hm.ch.push_back(Channel(tw, th, 3, 3)); hm.ch.push_back(Channel(int(nb), 2)); hm.ch.push_back(Channel(w, h));
      for (int y = 0; y < th; y++) for (int x = 0; x < tw; x++) { size_t o = size_t(cy0 / 8 + y) * pl.xt + cx0 / 8 + x; hm.ch[0].row(y)[x] = pl.ytox[o]; hm.ch[1].row(y)[x] = pl.ytob[o]; }
      uint32_t k = 0; for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { size_t o = size_t(cy0 + y) * pl.xb + cx0 + x; if (!pl.is_first[o]) continue; hm.ch[2].row(0)[k] = pl.strategy[o]; hm.ch[2].row(1)[k] = pl.hf_mul[o] - 1; k++; }
      for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) hm.ch[3].row(y)[x] = pl.sharp[size_t(cy0 + y) * pl.xb + cx0 + x];
      ModularTokenize(hm, 0, 4, StreamIdHfMeta(fh, uint32_t(g)), tree, &hfmeta_tokens[g]);
    });
  }

  // ---- entropy codes
  std::vector<const std::vector<Token>*> mstreams; mstreams.push_back(&global_tokens); for (auto& t : lf_tokens) mstreams.push_back(&t); for (auto& t : hfmeta_tokens) mstreams.push_back(&t); for (auto& t : mg_tokens) mstreams.push_back(&t);
  EncCode tree_code = BuildCode({&tree_tokens}, 6, topt); EncCode mcode = BuildCode(mstreams, NumLeaves(tree), mopt);
  EncCode ac_code[2]; BlockCtxMap bctx0;
  if (!p.lossless) for (uint32_t pass = 0; pass < np; pass++) { std::vector<const std::vector<Token>*> as; for (uint32_t g = 0; g < ng; g++) as.push_back(&ac_tokens[size_t(pass) * ng + g]); EncOptions aopt; aopt.cfg = HybridCfg{4, 2, 0}; aopt.use_prefix = p.use_prefix; aopt.max_clusters = 64; ac_code[pass] = BuildCode(as, size_t(495) * bctx0.num_ctxs, aopt); }

  // ---- sections
  { BitWriter& bw = sw.Get(0);   // LfGlobal
    bw.Bool(true);   /* LfChannelDequantization all_default: present for Modular frames too */
    if (!p.lossless) { bw.U32(BitsOffset(11, 1), BitsOffset(11, 2049), BitsOffset(12, 4097), BitsOffset(16, 8193), pl.q.global_scale); bw.U32(Val(16), BitsOffset(5, 1), BitsOffset(8, 1), BitsOffset(16, 1), pl.q.quant_lf); bw.Bool(true); bw.Bool(true); }
    bw.Bool(true); WriteCode(bw, tree_code); WriteTokens(bw, tree_code, tree_tokens); WriteCode(bw, mcode);
    if (!gimg.ch.empty()) { WriteGroupHeader(bw, gheader); if (global_n) WriteTokens(bw, mcode, global_tokens); } }
  GroupHeader plain; plain.use_global_tree = true;
  for (uint32_t g = 0; g < nlf; g++) { BitWriter& bw = sw.Get(1 + g);
    if (!p.lossless) { bw.Write(2, 0); WriteGroupHeader(bw, plain); WriteTokens(bw, mcode, lf_tokens[g]);
      int gx = int(g % fh.xlfgroups), gy = int(g / fh.xlfgroups), w = std::min(256, pl.xb - gx * 256), h = std::min(256, pl.yb - gy * 256);
      bw.Write(CeilLog2(uint64_t(w) * h), hfmeta_nb[g] - 1); WriteGroupHeader(bw, plain); WriteTokens(bw, mcode, hfmeta_tokens[g]); } }
  { BitWriter& bw = sw.Get(1 + nlf);   // HfGlobal
    if (!p.lossless) { bw.Bool(true); bw.Write(CeilLog2(ng), 0); for (uint32_t pass = 0; pass < np; pass++) { bw.U32(Val(0x5F), Val(0x13), Val(0), Bits(13), 0); WriteCode(bw, ac_code[pass]); } } }
  for (uint32_t pass = 0; pass < np; pass++) for (uint32_t g = 0; g < ng; g++) { BitWriter& bw = sw.Get(2 + nlf + size_t(pass) * ng + g);
    if (!p.lossless) WriteTokens(bw, ac_code[pass], ac_tokens[size_t(pass) * ng + g]);
    if (mg_present[g] && pass + 1 == np) { WriteGroupHeader(bw, plain); WriteTokens(bw, mcode, mg_tokens[g]); } }

  // ---- assemble codestream
  BitWriter cs; if (!p.frame_only) { cs.Write(16, 0x0AFF); WriteImageHeaders(cs, m); } WriteFrameHeader(cs, fh, m);
  std::vector<std::vector<uint8_t>> secs; std::vector<size_t> sizes;
  if (sw.single) { secs.push_back(sw.shared.Finish()); } else for (auto& w : sw.w) secs.push_back(w.Finish());
  for (auto& s : secs) sizes.push_back(s.size());
  WriteToc(cs, sizes); std::vector<uint8_t> out = cs.Finish(); for (auto& s : secs) out.insert(out.end(), s.begin(), s.end());
  if (!p.container || p.frame_only) return out;
  // container: always used by the reference (JxlEncoderUseBoxes, N/Encoder/JxlEncoder.cpp:201); Exif / xml boxes uncompressed (:284-310)
  std::vector<uint8_t> file = ContainerPrologue();
  if (in.exif_size) AppendBox(file, "Exif", in.exif, in.exif_size);   // blob already carries the 4-byte TIFF offset (S/Exif/ExifWriter.cs:78-90)
  if (in.xmp_size) AppendBox(file, "xml ", in.xmp, in.xmp_size);
  AppendBox(file, "jxlc", out.data(), out.size());
  return file;
}

}  // namespace jxlo
