// ORACLE — test infrastructure only (see jxlo_bits.h header). PARITY UNPINNED.
// Codestream headers: SizeHeader, ImageMetadata, CustomTransformData, FrameHeader, TOC,
// and the ISO/IEC 18181-2 container boxes. Restates SURVEY.md Appendix A.1, A.3-A.5.
// The reference sees these only as JxlBasicInfo / JxlColorEncoding / JxlFrameHeader values
// (N/Decoder/JxlDecoder.cpp:461-561, 562-686, 259-288) and box events (:687-784).
#pragma once
#include "jxlo_bits.h"
#include "jxlo_entropy.h"

namespace jxlo {

struct BitDepth { bool float_sample = false; uint32_t bits = 8, exp_bits = 0; };
struct ExtraChannelInfo { uint32_t type = 0; BitDepth bd; uint32_t dim_shift = 0; std::string name; bool alpha_associated = false; float spot[4] = {0, 0, 0, 0}; uint32_t cfa = 1; };
enum { kEcAlpha = 0, kEcDepth = 1, kEcSpot = 2, kEcSelection = 3, kEcBlack = 4, kEcCFA = 5, kEcThermal = 6 };
enum { kCsRGB = 0, kCsGray = 1, kCsXYB = 2, kCsUnknown = 3 };
enum { kWpD65 = 1, kWpCustom = 2, kWpE = 10, kWpDCI = 11 };
enum { kPrSRGB = 1, kPrCustom = 2, kPr2100 = 9, kPrP3 = 11 };
enum { kTf709 = 1, kTfUnknown = 2, kTfLinear = 8, kTfSRGB = 13, kTfPQ = 16, kTfDCI = 17, kTfHLG = 18 };
struct ColorEncoding {
  bool want_icc = false; uint32_t color_space = kCsRGB, white_point = kWpD65, primaries = kPrSRGB; bool have_gamma = false; uint32_t gamma = 0;
  uint32_t tf = kTfSRGB, intent = 1; int32_t white_xy[2] = {0, 0}; int32_t prim_xy[3][2] = {{0, 0}, {0, 0}, {0, 0}};   // custom xy in 1e-6 units
  bool IsDefault() const { return !want_icc && color_space == kCsRGB && white_point == kWpD65 && primaries == kPrSRGB && !have_gamma && tf == kTfSRGB && intent == 1; }
};
struct ToneMapping { float intensity_target = 255.f, min_nits = 0.f; bool relative_to_max_display = false; float linear_below = 0.f; };
struct OpsinInverse {
  float inv[9] = {11.031566901960783f, -9.866943921568629f, -0.16462299647058826f, -3.254147380392157f, 4.418770392156863f, -0.16462299647058826f,
                  -3.6588512862745097f, 2.7129230470588235f, 1.9459282392156863f};
  float bias[3] = {-0.0037930732552754493f, -0.0037930732552754493f, -0.0037930732552754493f};
  float quant_bias[4] = {1.0f - 0.05465007330715401f, 1.0f - 0.07005449891748593f, 1.0f - 0.049935103337343655f, 0.145f};
};
struct Animation { uint32_t tps_num = 100, tps_den = 1, loops = 0; bool have_timecodes = false; };
struct ImageMetadata {
  uint32_t xsize = 0, ysize = 0;
  uint32_t orientation = 1; bool have_intrinsic = false, have_preview = false, have_animation = false; Animation anim;
  uint32_t intrinsic_x = 0, intrinsic_y = 0, preview_x = 0, preview_y = 0;
  BitDepth bd; bool modular_16bit = true; std::vector<ExtraChannelInfo> ec; bool xyb_encoded = true; ColorEncoding ce; ToneMapping tm;
  OpsinInverse opsin; uint32_t cw_mask = 0; std::vector<float> up2, up4, up8;
  std::vector<uint8_t> icc;   // decoded ICC profile when ce.want_icc
  int num_color_channels() const { return ce.color_space == kCsGray ? 1 : 3; }
  int alpha_index() const { for (size_t i = 0; i < ec.size(); i++) if (ec[i].type == kEcAlpha) return int(i); return -1; }
  int black_index() const { for (size_t i = 0; i < ec.size(); i++) if (ec[i].type == kEcBlack) return int(i); return -1; }
};

inline void ReadSize(BitReader& br, uint32_t* xs, uint32_t* ys) {
  bool small = br.Bool(); uint32_t y = small ? (br.ReadBits(5) + 1) * 8 : br.U32(BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  uint32_t ratio = br.ReadBits(3), x;
  if (ratio == 0) x = small ? (br.ReadBits(5) + 1) * 8 : br.U32(BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  else { static const uint32_t num[8] = {0, 1, 12, 4, 3, 16, 5, 2}, den[8] = {1, 1, 10, 3, 2, 9, 4, 1}; x = uint32_t(uint64_t(y) * num[ratio] / den[ratio]); }
  *xs = x; *ys = y;
}
inline void WriteSize(BitWriter& bw, uint32_t xs, uint32_t ys) {
  bool small = xs % 8 == 0 && ys % 8 == 0 && xs <= 256 && ys <= 256; bw.Bool(small);
  if (small) bw.Write(5, ys / 8 - 1); else bw.U32(BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1), ys);
  static const uint32_t num[8] = {0, 1, 12, 4, 3, 16, 5, 2}, den[8] = {1, 1, 10, 3, 2, 9, 4, 1}; uint32_t ratio = 0;
  for (uint32_t r = 1; r < 8; r++) if (uint32_t(uint64_t(ys) * num[r] / den[r]) == xs) { ratio = r; break; }
  bw.Write(3, ratio);
  if (!ratio) { if (small) bw.Write(5, xs / 8 - 1); else bw.U32(BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1), xs); }
}
inline BitDepth ReadBitDepth(BitReader& br) {
  BitDepth b; b.float_sample = br.Bool();
  if (!b.float_sample) { b.bits = br.U32(Val(8), Val(10), Val(12), BitsOffset(6, 1)); b.exp_bits = 0; JXLO_CHECK(b.bits <= 31, "bit depth"); }
  else { b.bits = br.U32(Val(32), Val(16), Val(24), BitsOffset(6, 1)); b.exp_bits = br.ReadBits(4) + 1; JXLO_CHECK(b.exp_bits >= 2 && b.exp_bits <= 8 && b.bits <= 32 && b.bits > b.exp_bits + 2, "float bit depth"); }
  return b;
}
inline void WriteBitDepth(BitWriter& bw, const BitDepth& b) {
  bw.Bool(b.float_sample);
  if (!b.float_sample) bw.U32(Val(8), Val(10), Val(12), BitsOffset(6, 1), b.bits);
  else { bw.U32(Val(32), Val(16), Val(24), BitsOffset(6, 1), b.bits); bw.Write(4, b.exp_bits - 1); }
}
inline int32_t ReadCustomXY(BitReader& br) { return UnpackSigned(br.U32(Bits(19), BitsOffset(19, 524288), BitsOffset(20, 1048576), BitsOffset(21, 2097152))); }
inline void WriteCustomXY(BitWriter& bw, int32_t v) { bw.U32(Bits(19), BitsOffset(19, 524288), BitsOffset(20, 1048576), BitsOffset(21, 2097152), PackSigned(v)); }

inline ColorEncoding ReadColorEncoding(BitReader& br) {
  ColorEncoding c; if (br.Bool()) return c;
  c.want_icc = br.Bool(); c.color_space = br.Enum(); JXLO_CHECK(c.color_space <= 3, "color space");
  if (!c.want_icc) {
    if (c.color_space != kCsXYB) { c.white_point = br.Enum(); if (c.white_point == kWpCustom) { c.white_xy[0] = ReadCustomXY(br); c.white_xy[1] = ReadCustomXY(br); } }
    if (c.color_space != kCsXYB && c.color_space != kCsGray) { c.primaries = br.Enum(); if (c.primaries == kPrCustom) for (int i = 0; i < 3; i++) { c.prim_xy[i][0] = ReadCustomXY(br); c.prim_xy[i][1] = ReadCustomXY(br); } }
    if (c.color_space != kCsXYB) { c.have_gamma = br.Bool(); if (c.have_gamma) { c.gamma = br.ReadBits(24); JXLO_CHECK(c.gamma > 0 && c.gamma <= 10000000, "gamma"); } else c.tf = br.Enum(); }
    else { c.have_gamma = true; c.gamma = 3333333; }
    c.intent = br.Enum(); JXLO_CHECK(c.intent <= 3, "rendering intent");
  }
  return c;
}
inline void WriteColorEncoding(BitWriter& bw, const ColorEncoding& c) {
  if (c.IsDefault()) { bw.Bool(true); return; }
  bw.Bool(false); bw.Bool(c.want_icc); bw.Enum(c.color_space);
  if (!c.want_icc) {
    if (c.color_space != kCsXYB) { bw.Enum(c.white_point); if (c.white_point == kWpCustom) { WriteCustomXY(bw, c.white_xy[0]); WriteCustomXY(bw, c.white_xy[1]); } }
    if (c.color_space != kCsXYB && c.color_space != kCsGray) { bw.Enum(c.primaries); if (c.primaries == kPrCustom) for (int i = 0; i < 3; i++) { WriteCustomXY(bw, c.prim_xy[i][0]); WriteCustomXY(bw, c.prim_xy[i][1]); } }
    if (c.color_space != kCsXYB) { bw.Bool(c.have_gamma); if (c.have_gamma) bw.Write(24, c.gamma); else bw.Enum(c.tf); }
    bw.Enum(c.intent);
  }
}

// ---- ICC stream (A.3 "ICC stream" [M/L]): entropy stream + predictor. See jxlo_icc.h.
std::vector<uint8_t> ReadIccStream(BitReader& br);
void WriteIccStream(BitWriter& bw, const std::vector<uint8_t>& icc);

inline ImageMetadata ReadImageHeaders(BitReader& br) {
  ImageMetadata m; ReadSize(br, &m.xsize, &m.ysize);
  if (!br.Bool()) {
    bool extra = br.Bool();
    if (extra) {
      m.orientation = 1 + br.ReadBits(3);
      m.have_intrinsic = br.Bool(); if (m.have_intrinsic) ReadSize(br, &m.intrinsic_x, &m.intrinsic_y);
      m.have_preview = br.Bool();
      if (m.have_preview) {
        bool div8 = br.Bool(); uint32_t y = div8 ? 8 * br.U32(Val(16), Val(32), BitsOffset(5, 1), BitsOffset(9, 33)) : br.U32(BitsOffset(6, 1), BitsOffset(8, 65), BitsOffset(10, 321), BitsOffset(12, 1345));
        uint32_t ratio = br.ReadBits(3), x; static const uint32_t num[8] = {0, 1, 12, 4, 3, 16, 5, 2}, den[8] = {1, 1, 10, 3, 2, 9, 4, 1};
        if (!ratio) x = div8 ? 8 * br.U32(Val(16), Val(32), BitsOffset(5, 1), BitsOffset(9, 33)) : br.U32(BitsOffset(6, 1), BitsOffset(8, 65), BitsOffset(10, 321), BitsOffset(12, 1345));
        else x = uint32_t(uint64_t(y) * num[ratio] / den[ratio]);
        m.preview_x = x; m.preview_y = y;
      }
      m.have_animation = br.Bool();
      if (m.have_animation) { m.anim.tps_num = br.U32(Val(100), Val(1000), BitsOffset(10, 1), BitsOffset(30, 1)); m.anim.tps_den = br.U32(Val(1), Val(1001), BitsOffset(8, 1), BitsOffset(10, 1));
        m.anim.loops = br.U32(Val(0), Bits(3), Bits(16), Bits(32)); m.anim.have_timecodes = br.Bool(); }
    }
    m.bd = ReadBitDepth(br); m.modular_16bit = br.Bool();
    uint32_t nec = br.U32(Val(0), Val(1), BitsOffset(4, 2), BitsOffset(12, 1));
    for (uint32_t i = 0; i < nec; i++) {
      ExtraChannelInfo e;
      if (!br.Bool()) {
        e.type = br.Enum(); e.bd = ReadBitDepth(br); e.dim_shift = br.U32(Val(0), Val(3), Val(4), BitsOffset(3, 1));
        uint32_t nl = br.U32(Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48)); for (uint32_t k = 0; k < nl; k++) e.name.push_back(char(br.ReadBits(8)));
        if (e.type == kEcAlpha) e.alpha_associated = br.Bool();
        if (e.type == kEcSpot) for (int k = 0; k < 4; k++) e.spot[k] = br.F16();
        if (e.type == kEcCFA) e.cfa = br.U32(Val(1), Bits(2), BitsOffset(4, 3), BitsOffset(8, 19));
      }
      m.ec.push_back(e);
    }
    m.xyb_encoded = br.Bool(); m.ce = ReadColorEncoding(br);
    if (extra && !br.Bool()) { m.tm.intensity_target = br.F16(); m.tm.min_nits = br.F16(); m.tm.relative_to_max_display = br.Bool(); m.tm.linear_below = br.F16(); JXLO_CHECK(m.tm.intensity_target > 0, "intensity target"); }
    br.Extensions();
  }
  if (!br.Bool()) {   // CustomTransformData
    if (m.xyb_encoded && !br.Bool()) { for (auto& v : m.opsin.inv) v = br.F16(); for (auto& v : m.opsin.bias) v = br.F16(); for (auto& v : m.opsin.quant_bias) v = br.F16(); }
    m.cw_mask = br.ReadBits(3);
    if (m.cw_mask & 1) { m.up2.resize(15); for (auto& v : m.up2) v = br.F16(); }
    if (m.cw_mask & 2) { m.up4.resize(55); for (auto& v : m.up4) v = br.F16(); }
    if (m.cw_mask & 4) { m.up8.resize(210); for (auto& v : m.up8) v = br.F16(); }
    // (CustomTransformData carries no extensions field)
  }
  if (m.ce.want_icc) m.icc = ReadIccStream(br);
  br.ZeroPadToByte();
  JXLO_CHECK(!br.overrun, "image header truncated");
  return m;
}

inline void WriteImageHeaders(BitWriter& bw, const ImageMetadata& m) {
  WriteSize(bw, m.xsize, m.ysize);
  bool tm_default = m.tm.intensity_target == 255.f && m.tm.min_nits == 0.f && !m.tm.relative_to_max_display && m.tm.linear_below == 0.f;
  bool extra = m.orientation != 1 || !tm_default;
  bool all_default = !extra && !m.bd.float_sample && m.bd.bits == 8 && m.modular_16bit && m.ec.empty() && m.xyb_encoded && m.ce.IsDefault();
  bw.Bool(all_default);
  if (!all_default) {
    bw.Bool(extra);
    if (extra) { bw.Write(3, m.orientation - 1); bw.Bool(false); bw.Bool(false); bw.Bool(false); }
    WriteBitDepth(bw, m.bd); bw.Bool(m.modular_16bit);
    bw.U32(Val(0), Val(1), BitsOffset(4, 2), BitsOffset(12, 1), uint32_t(m.ec.size()));
    for (const ExtraChannelInfo& e : m.ec) {
      bool d_alpha = e.type == kEcAlpha && !e.bd.float_sample && e.bd.bits == 8 && e.dim_shift == 0 && e.name.empty() && !e.alpha_associated;
      bw.Bool(d_alpha); if (d_alpha) continue;
      bw.Enum(e.type); WriteBitDepth(bw, e.bd); bw.U32(Val(0), Val(3), Val(4), BitsOffset(3, 1), e.dim_shift);
      bw.U32(Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48), uint32_t(e.name.size())); for (char ch : e.name) bw.Write(8, uint8_t(ch));
      if (e.type == kEcAlpha) bw.Bool(e.alpha_associated);
      if (e.type == kEcSpot) for (int k = 0; k < 4; k++) bw.F16(e.spot[k]);
      if (e.type == kEcCFA) bw.U32(Val(1), Bits(2), BitsOffset(4, 3), BitsOffset(8, 19), e.cfa);
    }
    bw.Bool(m.xyb_encoded); WriteColorEncoding(bw, m.ce);
    if (extra) { bw.Bool(tm_default); if (!tm_default) { bw.F16(m.tm.intensity_target); bw.F16(m.tm.min_nits); bw.Bool(m.tm.relative_to_max_display); bw.F16(m.tm.linear_below); } }
    bw.U64(0);
  }
  bw.Bool(true);   // default CustomTransformData
  if (m.ce.want_icc) WriteIccStream(bw, m.icc);
  bw.ZeroPadToByte();
}

// ------------------------------------------------------------------ frame header
struct Passes { uint32_t num_passes = 1, num_ds = 0; uint32_t shift[8] = {0}, downsample[8] = {0}, last_pass[8] = {0}; };
struct BlendingInfo { uint32_t mode = 0, alpha_channel = 0; bool clamp = false; uint32_t source = 0; };
struct LoopFilter {
  bool gab = true; float gab_w[6] = {1.1f * 0.104699568f, 1.1f * 0.055680538f, 1.1f * 0.104699568f, 1.1f * 0.055680538f, 1.1f * 0.104699568f, 1.1f * 0.055680538f};
  uint32_t epf_iters = 2; float epf_sharp_lut[8] = {0.f, 1.f / 7, 2.f / 7, 3.f / 7, 4.f / 7, 5.f / 7, 6.f / 7, 1.f};
  float epf_channel_scale[3] = {40.0f, 5.0f, 3.5f}; float epf_quant_mul = 0.46f, epf_pass0_sigma_scale = 0.9f, epf_pass2_sigma_scale = 6.5f, epf_border_sad_mul = 2.0f / 3.0f, epf_sigma_for_modular = 1.0f;
};
enum { kFrameRegular = 0, kFrameLF = 1, kFrameReferenceOnly = 2, kFrameSkipProgressive = 3 };
enum { kFlagNoise = 1, kFlagPatches = 2, kFlagSplines = 16, kFlagUseLfFrame = 32, kFlagSkipAdaptiveLfSmoothing = 128 };
struct FrameHeader {
  uint32_t frame_type = kFrameRegular, encoding = 0; uint64_t flags = 0; bool do_ycbcr = false; uint32_t jpeg_upsampling[3] = {0, 0, 0};
  uint32_t upsampling = 1; std::vector<uint32_t> ec_upsampling; uint32_t group_size_shift = 1, x_qm_scale = 3, b_qm_scale = 2; Passes passes; uint32_t lf_level = 0;
  bool have_crop = false; int32_t x0 = 0, y0 = 0; uint32_t width = 0, height = 0; BlendingInfo blending; std::vector<BlendingInfo> ec_blending;
  uint32_t duration = 0, timecode = 0; bool is_last = true; uint32_t save_as_reference = 0; bool save_before_ct = false; std::string name; LoopFilter lf;
  // derived
  uint32_t xsize = 0, ysize = 0;          // frame size in pixels (after upsampling division)
  uint32_t group_dim = 256; uint32_t xgroups = 0, ygroups = 0, xlfgroups = 0, ylfgroups = 0, num_groups = 0, num_lf_groups = 0;
  uint32_t xblocks = 0, yblocks = 0;      // 8x8 blocks (ceil)
};
inline uint32_t DivCeil(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
inline void DeriveFrameDims(FrameHeader& f, const ImageMetadata& m) {
  uint32_t xs = f.have_crop ? f.width : m.xsize, ys = f.have_crop ? f.height : m.ysize;
  xs = DivCeil(xs, f.upsampling); ys = DivCeil(ys, f.upsampling);
  if (f.frame_type == kFrameLF) { xs = DivCeil(xs, 1u << (3 * f.lf_level)); ys = DivCeil(ys, 1u << (3 * f.lf_level)); }
  f.xsize = xs; f.ysize = ys; f.group_dim = 128u << f.group_size_shift;
  f.xgroups = DivCeil(xs, f.group_dim); f.ygroups = DivCeil(ys, f.group_dim); f.xlfgroups = DivCeil(xs, f.group_dim * 8); f.ylfgroups = DivCeil(ys, f.group_dim * 8);
  f.num_groups = f.xgroups * f.ygroups; f.num_lf_groups = f.xlfgroups * f.ylfgroups; f.xblocks = DivCeil(xs, 8); f.yblocks = DivCeil(ys, 8);
}
inline BlendingInfo ReadBlending(BitReader& br, bool have_ec, bool partial) {
  BlendingInfo b; b.mode = br.U32(Val(0), Val(1), Val(2), BitsOffset(2, 3)); JXLO_CHECK(b.mode <= 4, "blend mode");
  if (have_ec && (b.mode == 2 || b.mode == 3)) b.alpha_channel = br.U32(Val(0), Val(1), Val(2), BitsOffset(3, 3));
  if ((have_ec && (b.mode == 2 || b.mode == 3)) || b.mode == 4) b.clamp = br.Bool();
  if (b.mode != 0 || partial) b.source = br.ReadBits(2);
  return b;
}
inline FrameHeader ReadFrameHeader(BitReader& br, const ImageMetadata& m) {
  FrameHeader f; f.ec_upsampling.assign(m.ec.size(), 1); f.ec_blending.assign(m.ec.size(), BlendingInfo());
  if (!br.Bool()) {
    f.frame_type = br.ReadBits(2); f.encoding = br.ReadBits(1); f.flags = br.U64();
    if (!m.xyb_encoded) f.do_ycbcr = br.Bool();
    if (!(f.flags & kFlagUseLfFrame)) {
      if (f.do_ycbcr) for (auto& u : f.jpeg_upsampling) u = br.ReadBits(2);
      f.upsampling = br.U32(Val(1), Val(2), Val(4), Val(8)); for (auto& u : f.ec_upsampling) u = br.U32(Val(1), Val(2), Val(4), Val(8));
    }
    if (f.encoding == 1) f.group_size_shift = br.ReadBits(2);
    if (f.encoding == 0 && m.xyb_encoded) { f.x_qm_scale = br.ReadBits(3); f.b_qm_scale = br.ReadBits(3); }
    if (f.frame_type != kFrameReferenceOnly) {
      Passes& p = f.passes; p.num_passes = br.U32(Val(1), Val(2), Val(3), BitsOffset(3, 4));
      if (p.num_passes != 1) { p.num_ds = br.U32(Val(0), Val(1), Val(2), BitsOffset(1, 3)); JXLO_CHECK(p.num_ds < p.num_passes, "passes");
        for (uint32_t i = 0; i + 1 < p.num_passes; i++) p.shift[i] = br.ReadBits(2);
        for (uint32_t i = 0; i < p.num_ds; i++) p.downsample[i] = br.U32(Val(1), Val(2), Val(4), Val(8));
        for (uint32_t i = 0; i < p.num_ds; i++) p.last_pass[i] = br.U32(Val(0), Val(1), Val(2), Bits(3)); }
    }
    bool full_frame = true;
    if (f.frame_type == kFrameLF) f.lf_level = 1 + br.ReadBits(2);
    else {
      f.have_crop = br.Bool();
      if (f.have_crop) {
        auto d = [&]() { return br.U32(Bits(8), BitsOffset(11, 256), BitsOffset(14, 2304), BitsOffset(30, 18688)); };
        if (f.frame_type != kFrameReferenceOnly) { f.x0 = UnpackSigned(d()); f.y0 = UnpackSigned(d()); }
        f.width = d(); f.height = d();
        full_frame = f.x0 <= 0 && f.y0 <= 0 && int64_t(f.x0) + f.width >= int64_t(m.xsize) && int64_t(f.y0) + f.height >= int64_t(m.ysize);
      }
    }
    if (f.frame_type == kFrameRegular || f.frame_type == kFrameSkipProgressive) {
      f.blending = ReadBlending(br, !m.ec.empty(), !full_frame); for (auto& b : f.ec_blending) b = ReadBlending(br, !m.ec.empty(), !full_frame);
      if (m.have_animation) { f.duration = br.U32(Val(0), Val(1), Bits(8), Bits(32)); if (m.anim.have_timecodes) f.timecode = br.ReadBits(32); }
      f.is_last = br.Bool();
    } else f.is_last = false;
    if (f.frame_type != kFrameLF && !f.is_last) f.save_as_reference = br.ReadBits(2);
    if (f.frame_type != kFrameLF) {
      bool can_ref = !f.is_last && (f.duration == 0 || f.save_as_reference != 0);
      if (f.frame_type == kFrameReferenceOnly) f.save_before_ct = br.Bool();
      else if (can_ref && f.blending.mode == 0 && full_frame) f.save_before_ct = br.Bool();
    }
    uint32_t nl = br.U32(Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48)); for (uint32_t k = 0; k < nl; k++) f.name.push_back(char(br.ReadBits(8)));
    LoopFilter& l = f.lf;
    if (!br.Bool()) {
      l.gab = br.Bool(); if (l.gab && br.Bool()) for (auto& w : l.gab_w) w = br.F16();
      l.epf_iters = br.ReadBits(2);
      if (l.epf_iters) {
        if (f.encoding == 0 && br.Bool()) for (auto& s : l.epf_sharp_lut) s = br.F16();
        if (br.Bool()) { for (auto& s : l.epf_channel_scale) s = br.F16(); br.ReadBits(32); }
        if (br.Bool()) { if (f.encoding == 0) l.epf_quant_mul = br.F16(); l.epf_pass0_sigma_scale = br.F16(); l.epf_pass2_sigma_scale = br.F16(); l.epf_border_sad_mul = br.F16(); }
        if (f.encoding == 1) l.epf_sigma_for_modular = br.F16();
      }
      br.Extensions();
    }
    br.Extensions();
  }
  JXLO_CHECK(!br.overrun, "frame header truncated");
  DeriveFrameDims(f, m);
  return f;
}
// The oracle encoder only writes single, full-canvas, last frames.
inline void WriteFrameHeader(BitWriter& bw, const FrameHeader& f, const ImageMetadata& m) {
  bw.Bool(false); bw.Write(2, f.frame_type); bw.Write(1, f.encoding); bw.U64(f.flags);
  if (!m.xyb_encoded) bw.Bool(false);
  bw.U32(Val(1), Val(2), Val(4), Val(8), 1); for (size_t i = 0; i < m.ec.size(); i++) bw.U32(Val(1), Val(2), Val(4), Val(8), 1);
  if (f.encoding == 1) bw.Write(2, f.group_size_shift);
  if (f.encoding == 0 && m.xyb_encoded) { bw.Write(3, f.x_qm_scale); bw.Write(3, f.b_qm_scale); }
  bw.U32(Val(1), Val(2), Val(3), BitsOffset(3, 4), f.passes.num_passes);
  if (f.passes.num_passes != 1) { bw.U32(Val(0), Val(1), Val(2), BitsOffset(1, 3), f.passes.num_ds); for (uint32_t i = 0; i + 1 < f.passes.num_passes; i++) bw.Write(2, f.passes.shift[i]);
    for (uint32_t i = 0; i < f.passes.num_ds; i++) bw.U32(Val(1), Val(2), Val(4), Val(8), f.passes.downsample[i]); for (uint32_t i = 0; i < f.passes.num_ds; i++) bw.U32(Val(0), Val(1), Val(2), Bits(3), f.passes.last_pass[i]); }
  // crop, blending, is_last, save_as_reference: mirrors ReadFrameHeader field by field (regular frames of a still image: no duration)
  JXLO_CHECK(f.frame_type == kFrameRegular, "the oracle encoder writes regular frames only");
  bw.Bool(f.have_crop); bool full_frame = true;
  if (f.have_crop) {
    auto d = [&](uint32_t v) { bw.U32(Bits(8), BitsOffset(11, 256), BitsOffset(14, 2304), BitsOffset(30, 18688), v); };
    d(PackSigned(f.x0)); d(PackSigned(f.y0)); d(f.width); d(f.height);
    full_frame = f.x0 <= 0 && f.y0 <= 0 && int64_t(f.x0) + f.width >= int64_t(m.xsize) && int64_t(f.y0) + f.height >= int64_t(m.ysize);
  }
  auto blending = [&](const BlendingInfo& b) {
    bw.U32(Val(0), Val(1), Val(2), BitsOffset(2, 3), b.mode); const bool have_ec = !m.ec.empty();
    if (have_ec && (b.mode == 2 || b.mode == 3)) bw.U32(Val(0), Val(1), Val(2), BitsOffset(3, 3), b.alpha_channel);
    if ((have_ec && (b.mode == 2 || b.mode == 3)) || b.mode == 4) bw.Bool(b.clamp);
    if (b.mode != 0 || !full_frame) bw.Write(2, b.source);
  };
  blending(f.blending); for (size_t i = 0; i < m.ec.size(); i++) blending(i < f.ec_blending.size() ? f.ec_blending[i] : BlendingInfo());
  bw.Bool(f.is_last);
  if (!f.is_last) { bw.Write(2, f.save_as_reference); if (f.blending.mode == 0 && full_frame) bw.Bool(f.save_before_ct); }   // can_ref holds: duration is 0
  bw.U32(Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48), uint32_t(f.name.size())); for (char ch : f.name) bw.Write(8, uint8_t(ch));
  const LoopFilter& l = f.lf; LoopFilter d;
  bool gab_custom = memcmp(l.gab_w, d.gab_w, sizeof(d.gab_w)) != 0;
  bool all_default = l.gab && !gab_custom && l.epf_iters == 2;
  bw.Bool(all_default);
  if (!all_default) {
    bw.Bool(l.gab); if (l.gab) { bw.Bool(gab_custom); if (gab_custom) for (float w : l.gab_w) bw.F16(w); }
    bw.Write(2, l.epf_iters);
    if (l.epf_iters) { if (f.encoding == 0) bw.Bool(false); bw.Bool(false); bw.Bool(false); if (f.encoding == 1) bw.F16(l.epf_sigma_for_modular); }
    bw.U64(0);
  }
  bw.U64(0);
}

// ------------------------------------------------------------------ TOC (A.5)
inline size_t NumTocEntries(const FrameHeader& f) { return (f.num_groups == 1 && f.passes.num_passes == 1) ? 1 : 2 + f.num_lf_groups + size_t(f.num_groups) * f.passes.num_passes; }
inline std::vector<uint32_t> DecodeLehmer(const std::vector<uint32_t>& lehmer, size_t n) {
  std::vector<uint32_t> temp(n), perm(n); for (size_t i = 0; i < n; i++) temp[i] = uint32_t(i);
  for (size_t i = 0; i < n; i++) { JXLO_CHECK(lehmer[i] < temp.size(), "lehmer code"); perm[i] = temp[lehmer[i]]; temp.erase(temp.begin() + lehmer[i]); }
  return perm;
}
inline uint32_t PermCtx(uint32_t v) { return std::min<uint32_t>(7, v == 0 ? 0 : FloorLog2(v) + 1); }
inline std::vector<uint32_t> ReadPermutation(SymbolReader& r, size_t skip, size_t size) {
  uint32_t end = r.Read(PermCtx(uint32_t(size))); JXLO_CHECK(end <= size - skip, "permutation length");
  std::vector<uint32_t> lehmer(size, 0); uint32_t last = 0;
  for (uint32_t i = 0; i < end; i++) { lehmer[skip + i] = last = r.Read(PermCtx(last)); JXLO_CHECK(last < size - skip - i, "lehmer value"); }
  return DecodeLehmer(lehmer, size);
}
// returns byte offsets+sizes of sections in bitstream order, indexed by logical section id
struct Toc { std::vector<size_t> offset, size; size_t total = 0; };
inline Toc ReadToc(BitReader& br, const FrameHeader& f) {
  size_t n = NumTocEntries(f); Toc t; std::vector<uint32_t> perm;
  if (br.Bool()) { Code c = DecodeCode(br, 8); SymbolReader r(&c, &br); perm = ReadPermutation(r, 0, n); JXLO_CHECK(r.CheckFinal(), "TOC permutation final state"); }
  br.ZeroPadToByte();
  std::vector<size_t> sizes(n); for (auto& s : sizes) s = br.U32(Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712));
  br.ZeroPadToByte(); JXLO_CHECK(!br.overrun, "TOC truncated");
  t.offset.assign(n, 0); t.size.assign(n, 0); size_t pos = 0;
  // Sizes are coded in file order; the permutation maps a LOGICAL section to its file slot (libjxl's writer stores section i at
  // slot permutation[i], and its reader takes offsets[permutation[i]] for section i).
  std::vector<size_t> slot_pos(n); for (size_t i = 0; i < n; i++) { slot_pos[i] = pos; pos += sizes[i]; }
  for (size_t i = 0; i < n; i++) { size_t slot = perm.empty() ? i : perm[i]; t.offset[i] = slot_pos[slot]; t.size[i] = sizes[slot]; }
  t.total = pos; return t;
}
inline void WriteToc(BitWriter& bw, const std::vector<size_t>& sizes) {
  bw.Bool(false); bw.ZeroPadToByte();
  for (size_t s : sizes) bw.U32(Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712), uint32_t(s));
  bw.ZeroPadToByte();
}

// ------------------------------------------------------------------ container (A.1)
struct Box { char type[5]; const uint8_t* data; size_t size; };
struct ContainerInfo { bool is_container = false; std::vector<uint8_t> codestream; std::vector<Box> boxes; };
inline int SignatureCheck(const uint8_t* d, size_t n) {   // 0 invalid/not enough, 1 codestream, 2 container
  static const uint8_t sig[12] = {0, 0, 0, 0xC, 'J', 'X', 'L', ' ', 0xD, 0xA, 0x87, 0xA};
  if (n >= 2 && d[0] == 0xFF && d[1] == 0x0A) return 1;
  if (n >= 12 && memcmp(d, sig, 12) == 0) return 2;
  return 0;
}
inline ContainerInfo ParseContainer(const uint8_t* d, size_t n) {
  ContainerInfo c; int s = SignatureCheck(d, n); JXLO_CHECK(s != 0, "invalid signature");
  if (s == 1) { c.codestream.assign(d, d + n); return c; }
  c.is_container = true; size_t pos = 0; bool seen_last = false;
  while (pos + 8 <= n) {
    uint64_t size = (uint64_t(d[pos]) << 24) | (d[pos + 1] << 16) | (d[pos + 2] << 8) | d[pos + 3]; size_t hdr = 8; Box b; memcpy(b.type, d + pos + 4, 4); b.type[4] = 0;
    if (size == 1) { JXLO_CHECK(pos + 16 <= n, "box header truncated"); size = 0; for (int i = 0; i < 8; i++) size = (size << 8) | d[pos + 8 + i]; hdr = 16; }
    else if (size == 0) size = n - pos;
    JXLO_CHECK(size >= hdr && size <= uint64_t(n - pos), "box size");   /* written so that a 64-bit extended size cannot wrap: pos always advances by at least hdr */ b.data = d + pos + hdr; b.size = size - hdr;
    if (!strcmp(b.type, "jxlc")) c.codestream.insert(c.codestream.end(), b.data, b.data + b.size);
    else if (!strcmp(b.type, "jxlp")) { JXLO_CHECK(b.size >= 4 && !seen_last, "jxlp box"); if (b.data[0] & 0x80) seen_last = true; c.codestream.insert(c.codestream.end(), b.data + 4, b.data + b.size); }
    c.boxes.push_back(b); pos += size;
  }
  JXLO_CHECK(!c.codestream.empty(), "container without codestream");
  return c;
}
inline void AppendBox(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t size) {
  uint64_t total = size + 8;
  if (total > 0xffffffffull) { total += 8; uint8_t h[16] = {0, 0, 0, 1, uint8_t(type[0]), uint8_t(type[1]), uint8_t(type[2]), uint8_t(type[3])}; for (int i = 0; i < 8; i++) h[8 + i] = uint8_t(total >> (56 - 8 * i)); out.insert(out.end(), h, h + 16); }
  else { uint8_t h[8] = {uint8_t(total >> 24), uint8_t(total >> 16), uint8_t(total >> 8), uint8_t(total), uint8_t(type[0]), uint8_t(type[1]), uint8_t(type[2]), uint8_t(type[3])}; out.insert(out.end(), h, h + 8); }
  out.insert(out.end(), data, data + size);
}
inline std::vector<uint8_t> ContainerPrologue() {
  static const uint8_t sig[12] = {0, 0, 0, 0xC, 'J', 'X', 'L', ' ', 0xD, 0xA, 0x87, 0xA};
  static const uint8_t ftyp[12] = {'j', 'x', 'l', ' ', 0, 0, 0, 0, 'j', 'x', 'l', ' '};
  std::vector<uint8_t> out(sig, sig + 12); AppendBox(out, "ftyp", ftyp, 12); return out;
}

}  // namespace jxlo
