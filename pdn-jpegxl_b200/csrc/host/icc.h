// pdn-jpegxl_b200 engine — ICC profile stream (ISO/IEC 18181-1 Annex E as digested in SURVEY.md A.3), host side.
// Reached in the reference through JxlDecoderGetColorAsICCProfile (N/Decoder/JxlDecoder.cpp:606-631,658-681 -> setIccProfile) and
// JxlEncoderSetICCProfile (N/Encoder/JxlEncoder.cpp:258-262). The codestream carries the profile as: U64 enc_size, an entropy-coded
// byte stream with 41 contexts (context from the two previous bytes), whose bytes are a *predicted* ICC: varint output size, varint
// command-stream size, the command stream, then the data stream. The reader replays every command (header prediction, tag-table
// shortcuts, raw / shuffled / N-th order predicted runs); the writer emits the plain subset (predicted header + one insert run).
#pragma once
#include "entropy.h"
#include "headers.h"
#include <cmath>
#include <string>

namespace jxlgpu {

static const size_t kIccHeaderSize = 128;
static const size_t kNumIccContexts = 41;

inline uint32_t IccContext(size_t i, uint32_t b1, uint32_t b2) {
  if (i <= 128) return 0;
  auto letter = [](uint32_t b) { return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z'); };
  auto digit = [](uint32_t b) { return (b >= '0' && b <= '9') || b == '.' || b == ','; };
  uint32_t p1, p2;
  if (letter(b1)) p1 = 0; else if (digit(b1)) p1 = 1; else if (b1 <= 1) p1 = 2 + b1; else if (b1 > 1 && b1 < 16) p1 = 4; else if (b1 > 240 && b1 < 255) p1 = 5; else if (b1 == 255) p1 = 6; else p1 = 7;
  if (letter(b2)) p2 = 0; else if (digit(b2)) p2 = 1; else if (b2 < 16) p2 = 2; else if (b2 > 240) p2 = 3; else p2 = 4;
  return 1 + p1 + p2 * 8;
}

inline uint64_t IccVarInt(const std::vector<uint8_t>& d, size_t* pos, size_t end) {
  uint64_t v = 0; int shift = 0;
  for (;;) { JXLG_CHECK(*pos < end && shift < 63, "ICC varint"); uint8_t b = d[(*pos)++]; v |= uint64_t(b & 127) << shift; if (!(b & 128)) break; shift += 7; }
  return v;
}
inline void IccPutVarInt(std::vector<uint8_t>& d, uint64_t v) { while (v > 127) { d.push_back(uint8_t(v & 127) | 128); v >>= 7; } d.push_back(uint8_t(v)); }
inline void IccPut32(std::vector<uint8_t>& d, uint64_t v) { JXLG_CHECK(v <= 0xffffffffull, "ICC value exceeds 32 bits"); d.push_back(uint8_t(v >> 24)); d.push_back(uint8_t(v >> 16)); d.push_back(uint8_t(v >> 8)); d.push_back(uint8_t(v)); }
inline void IccPutTag(std::vector<uint8_t>& d, const char* t) { for (int i = 0; i < 4; i++) d.push_back(uint8_t(t[i])); }

inline std::vector<uint8_t> IccInitialHeader(uint64_t osize) {
  std::vector<uint8_t> h(kIccHeaderSize, 0);
  h[0] = uint8_t(osize >> 24); h[1] = uint8_t(osize >> 16); h[2] = uint8_t(osize >> 8); h[3] = uint8_t(osize);
  h[8] = 4; memcpy(&h[12], "mntr", 4); memcpy(&h[16], "RGB ", 4); memcpy(&h[20], "XYZ ", 4); memcpy(&h[36], "acsp", 4);
  const uint8_t d50[12] = {0, 0, 0xF6, 0xD6, 0, 1, 0, 0, 0, 0, 0xD3, 0x2D}; memcpy(&h[68], d50, 12);
  return h;
}
// position-dependent refinements of the header prediction, from the bytes already known
inline void IccPredictHeader(const std::vector<uint8_t>& icc, std::vector<uint8_t>& h, size_t pos) {
  const size_t size = icc.size();
  if (pos == 8 && size >= 8) { h[80] = icc[4]; h[81] = icc[5]; h[82] = icc[6]; h[83] = icc[7]; }
  if (pos == 41 && size >= 41) { if (icc[40] == 'A') { h[41] = 'P'; h[42] = 'P'; h[43] = 'L'; } if (icc[40] == 'M') { h[41] = 'S'; h[42] = 'F'; h[43] = 'T'; } }
  if (pos == 42 && size >= 42) { if (icc[40] == 'S' && icc[41] == 'G') { h[42] = 'I'; h[43] = ' '; } if (icc[40] == 'S' && icc[41] == 'U') { h[42] = 'N'; h[43] = 'W'; } }
}
inline void IccShuffle(std::vector<uint8_t>& d, size_t width) {   // inverse of the encoder's byte-plane split
  const size_t size = d.size(), height = (size + width - 1) / width; std::vector<uint8_t> r(size); size_t s = 0, j = 0;
  for (size_t i = 0; i < size; i++) { r[i] = d[j]; j += height; if (j >= size) j = ++s; }
  d.swap(r);
}
inline uint8_t IccLinearPredict(const std::vector<uint8_t>& d, size_t start, size_t i, size_t stride, size_t width, int order) {
  auto pred = [order](uint32_t p1, uint32_t p2, uint32_t p3) { return order == 0 ? p1 : order == 1 ? 2 * p1 - p2 : 3 * p1 - 3 * p2 + p3; };
  if (width == 1) { size_t pos = start + i; return uint8_t(pred(d[pos - stride], d[pos - stride * 2], d[pos - stride * 3])); }
  const size_t p = start + (i & ~(width - 1)); uint32_t v[3];
  for (int k = 0; k < 3; k++) { uint32_t x = 0; for (size_t b = 0; b < width; b++) x = (x << 8) | d[p - stride * (k + 1) + b]; v[k] = x; }
  const uint32_t r = pred(v[0], v[1], v[2]); const size_t shift = (width - 1 - (i & (width - 1))) * 8; return uint8_t(r >> shift);
}

inline std::vector<uint8_t> UnpredictIcc(const std::vector<uint8_t>& enc) {
  static const char* kTagStrings[17] = {"cprt", "wtpt", "bkpt", "rXYZ", "gXYZ", "bXYZ", "kXYZ", "rTRC", "gTRC", "bTRC", "kTRC", "chad", "desc", "chrm", "dmnd", "dmdd", "lumi"};
  static const char* kTypeStrings[8] = {"XYZ ", "desc", "text", "mluc", "para", "curv", "sf32", "gbd "};
  const size_t size = enc.size(); size_t pos = 0; std::vector<uint8_t> out;
  const uint64_t osize = IccVarInt(enc, &pos, size); JXLG_CHECK(osize <= (1ull << 28), "ICC profile too large");
  const uint64_t csize = IccVarInt(enc, &pos, size); size_t cpos = pos; JXLG_CHECK(csize <= size - cpos, "ICC command stream size");
  const size_t cend = cpos + size_t(csize); pos = cend;
  std::vector<uint8_t> header = IccInitialHeader(osize);
  for (size_t i = 0; i <= kIccHeaderSize; i++) {
    if (out.size() == osize) { JXLG_CHECK(cpos == cend && pos == size, "ICC stream has trailing data"); return out; }
    if (i == kIccHeaderSize) break;
    IccPredictHeader(out, header, i); JXLG_CHECK(pos < size, "ICC header truncated"); out.push_back(uint8_t(enc[pos++] + header[i]));
  }
  JXLG_CHECK(cpos < cend, "ICC tag list missing");
  uint64_t numtags = IccVarInt(enc, &cpos, cend);
  if (numtags != 0) {
    numtags--; IccPut32(out, numtags); uint64_t prevstart = kIccHeaderSize + numtags * 12, prevsize = 0;
    for (;;) {
      JXLG_CHECK(out.size() <= osize && cpos <= cend, "ICC tag list overrun"); if (cpos == cend) break;
      const uint8_t command = enc[cpos++], tagcode = command & 63; char tag[5] = {0, 0, 0, 0, 0};
      if (tagcode == 0) break;
      else if (tagcode == 1) { JXLG_CHECK(pos + 4 <= size, "ICC tag keyword"); memcpy(tag, &enc[pos], 4); pos += 4; }
      else if (tagcode == 2) memcpy(tag, "rTRC", 4); else if (tagcode == 3) memcpy(tag, "rXYZ", 4);
      else { JXLG_CHECK(tagcode - 4 < 17, "ICC tag code"); memcpy(tag, kTagStrings[tagcode - 4], 4); }
      IccPutTag(out, tag);
      uint64_t tagstart, tagsize = prevsize;
      if (!memcmp(tag, "rXYZ", 4) || !memcmp(tag, "gXYZ", 4) || !memcmp(tag, "bXYZ", 4) || !memcmp(tag, "kXYZ", 4) || !memcmp(tag, "wtpt", 4) || !memcmp(tag, "bkpt", 4) || !memcmp(tag, "lumi", 4)) tagsize = 20;
      if (command & 64) tagstart = IccVarInt(enc, &cpos, cend); else tagstart = prevstart + prevsize;
      IccPut32(out, tagstart);
      if (command & 128) tagsize = IccVarInt(enc, &cpos, cend);
      IccPut32(out, tagsize); prevstart = tagstart; prevsize = tagsize;
      if (tagcode == 2) { IccPutTag(out, "gTRC"); IccPut32(out, tagstart); IccPut32(out, tagsize); IccPutTag(out, "bTRC"); IccPut32(out, tagstart); IccPut32(out, tagsize); }
      if (tagcode == 3) { IccPutTag(out, "gXYZ"); IccPut32(out, tagstart + tagsize); IccPut32(out, tagsize); IccPutTag(out, "bXYZ"); IccPut32(out, tagstart + tagsize * 2); IccPut32(out, tagsize); }
    }
  }
  for (;;) {   // main content
    JXLG_CHECK(out.size() <= osize && cpos <= cend, "ICC content overrun"); if (cpos == cend) break;
    const uint8_t command = enc[cpos++];
    if (command == 1) { const uint64_t num = IccVarInt(enc, &cpos, cend); JXLG_CHECK(num <= size - pos, "ICC insert run"); out.insert(out.end(), enc.begin() + pos, enc.begin() + pos + num); pos += num; }
    else if (command == 2 || command == 3) { const uint64_t num = IccVarInt(enc, &cpos, cend); JXLG_CHECK(num <= size - pos, "ICC shuffle run");
      std::vector<uint8_t> sh(enc.begin() + pos, enc.begin() + pos + num); IccShuffle(sh, command == 2 ? 2 : 4); out.insert(out.end(), sh.begin(), sh.end()); pos += num; }
    else if (command == 4) {
      JXLG_CHECK(cpos + 2 <= cend, "ICC predict command"); const uint8_t flags = enc[cpos++]; const size_t width = (flags & 3) + 1; JXLG_CHECK(width != 3, "ICC predict width"); const int order = (flags & 12) >> 2; JXLG_CHECK(order != 3, "ICC predict order");
      uint64_t stride = width; if (flags & 16) { stride = IccVarInt(enc, &cpos, cend); JXLG_CHECK(stride >= width, "ICC predict stride"); }
      JXLG_CHECK(!out.empty() && ((out.size() - 1) >> 2) >= stride, "ICC predict stride exceeds the decoded prefix");
      const uint64_t num = IccVarInt(enc, &cpos, cend); JXLG_CHECK(num <= size - pos, "ICC predict run");
      std::vector<uint8_t> sh(enc.begin() + pos, enc.begin() + pos + num); if (width > 1) IccShuffle(sh, width);
      const size_t start = out.size(); for (size_t i = 0; i < num; i++) out.push_back(uint8_t(IccLinearPredict(out, start, i, size_t(stride), width, order) + sh[i]));
      pos += num;
    }
    else if (command == 10) { IccPutTag(out, "XYZ "); for (int i = 0; i < 4; i++) out.push_back(0); JXLG_CHECK(pos + 12 <= size, "ICC XYZ command"); out.insert(out.end(), enc.begin() + pos, enc.begin() + pos + 12); pos += 12; }
    else if (command >= 16 && command < 24) { IccPutTag(out, kTypeStrings[command - 16]); for (int i = 0; i < 4; i++) out.push_back(0); }
    else JXLG_CHECK(false, "unknown ICC command");
  }
  JXLG_CHECK(pos == size && out.size() == osize, "ICC stream size mismatch");
  return out;
}

// Writer (plain subset): predicted header, empty tag list marker, one insert run for everything after the header.
inline std::vector<uint8_t> PredictIcc(const std::vector<uint8_t>& icc) {
  std::vector<uint8_t> cmds, data, enc; const size_t n = icc.size();
  std::vector<uint8_t> header = IccInitialHeader(n), prefix;
  for (size_t i = 0; i < std::min(n, kIccHeaderSize); i++) { IccPredictHeader(prefix, header, i); data.push_back(uint8_t(icc[i] - header[i])); prefix.push_back(icc[i]); }
  if (n > kIccHeaderSize) { IccPutVarInt(cmds, 0); cmds.push_back(1); IccPutVarInt(cmds, n - kIccHeaderSize); data.insert(data.end(), icc.begin() + kIccHeaderSize, icc.end()); }
  IccPutVarInt(enc, n); IccPutVarInt(enc, cmds.size()); enc.insert(enc.end(), cmds.begin(), cmds.end()); enc.insert(enc.end(), data.begin(), data.end());
  return enc;
}

// (declared in headers.h; defined once, in the translation unit that sets JXLG_ICC_DEFINE_STREAM_FUNCS: decode_engine.cu)
#ifdef JXLG_ICC_DEFINE_STREAM_FUNCS
std::vector<uint8_t> ReadIccStream(BitReader& br) {
  const uint64_t enc_size = br.U64(); JXLG_CHECK(enc_size > 0 && enc_size <= (1ull << 28), "ICC stream size");
  Code code = DecodeCode(br, kNumIccContexts); SymbolReader rd(&code, &br);
  std::vector<uint8_t> enc(enc_size);
  for (size_t i = 0; i < enc_size; i++) { uint32_t v = rd.Read(IccContext(i, i > 0 ? enc[i - 1] : 0, i > 1 ? enc[i - 2] : 0)); JXLG_CHECK(v < 256, "ICC byte out of range"); enc[i] = uint8_t(v); }
  JXLG_CHECK(rd.CheckFinal(), "ICC stream ANS final state");
  return UnpredictIcc(enc);
}
void WriteIccStream(BitWriter& bw, const std::vector<uint8_t>& icc) {
  JXLG_CHECK(!icc.empty(), "empty ICC profile");
  const std::vector<uint8_t> enc = PredictIcc(icc); bw.U64(enc.size());
  std::vector<Token> toks(enc.size());
  for (size_t i = 0; i < enc.size(); i++) toks[i] = Token{IccContext(i, i > 0 ? enc[i - 1] : 0, i > 1 ? enc[i - 2] : 0), enc[i]};
  EncOptions opt; opt.max_clusters = 8;
  std::vector<const std::vector<Token>*> streams{&toks}; EncCode ec = BuildCode(streams, kNumIccContexts, opt); WriteCode(bw, ec); WriteTokens(bw, ec, toks);
}
#endif


// ---------------------------------------------------------------------------------------------------------------------------
// ICC synthesis for enumerated colour encodings. The reference asks libjxl for an ICC blob whenever the encoding is
// not one of its 8 KnownColorProfile values (N/Decoder/JxlDecoder.cpp:600-631 -> setIccProfile); libjxl synthesises one. This is a
// plain ICC v4.4 display profile: matrix/TRC for RGB (rXYZ/gXYZ/bXYZ + one shared parametric curve), kTRC for gray, with the
// chromatic-adaptation tag. PQ / HLG need LUT-based profiles and are not synthesised (no profile is reported, as before).
// ---------------------------------------------------------------------------------------------------------------------------
namespace iccsyn {
inline void Mul3(const double a[9], const double b[9], double o[9]) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { double v = 0; for (int k = 0; k < 3; k++) v += a[3 * r + k] * b[3 * k + c]; o[3 * r + c] = v; } }
inline bool Inv3(const double m[9], double o[9]) {
  double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]); if (std::fabs(det) < 1e-12) return false; double id = 1 / det;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id; o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id; return true;
}
static const double kD50[3] = {0.9642, 1.0, 0.8249};   // PCS illuminant, as stored in ICC headers (0xF6D6, 0x10000, 0xD32D)
static const double kBradford[9] = {0.8951, 0.2664, -0.1614, -0.7502, 1.7135, 0.0367, 0.0389, -0.0685, 1.0296};
// Bradford adaptation matrix from white (wx, wy) to D50
inline bool AdaptToD50(double wx, double wy, double out[9]) {
  if (wy <= 0) return false; const double w[3] = {wx / wy, 1.0, (1 - wx - wy) / wy}; double bi[9]; if (!Inv3(kBradford, bi)) return false;
  double lw[3], ld[3]; for (int i = 0; i < 3; i++) { lw[i] = kBradford[3 * i] * w[0] + kBradford[3 * i + 1] * w[1] + kBradford[3 * i + 2] * w[2]; ld[i] = kBradford[3 * i] * kD50[0] + kBradford[3 * i + 1] * kD50[1] + kBradford[3 * i + 2] * kD50[2]; if (std::fabs(lw[i]) < 1e-12) return false; }
  double d[9] = {ld[0] / lw[0], 0, 0, 0, ld[1] / lw[1], 0, 0, 0, ld[2] / lw[2]}, t[9]; Mul3(d, kBradford, t); Mul3(bi, t, out); return true;
}
inline void WhiteXy(const ColorEncoding& c, double* x, double* y) {
  switch (c.white_point) { case kWpCustom: *x = c.white_xy[0] * 1e-6; *y = c.white_xy[1] * 1e-6; break; case kWpE: *x = *y = 1.0 / 3; break; case kWpDCI: *x = 0.314; *y = 0.351; break; default: *x = 0.3127; *y = 0.3290; }
}
inline void PrimariesXy(const ColorEncoding& c, double p[3][2]) {
  static const double srgb[3][2] = {{0.639998686, 0.330010138}, {0.300003784, 0.600003357}, {0.150002046, 0.059997204}}, bt2100[3][2] = {{0.708, 0.292}, {0.170, 0.797}, {0.131, 0.046}}, p3[3][2] = {{0.680, 0.320}, {0.265, 0.690}, {0.150, 0.060}};
  const double (*src)[2] = c.primaries == kPr2100 ? bt2100 : c.primaries == kPrP3 ? p3 : srgb;
  for (int i = 0; i < 3; i++) for (int k = 0; k < 2; k++) p[i][k] = c.primaries == kPrCustom ? c.prim_xy[i][k] * 1e-6 : src[i][k];
}
// RGB(linear) -> XYZ for the given primaries and white (columns scaled so that RGB = 1,1,1 maps to the white point)
inline bool RgbToXyz(const double p[3][2], double wx, double wy, double m[9]) {
  double P[9]; for (int i = 0; i < 3; i++) { if (p[i][1] == 0) return false; P[i] = p[i][0] / p[i][1]; P[3 + i] = 1.0; P[6 + i] = (1 - p[i][0] - p[i][1]) / p[i][1]; }
  double Pi[9]; if (!Inv3(P, Pi) || wy <= 0) return false; const double W[3] = {wx / wy, 1.0, (1 - wx - wy) / wy}; double S[3];
  for (int i = 0; i < 3; i++) S[i] = Pi[3 * i] * W[0] + Pi[3 * i + 1] * W[1] + Pi[3 * i + 2] * W[2];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) m[3 * r + c] = P[3 * r + c] * S[c]; return true;
}
inline void Put16(std::vector<uint8_t>& d, uint32_t v) { d.push_back(uint8_t(v >> 8)); d.push_back(uint8_t(v)); }
inline void Put32(std::vector<uint8_t>& d, uint32_t v) { d.push_back(uint8_t(v >> 24)); d.push_back(uint8_t(v >> 16)); d.push_back(uint8_t(v >> 8)); d.push_back(uint8_t(v)); }
inline void PutS15(std::vector<uint8_t>& d, double v) { Put32(d, uint32_t(int32_t(std::lround(v * 65536.0)))); }
inline void PutSig(std::vector<uint8_t>& d, const char* s) { for (int i = 0; i < 4; i++) d.push_back(uint8_t(s[i])); }
inline std::vector<uint8_t> Mluc(const std::string& text) { std::vector<uint8_t> t; PutSig(t, "mluc"); Put32(t, 0); Put32(t, 1); Put32(t, 12); PutSig(t, "enUS"); Put32(t, uint32_t(text.size() * 2)); Put32(t, 28); for (char ch : text) Put16(t, uint8_t(ch)); return t; }
inline std::vector<uint8_t> XyzTag(const double v[3]) { std::vector<uint8_t> t; PutSig(t, "XYZ "); Put32(t, 0); for (int i = 0; i < 3; i++) PutS15(t, v[i]); return t; }
inline std::vector<uint8_t> ParaTag(int type, const double* params, int n) { std::vector<uint8_t> t; PutSig(t, "para"); Put32(t, 0); Put16(t, uint32_t(type)); Put16(t, 0); for (int i = 0; i < n; i++) PutS15(t, params[i]); return t; }
}  // namespace iccsyn

// Returns an empty vector when the encoding cannot be expressed as a matrix/TRC (or gray TRC) profile.
inline std::vector<uint8_t> SynthesizeIcc(const ColorEncoding& c) {
  using namespace iccsyn; std::vector<uint8_t> none;
  if (c.want_icc || (c.color_space != kCsRGB && c.color_space != kCsGray)) return none;
  std::vector<uint8_t> trc;
  if (c.have_gamma) { if (c.gamma == 0) return none; const double g = 1.0 / (c.gamma * 1e-7); trc = ParaTag(0, &g, 1); }
  else if (c.tf == kTfSRGB) { const double p[5] = {2.4, 1.0 / 1.055, 0.055 / 1.055, 1.0 / 12.92, 0.04045}; trc = ParaTag(3, p, 5); }
  else if (c.tf == kTf709) { const double p[5] = {1.0 / 0.45, 1.0 / 1.099, 0.099 / 1.099, 1.0 / 4.5, 0.081}; trc = ParaTag(3, p, 5); }
  else if (c.tf == kTfLinear) { const double g = 1.0; trc = ParaTag(0, &g, 1); }
  else if (c.tf == kTfDCI) { const double g = 2.6; trc = ParaTag(0, &g, 1); }
  else return none;   // PQ, HLG, unknown
  double wx, wy; WhiteXy(c, &wx, &wy); double chad[9]; if (!AdaptToD50(wx, wy, chad)) return none;
  const bool gray = c.color_space == kCsGray;
  std::vector<std::pair<std::string, std::vector<uint8_t>>> tags;
  std::string name = std::string(gray ? "Gray" : "RGB") + (c.have_gamma ? "_gamma" : c.tf == kTfSRGB ? "_sRGB-TRC" : c.tf == kTf709 ? "_709-TRC" : c.tf == kTfLinear ? "_linear" : "_DCI") + (gray ? "" : c.primaries == kPrSRGB ? "_sRGB" : c.primaries == kPrP3 ? "_P3" : c.primaries == kPr2100 ? "_2100" : "_custom");
  tags.push_back({"desc", Mluc(name)}); tags.push_back({"cprt", Mluc("CC0")}); tags.push_back({"wtpt", XyzTag(kD50)});
  { std::vector<uint8_t> t; PutSig(t, "sf32"); Put32(t, 0); for (int i = 0; i < 9; i++) PutS15(t, chad[i]); tags.push_back({"chad", t}); }
  if (gray) tags.push_back({"kTRC", trc});
  else {
    double p[3][2], m[9], md50[9]; PrimariesXy(c, p); if (!RgbToXyz(p, wx, wy, m)) return none; Mul3(chad, m, md50);
    const char* names[3] = {"rXYZ", "gXYZ", "bXYZ"}; for (int k = 0; k < 3; k++) { const double col[3] = {md50[k], md50[3 + k], md50[6 + k]}; tags.push_back({names[k], XyzTag(col)}); }
    tags.push_back({"rTRC", trc}); tags.push_back({"gTRC", trc}); tags.push_back({"bTRC", trc});
  }
  // layout: header, tag table, tag data (4-byte aligned; the three TRC tags share one element)
  std::vector<uint8_t> body; std::vector<uint32_t> off(tags.size()), len(tags.size()); const uint32_t base = uint32_t(128 + 4 + 12 * tags.size());
  for (size_t i = 0; i < tags.size(); i++) {
    if (i > 0 && tags[i].second == tags[i - 1].second && tags[i].first.substr(1) == "TRC") { off[i] = off[i - 1]; len[i] = len[i - 1]; continue; }
    while (body.size() & 3) body.push_back(0); off[i] = base + uint32_t(body.size()); len[i] = uint32_t(tags[i].second.size()); body.insert(body.end(), tags[i].second.begin(), tags[i].second.end());
  }
  while (body.size() & 3) body.push_back(0);
  std::vector<uint8_t> icc; const uint32_t total = base + uint32_t(body.size());
  Put32(icc, total); PutSig(icc, "jxl "); Put32(icc, 0x04400000u); PutSig(icc, "mntr"); PutSig(icc, gray ? "GRAY" : "RGB "); PutSig(icc, "XYZ ");
  Put16(icc, 2019); Put16(icc, 12); Put16(icc, 1); Put16(icc, 0); Put16(icc, 0); Put16(icc, 0);   // fixed creation date: the profile is a pure function of the encoding
  PutSig(icc, "acsp"); PutSig(icc, "APPL"); Put32(icc, 0); Put32(icc, 0); Put32(icc, 0); Put32(icc, 0); Put32(icc, 0); Put32(icc, c.intent & 3);
  for (int i = 0; i < 3; i++) PutS15(icc, kD50[i]); PutSig(icc, "jxl "); while (icc.size() < 128) icc.push_back(0);
  Put32(icc, uint32_t(tags.size())); for (size_t i = 0; i < tags.size(); i++) { PutSig(icc, tags[i].first.c_str()); Put32(icc, off[i]); Put32(icc, len[i]); }
  icc.insert(icc.end(), body.begin(), body.end());
  return icc;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Matrix/TRC reading for the encoder (SaveImage with metadata->iccProfile, N/Encoder/JxlEncoder.cpp:258-262). libjxl hands the
// profile to a CMS to reach XYB; the engine has no CMS on any path, so it reads what a matrix/TRC RGB profile states directly:
// the D50-adapted colorants and the three tone curves ('curv' tables / gamma, 'para' types 0-4). Anything else (LUT-based, CMYK,
// Lab PCS) is refused with a clear error instead of being encoded with wrong colours.
// ---------------------------------------------------------------------------------------------------------------------------
struct IccMatrixTrc { double to_linear_srgb[9]; float lut[3][256]; };
inline bool ParseMatrixTrcIcc(const uint8_t* icc, size_t n, IccMatrixTrc* out, std::string* why) {
  auto fail = [&](const char* m) { if (why) *why = m; return false; };
  auto u32 = [&](size_t o) { return (uint32_t(icc[o]) << 24) | (uint32_t(icc[o + 1]) << 16) | (uint32_t(icc[o + 2]) << 8) | icc[o + 3]; };
  auto s15 = [&](size_t o) { return double(int32_t(u32(o))) / 65536.0; };
  if (n < 132) return fail("ICC profile too short");
  if (memcmp(icc + 16, "RGB ", 4) != 0) return fail("ICC profile is not an RGB profile");
  if (memcmp(icc + 20, "XYZ ", 4) != 0) return fail("ICC profile connection space is not XYZ (LUT-based profile)");
  const uint32_t count = u32(128); if (count > 1000 || size_t(132) + size_t(count) * 12 > n) return fail("ICC tag table truncated");
  auto find = [&](const char* sig, size_t* off, size_t* len) { for (uint32_t i = 0; i < count; i++) { size_t e = 132 + size_t(i) * 12; if (!memcmp(icc + e, sig, 4)) { *off = u32(e + 4); *len = u32(e + 8); return *off <= n && *len <= n - *off; } } return false; };
  double m[9]; const char* xyz[3] = {"rXYZ", "gXYZ", "bXYZ"}; const char* trc[3] = {"rTRC", "gTRC", "bTRC"};
  for (int k = 0; k < 3; k++) { size_t o, l; if (!find(xyz[k], &o, &l) || l < 20 || memcmp(icc + o, "XYZ ", 4) != 0) return fail("ICC profile has no matrix colorants (only matrix/TRC RGB profiles can be encoded without a CMS)"); for (int r = 0; r < 3; r++) m[3 * r + k] = s15(o + 8 + 4 * r); }
  for (int k = 0; k < 3; k++) {
    size_t o, l; if (!find(trc[k], &o, &l) || l < 12) return fail("ICC profile has no tone curves");
    if (!memcmp(icc + o, "curv", 4)) {
      const uint32_t cnt = u32(o + 8); if (size_t(12) + size_t(cnt) * 2 > l) return fail("ICC curve truncated");
      for (int v = 0; v < 256; v++) { const double x = v / 255.0; double y;
        if (cnt == 0) y = x; else if (cnt == 1) y = std::pow(x, ((icc[o + 12] << 8) | icc[o + 13]) / 256.0);
        else { const double pos = x * (cnt - 1); const uint32_t i0 = uint32_t(pos), i1 = std::min(i0 + 1, cnt - 1); const double f = pos - i0; auto at = [&](uint32_t i) { return ((icc[o + 12 + 2 * i] << 8) | icc[o + 13 + 2 * i]) / 65535.0; }; y = at(i0) * (1 - f) + at(i1) * f; }
        out->lut[k][v] = float(y); }
    } else if (!memcmp(icc + o, "para", 4)) {
      const uint32_t type = (icc[o + 8] << 8) | icc[o + 9]; static const int np[5] = {1, 3, 4, 5, 7}; if (type > 4 || size_t(12) + size_t(np[type]) * 4 > l) return fail("ICC parametric curve type");
      double p[7] = {0, 0, 0, 0, 0, 0, 0}; for (int i = 0; i < np[type]; i++) p[i] = s15(o + 12 + 4 * i); const double g = p[0], a = p[1], b = p[2], c = p[3], d = p[4], e = p[5], f = p[6];
      for (int v = 0; v < 256; v++) { const double x = v / 255.0; double y;
        switch (type) { case 0: y = std::pow(x, g); break; case 1: y = (a != 0 && x >= -b / a) ? std::pow(a * x + b, g) : 0; break; case 2: y = (a != 0 && x >= -b / a) ? std::pow(a * x + b, g) + c : c; break;
          case 3: y = x >= d ? std::pow(a * x + b, g) : c * x; break; default: y = x >= d ? std::pow(a * x + b, g) + e : c * x + f; }
        out->lut[k][v] = float(std::min(1.0, std::max(0.0, y))); }
    } else return fail("ICC tone curve type");
  }
  // XYZ(D50) -> linear sRGB: inverse of the D50-adapted sRGB colorants
  double sp[3][2] = {{0.639998686, 0.330010138}, {0.300003784, 0.600003357}, {0.150002046, 0.059997204}}, s65[9], chad[9], s50[9], s50i[9];
  if (!iccsyn::RgbToXyz(sp, 0.3127, 0.3290, s65) || !iccsyn::AdaptToD50(0.3127, 0.3290, chad)) return fail("internal colour matrix"); iccsyn::Mul3(chad, s65, s50); if (!iccsyn::Inv3(s50, s50i)) return fail("internal colour matrix");
  iccsyn::Mul3(s50i, m, out->to_linear_srgb);
  return true;
}

}  // namespace jxlgpu
