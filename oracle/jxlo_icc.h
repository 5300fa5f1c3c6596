// ============================================================================================
// oracle/jxlo_icc.h — TEST INFRASTRUCTURE ONLY (see oracle/README in jxlo_capi.cc). PARITY UNPINNED:
// restated from memory of ISO/IEC 18181-1 Annex E ("ICC profile") / SURVEY.md A.3; no libjxl, no ICC-carrying .jxl file
// exists offline to check it against. Reached in the reference through JxlDecoderGetColorAsICCProfile
// (N/Decoder/JxlDecoder.cpp:606-631,658-681) and JxlEncoderSetICCProfile (N/Encoder/JxlEncoder.cpp:258-262).
//
// The codestream carries the profile as: U64 enc_size, an entropy-coded byte stream with 41 contexts (context from the two
// previous bytes), whose bytes are a *predicted* ICC: varint output size, varint command-stream size, the command stream, then the
// data stream. The decoder replays the commands (header prediction, tag-table shortcuts, raw / shuffled / N-th order predicted runs).
// The writer here uses the plain subset: predicted header, no tag-list shortcuts, one "insert" run.
// ============================================================================================
#pragma once
#include "jxlo_entropy.h"

namespace jxlo {

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
  for (;;) { JXLO_CHECK(*pos < end && shift < 63, "ICC varint"); uint8_t b = d[(*pos)++]; v |= uint64_t(b & 127) << shift; if (!(b & 128)) break; shift += 7; }
  return v;
}
inline void IccPutVarInt(std::vector<uint8_t>& d, uint64_t v) { while (v > 127) { d.push_back(uint8_t(v & 127) | 128); v >>= 7; } d.push_back(uint8_t(v)); }
inline void IccPut32(std::vector<uint8_t>& d, uint64_t v) { JXLO_CHECK(v <= 0xffffffffull, "ICC value exceeds 32 bits"); d.push_back(uint8_t(v >> 24)); d.push_back(uint8_t(v >> 16)); d.push_back(uint8_t(v >> 8)); d.push_back(uint8_t(v)); }
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
  const uint64_t osize = IccVarInt(enc, &pos, size); JXLO_CHECK(osize <= (1ull << 28), "ICC profile too large");
  const uint64_t csize = IccVarInt(enc, &pos, size); size_t cpos = pos; JXLO_CHECK(csize <= size - cpos, "ICC command stream size");
  const size_t cend = cpos + size_t(csize); pos = cend;
  std::vector<uint8_t> header = IccInitialHeader(osize);
  for (size_t i = 0; i <= kIccHeaderSize; i++) {
    if (out.size() == osize) { JXLO_CHECK(cpos == cend && pos == size, "ICC stream has trailing data"); return out; }
    if (i == kIccHeaderSize) break;
    IccPredictHeader(out, header, i); JXLO_CHECK(pos < size, "ICC header truncated"); out.push_back(uint8_t(enc[pos++] + header[i]));
  }
  JXLO_CHECK(cpos < cend, "ICC tag list missing");
  uint64_t numtags = IccVarInt(enc, &cpos, cend);
  if (numtags != 0) {
    numtags--; IccPut32(out, numtags); uint64_t prevstart = kIccHeaderSize + numtags * 12, prevsize = 0;
    for (;;) {
      JXLO_CHECK(out.size() <= osize && cpos <= cend, "ICC tag list overrun"); if (cpos == cend) break;
      const uint8_t command = enc[cpos++], tagcode = command & 63; char tag[5] = {0, 0, 0, 0, 0};
      if (tagcode == 0) break;
      else if (tagcode == 1) { JXLO_CHECK(pos + 4 <= size, "ICC tag keyword"); memcpy(tag, &enc[pos], 4); pos += 4; }
      else if (tagcode == 2) memcpy(tag, "rTRC", 4); else if (tagcode == 3) memcpy(tag, "rXYZ", 4);
      else { JXLO_CHECK(tagcode - 4 < 17, "ICC tag code"); memcpy(tag, kTagStrings[tagcode - 4], 4); }
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
    JXLO_CHECK(out.size() <= osize && cpos <= cend, "ICC content overrun"); if (cpos == cend) break;
    const uint8_t command = enc[cpos++];
    if (command == 1) { const uint64_t num = IccVarInt(enc, &cpos, cend); JXLO_CHECK(num <= size - pos, "ICC insert run"); out.insert(out.end(), enc.begin() + pos, enc.begin() + pos + num); pos += num; }
    else if (command == 2 || command == 3) { const uint64_t num = IccVarInt(enc, &cpos, cend); JXLO_CHECK(num <= size - pos, "ICC shuffle run");
      std::vector<uint8_t> sh(enc.begin() + pos, enc.begin() + pos + num); IccShuffle(sh, command == 2 ? 2 : 4); out.insert(out.end(), sh.begin(), sh.end()); pos += num; }
    else if (command == 4) {
      JXLO_CHECK(cpos + 2 <= cend, "ICC predict command"); const uint8_t flags = enc[cpos++]; const size_t width = (flags & 3) + 1; JXLO_CHECK(width != 3, "ICC predict width"); const int order = (flags & 12) >> 2; JXLO_CHECK(order != 3, "ICC predict order");
      uint64_t stride = width; if (flags & 16) { stride = IccVarInt(enc, &cpos, cend); JXLO_CHECK(stride >= width, "ICC predict stride"); }
      JXLO_CHECK(!out.empty() && ((out.size() - 1) >> 2) >= stride, "ICC predict stride exceeds the decoded prefix");
      const uint64_t num = IccVarInt(enc, &cpos, cend); JXLO_CHECK(num <= size - pos, "ICC predict run");
      std::vector<uint8_t> sh(enc.begin() + pos, enc.begin() + pos + num); if (width > 1) IccShuffle(sh, width);
      const size_t start = out.size(); for (size_t i = 0; i < num; i++) out.push_back(uint8_t(IccLinearPredict(out, start, i, size_t(stride), width, order) + sh[i]));
      pos += num;
    }
    else if (command == 10) { IccPutTag(out, "XYZ "); for (int i = 0; i < 4; i++) out.push_back(0); JXLO_CHECK(pos + 12 <= size, "ICC XYZ command"); out.insert(out.end(), enc.begin() + pos, enc.begin() + pos + 12); pos += 12; }
    else if (command >= 16 && command < 24) { IccPutTag(out, kTypeStrings[command - 16]); for (int i = 0; i < 4; i++) out.push_back(0); }
    else JXLO_CHECK(false, "unknown ICC command");
  }
  JXLO_CHECK(pos == size && out.size() == osize, "ICC stream size mismatch");
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

// (declared in jxlo_headers.h; defined once, in the translation unit that includes this file)
std::vector<uint8_t> ReadIccStream(BitReader& br) {
  const uint64_t enc_size = br.U64(); JXLO_CHECK(enc_size > 0 && enc_size <= (1ull << 28), "ICC stream size");
  Code code = DecodeCode(br, kNumIccContexts); SymbolReader rd(&code, &br);
  std::vector<uint8_t> enc(enc_size);
  for (size_t i = 0; i < enc_size; i++) { uint32_t v = rd.Read(IccContext(i, i > 0 ? enc[i - 1] : 0, i > 1 ? enc[i - 2] : 0)); JXLO_CHECK(v < 256, "ICC byte out of range"); enc[i] = uint8_t(v); }
  JXLO_CHECK(rd.CheckFinal(), "ICC stream ANS final state");
  return UnpredictIcc(enc);
}
void WriteIccStream(BitWriter& bw, const std::vector<uint8_t>& icc) {
  JXLO_CHECK(!icc.empty(), "empty ICC profile");
  const std::vector<uint8_t> enc = PredictIcc(icc); bw.U64(enc.size());
  std::vector<Token> toks(enc.size());
  for (size_t i = 0; i < enc.size(); i++) toks[i] = Token{IccContext(i, i > 0 ? enc[i - 1] : 0, i > 1 ? enc[i - 2] : 0), enc[i]};
  EncOptions opt; opt.max_clusters = 8;
  std::vector<const std::vector<Token>*> streams{&toks}; EncCode ec = BuildCode(streams, kNumIccContexts, opt); WriteCode(bw, ec); WriteTokens(bw, ec, toks);
}

}  // namespace jxlo
