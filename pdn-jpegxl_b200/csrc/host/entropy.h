// pdn-jpegxl_b200 engine — host-side bitstream front-end (entropy-code headers, host symbol reader, histogram/ANS table builder).
// Part of the product: parses what must be parsed serially on the CPU (container, image and
// frame headers, TOC, entropy-code headers, MA tree) and hands flat tables to the sm_100a
// kernels. Replaces the libjxl work reached from N/Decoder/JxlDecoder.cpp:252,454 and
// N/Encoder/JxlEncoder.cpp:128,367 of the reference. Field codes per ISO/IEC 18181-1 as
// digested in SURVEY.md Appendix A (A.6).
#pragma once
#include "bits.h"
#include <array>
#include <memory>

namespace jxlgpu {

struct HybridCfg {
  uint32_t split_exp = 4, msb = 2, lsb = 0;
  uint32_t split() const { return 1u << split_exp; }
};

// token/extra-bit split of a value (A.6 ReadHybridUint inverse)
inline void HybridEncode(const HybridCfg& c, uint32_t v, uint32_t* tok, uint32_t* nbits, uint32_t* bits) {
  if (v < c.split()) { *tok = v; *nbits = 0; *bits = 0; return; }
  uint32_t n = FloorLog2(v), m = v - (1u << n);
  *tok = c.split() + ((n - c.split_exp) << (c.msb + c.lsb)) + ((m >> (n - c.msb)) << c.lsb) + (m & ((1u << c.lsb) - 1));
  *nbits = n - c.msb - c.lsb;
  *bits = (m >> c.lsb) & ((1u << *nbits) - 1);
}

static const int kAnsLogTab = 12;
static const uint32_t kAnsTab = 1u << kAnsLogTab;
static const uint32_t kAnsSignature = 0x13u << 16;

// Alias table (A.6 "Alias table", normative slot -> (symbol, offset) mapping).
struct AnsTable {
  int log_alpha = 5;
  std::vector<uint16_t> freq;                 // per symbol, sums to 4096
  std::vector<uint16_t> cutoff, right, off1;  // per table entry
  void Init(std::vector<uint16_t> dist, int log_alpha_) {
    log_alpha = log_alpha_;
    size_t table_size = size_t(1) << log_alpha; uint32_t entry = kAnsTab >> log_alpha;
    while (!dist.empty() && dist.back() == 0) dist.pop_back();
    if (dist.empty()) dist.push_back(kAnsTab);
    JXLG_CHECK(dist.size() <= table_size, "ANS alphabet too large");
    freq = dist; freq.resize(table_size, 0);
    cutoff.assign(table_size, 0); right.assign(table_size, 0); off1.assign(table_size, 0);
    for (size_t s = 0; s < dist.size(); s++) if (dist[s] == kAnsTab) {
      for (size_t i = 0; i < table_size; i++) { right[i] = uint16_t(s); cutoff[i] = 0; off1[i] = uint16_t(entry * i); }
      return;
    }
    std::vector<uint32_t> under, over, cut(table_size, 0);
    for (size_t i = 0; i < dist.size(); i++) { cut[i] = dist[i]; if (cut[i] > entry) over.push_back(i); else if (cut[i] < entry) under.push_back(i); }
    for (size_t i = dist.size(); i < table_size; i++) under.push_back(i);
    while (!over.empty()) {
      uint32_t o = over.back(); over.pop_back();
      JXLG_CHECK(!under.empty(), "alias table");
      uint32_t u = under.back(); under.pop_back();
      uint32_t by = entry - cut[u]; cut[o] -= by; right[u] = uint16_t(o); off1[u] = uint16_t(cut[o]);
      if (cut[o] < entry) under.push_back(o); else if (cut[o] > entry) over.push_back(o);
    }
    for (size_t i = 0; i < table_size; i++) {
      if (cut[i] == entry) { right[i] = uint16_t(i); off1[i] = 0; cutoff[i] = 0; }
      else { off1[i] = uint16_t(off1[i] - cut[i]); cutoff[i] = uint16_t(cut[i]); }
    }
  }
  inline void Lookup(uint32_t v, uint32_t* sym, uint32_t* offset, uint32_t* f) const {
    uint32_t log_entry = kAnsLogTab - log_alpha; uint32_t i = v >> log_entry, pos = v & ((1u << log_entry) - 1);
    bool g = pos >= cutoff[i]; *sym = g ? right[i] : i; *offset = (g ? uint32_t(uint16_t(off1[i])) : 0u) + pos;
    *offset &= 0xffff; *f = freq[*sym];
  }
};

// Prefix code (A.6, RFC 7932 §3.4/3.5 canonical codes read LSB-first).
struct PrefixTable {
  int max_len = 0; std::vector<uint8_t> len; std::vector<uint16_t> lut_sym; std::vector<uint8_t> lut_len;
  std::vector<uint16_t> enc_code;   // bit-reversed canonical code per symbol (for the encoder)
  void Build(const std::vector<uint8_t>& lengths) {
    len = lengths; max_len = 0; for (uint8_t l : len) max_len = std::max<int>(max_len, l);
    enc_code.assign(len.size(), 0);
    if (max_len == 0) { lut_sym.assign(1, 0); lut_len.assign(1, 0);
      for (size_t s = 0; s < len.size(); s++) if (s == single) lut_sym[0] = uint16_t(s);
      return; }
    lut_sym.assign(size_t(1) << max_len, 0); lut_len.assign(size_t(1) << max_len, 0);
    uint32_t code = 0;
    for (int l = 1; l <= max_len; l++) {
      for (size_t s = 0; s < len.size(); s++) if (len[s] == l) {
        uint32_t rev = 0; for (int b = 0; b < l; b++) if (code >> b & 1) rev |= 1u << (l - 1 - b);
        enc_code[s] = uint16_t(rev);
        for (uint32_t i = rev; i < (1u << max_len); i += 1u << l) { lut_sym[i] = uint16_t(s); lut_len[i] = uint8_t(l); }
        code++;
      }
      code <<= 1;
    }
  }
  size_t single = 0;   // symbol when the code has a single zero-length entry
};

struct Code {
  size_t num_ctx = 0;                 // contexts visible to the caller (without the LZ77 distance ctx)
  bool lz77 = false; uint32_t lz_min_symbol = 0, lz_min_length = 0; HybridCfg lz_len_cfg;
  std::vector<uint8_t> ctx_map;       // size num_ctx (+1 if lz77)
  bool use_prefix = false; int log_alpha = 5;
  std::vector<HybridCfg> cfg;         // per cluster
  std::vector<AnsTable> ans; std::vector<PrefixTable> prefix;
  size_t num_clusters() const { return cfg.size(); }
};

// ---------------------------------------------------------------- decode side
namespace detail {
inline HybridCfg ReadHybridCfg(BitReader& br, int log_alpha) {
  HybridCfg c; c.split_exp = br.ReadBits(CeilLog2(log_alpha + 1)); c.msb = c.lsb = 0;
  JXLG_CHECK(int(c.split_exp) <= log_alpha, "split_exp");
  if (int(c.split_exp) != log_alpha) {
    c.msb = br.ReadBits(CeilLog2(c.split_exp + 1)); JXLG_CHECK(c.msb <= c.split_exp, "msb_in_token");
    c.lsb = br.ReadBits(CeilLog2(c.split_exp - c.msb + 1)); JXLG_CHECK(c.msb + c.lsb <= c.split_exp, "lsb_in_token");
  }
  return c;
}
static const uint8_t kLogCountLut[128][2] = {
  {3,10},{7,12},{3,7},{4,3},{3,6},{3,8},{3,9},{4,5},{3,10},{4,4},{3,7},{4,1},{3,6},{3,8},{3,9},{4,2},
  {3,10},{5,0},{3,7},{4,3},{3,6},{3,8},{3,9},{4,5},{3,10},{4,4},{3,7},{4,1},{3,6},{3,8},{3,9},{4,2},
  {3,10},{6,11},{3,7},{4,3},{3,6},{3,8},{3,9},{4,5},{3,10},{4,4},{3,7},{4,1},{3,6},{3,8},{3,9},{4,2},
  {3,10},{5,0},{3,7},{4,3},{3,6},{3,8},{3,9},{4,5},{3,10},{4,4},{3,7},{4,1},{3,6},{3,8},{3,9},{4,2},
  {3,10},{7,13},{3,7},{4,3},{3,6},{3,8},{3,9},{4,5},{3,10},{4,4},{3,7},{4,1},{3,6},{3,8},{3,9},{4,2},
  {3,10},{5,0},{3,7},{4,3},{3,6},{3,8},{3,9},{4,5},{3,10},{4,4},{3,7},{4,1},{3,6},{3,8},{3,9},{4,2},
  {3,10},{6,11},{3,7},{4,3},{3,6},{3,8},{3,9},{4,5},{3,10},{4,4},{3,7},{4,1},{3,6},{3,8},{3,9},{4,2},
  {3,10},{5,0},{3,7},{4,3},{3,6},{3,8},{3,9},{4,5},{3,10},{4,4},{3,7},{4,1},{3,6},{3,8},{3,9},{4,2}};
inline int PopCountPrecision(int logcount, int shift) { int r = std::min(logcount, shift - ((kAnsLogTab - logcount) >> 1)); return r < 0 ? 0 : r; }

inline std::vector<uint16_t> ReadAnsHistogram(BitReader& br) {
  std::vector<uint16_t> counts;
  if (br.ReadBits(1)) {  // simple
    int ns = br.ReadBits(1) + 1; uint32_t s[2] = {0, 0};
    for (int i = 0; i < ns; i++) s[i] = br.U8();
    counts.assign(std::max(s[0], s[1]) + 1, 0);
    if (ns == 1) counts[s[0]] = kAnsTab;
    else { JXLG_CHECK(s[0] != s[1], "ANS simple: equal symbols"); uint32_t c0 = br.ReadBits(12); counts[s[0]] = uint16_t(c0); counts[s[1]] = uint16_t(kAnsTab - c0); }
    return counts;
  }
  if (br.ReadBits(1)) {  // flat
    uint32_t n = br.U8() + 1; counts.assign(n, uint16_t(kAnsTab / n));
    for (uint32_t i = 0; i < kAnsTab % n; i++) counts[i]++;
    return counts;
  }
  int log = 0; for (; log < 3; log++) if (br.ReadBits(1) == 0) break;
  int shift = int((br.ReadBits(log) | (1u << log)) - 1); JXLG_CHECK(shift <= kAnsLogTab + 1, "ANS shift");
  uint32_t length = br.U8() + 3;
  std::vector<int> logcounts(length, 0), same(length, 0); int omit_log = -1, omit_pos = -1;
  for (uint32_t i = 0; i < length; i++) {
    uint32_t idx = uint32_t(br.Peek(7)); br.Skip(kLogCountLut[idx][0]); logcounts[i] = kLogCountLut[idx][1];
    if (logcounts[i] == kAnsLogTab + 1) { uint32_t rle = br.U8(); same[i] = int(rle + 5); i += rle + 3; continue; }
    if (logcounts[i] > omit_log) { omit_log = logcounts[i]; omit_pos = int(i); }
  }
  JXLG_CHECK(omit_pos >= 0, "ANS histogram: no omit position");
  JXLG_CHECK(!(size_t(omit_pos) + 1 < length && logcounts[omit_pos + 1] == kAnsLogTab + 1), "ANS histogram: RLE after omit");
  counts.assign(length, 0); int prev = 0, numsame = 0; int total = 0;
  for (uint32_t i = 0; i < length; i++) {
    if (same[i]) { numsame = same[i] - 1; prev = i > 0 ? counts[i - 1] : 0; }
    if (numsame > 0) { counts[i] = uint16_t(prev); numsame--; }
    else {
      int code = logcounts[i];
      if (int(i) == omit_pos || code == 0) continue;
      if (code == 1) counts[i] = 1;
      else { int bc = PopCountPrecision(code - 1, shift); counts[i] = uint16_t((1u << (code - 1)) + (br.ReadBits(bc) << (code - 1 - bc))); }
    }
    total += counts[i];
  }
  JXLG_CHECK(total < int(kAnsTab), "ANS histogram: counts exceed table");
  counts[omit_pos] = uint16_t(kAnsTab - total);
  return counts;
}

static const uint8_t kCodeLengthOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
static const uint8_t kClclVlc[16][2] = {{2,0},{2,4},{2,3},{3,2},{2,0},{2,4},{2,3},{4,1},{2,0},{2,4},{2,3},{3,2},{2,0},{2,4},{2,3},{4,5}};

inline PrefixTable ReadPrefixCode(BitReader& br, uint32_t alphabet) {
  PrefixTable t; std::vector<uint8_t> len(alphabet, 0);
  if (alphabet == 1) { t.single = 0; t.Build(len); return t; }
  uint32_t hskip = br.ReadBits(2);
  if (hskip == 1) {
    uint32_t ns = br.ReadBits(2) + 1; int max_bits = 0; for (uint32_t c = alphabet - 1; c; c >>= 1) max_bits++;
    uint32_t s[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < ns; i++) { s[i] = br.ReadBits(max_bits); JXLG_CHECK(s[i] < alphabet, "prefix simple symbol"); }
    for (uint32_t i = 0; i < ns; i++) for (uint32_t j = i + 1; j < ns; j++) JXLG_CHECK(s[i] != s[j], "prefix simple duplicate");
    if (ns == 1) { t.single = s[0]; }
    else if (ns == 2) { len[s[0]] = 1; len[s[1]] = 1; }
    else if (ns == 3) { len[s[0]] = 1; len[s[1]] = 2; len[s[2]] = 2; }
    else { if (br.ReadBits(1)) { len[s[0]] = 1; len[s[1]] = 2; len[s[2]] = 3; len[s[3]] = 3; } else { for (int i = 0; i < 4; i++) len[s[i]] = 2; } }
    t.Build(len); return t;
  }
  uint8_t clcl[18] = {0}; int space = 32, num_codes = 0;
  for (uint32_t i = hskip; i < 18 && space > 0; i++) {
    uint32_t idx = uint32_t(br.Peek(4)); br.Skip(kClclVlc[idx][0]); uint8_t v = kClclVlc[idx][1];
    clcl[kCodeLengthOrder[i]] = v; if (v) { space -= 32 >> v; num_codes++; }
  }
  JXLG_CHECK(num_codes == 1 || space == 0, "prefix code-length code");
  PrefixTable cl; { std::vector<uint8_t> l(clcl, clcl + 18); if (num_codes == 1) { for (int i = 0; i < 18; i++) if (clcl[i]) { cl.single = i; l[i] = 0; } } cl.Build(l); }
  uint32_t symbol = 0; int prev_len = 8, repeat = 0, repeat_len = 0; int sp = 32768;
  while (symbol < alphabet && sp > 0) {
    uint32_t cs;
    if (cl.max_len == 0) cs = uint32_t(cl.single);
    else { uint32_t idx = uint32_t(br.Peek(cl.max_len)); br.Skip(cl.lut_len[idx]); cs = cl.lut_sym[idx]; JXLG_CHECK(cl.lut_len[idx] > 0, "prefix cl code"); }
    if (cs < 16) { repeat = 0; len[symbol++] = uint8_t(cs); if (cs) { prev_len = int(cs); sp -= 32768 >> cs; } }
    else {
      int extra = int(cs) - 14; int new_len = cs == 16 ? prev_len : 0;
      if (repeat_len != new_len) { repeat = 0; repeat_len = new_len; }
      int old = repeat; if (repeat > 0) { repeat -= 2; repeat <<= extra; }
      repeat += int(br.ReadBits(extra)) + 3; int delta = repeat - old;
      JXLG_CHECK(symbol + delta <= alphabet, "prefix repeat overflow");
      for (int i = 0; i < delta; i++) len[symbol++] = uint8_t(repeat_len);
      if (repeat_len) sp -= delta << (15 - repeat_len);
    }
    JXLG_CHECK(!br.overrun, "prefix code truncated");
  }
  JXLG_CHECK(sp == 0, "prefix code lengths do not fill the code space");
  t.Build(len); return t;
}
}  // namespace detail

struct SymbolReader;
inline Code DecodeCode(BitReader& br, size_t num_ctx, bool allow_lz77 = true);

struct SymbolReader {
  const Code* c; BitReader* br; uint32_t state = 0;
  std::vector<uint32_t> window; uint32_t num_to_copy = 0, copy_pos = 0, num_decoded = 0;
  SymbolReader(const Code* code, BitReader* r) : c(code), br(r) {
    if (!c->use_prefix) state = br->ReadBits(32);
    if (c->lz77) window.assign(1u << 20, 0);
  }
  inline uint32_t ReadSym(uint32_t cluster) {
    if (c->use_prefix) {
      const PrefixTable& t = c->prefix[cluster]; if (t.max_len == 0) return uint32_t(t.single);
      uint32_t idx = uint32_t(br->Peek(t.max_len)); JXLG_CHECK(t.lut_len[idx] > 0, "invalid prefix code"); br->Skip(t.lut_len[idx]); return t.lut_sym[idx];
    }
    uint32_t sym, off, f; c->ans[cluster].Lookup(state & 0xfff, &sym, &off, &f);
    state = f * (state >> 12) + off;
    if (state < (1u << 16)) state = (state << 16) | br->ReadBits(16);
    return sym;
  }
  inline uint32_t Hybrid(const HybridCfg& h, uint32_t t) {
    uint32_t split = h.split(); if (t < split) return t;
    uint32_t n = h.split_exp - (h.msb + h.lsb) + ((t - split) >> (h.msb + h.lsb));
    JXLG_CHECK(n < 32, "hybrid uint too large");
    uint32_t low = t & ((1u << h.lsb) - 1); t >>= h.lsb;
    uint32_t hi = (t & ((1u << h.msb) - 1)) | (1u << h.msb);
    return (((hi << n) | br->ReadBits(n)) << h.lsb) | low;
  }
  static void SpecialDistance(uint32_t i, int* dx, int* dy) {
    static const int8_t k[120][2] = {{0,1},{1,0},{1,1},{-1,1},{0,2},{2,0},{1,2},{-1,2},{2,1},{-2,1},{2,2},{-2,2},{0,3},{3,0},{1,3},{-1,3},{3,1},{-3,1},{2,3},{-2,3},{3,2},{-3,2},{0,4},{4,0},{1,4},{-1,4},{4,1},{-4,1},{3,3},{-3,3},{2,4},{-2,4},{4,2},{-4,2},{0,5},{3,4},{-3,4},{4,3},{-4,3},{5,0},{1,5},{-1,5},{5,1},{-5,1},{2,5},{-2,5},{5,2},{-5,2},{4,4},{-4,4},{3,5},{-3,5},{5,3},{-5,3},{0,6},{6,0},{1,6},{-1,6},{6,1},{-6,1},{2,6},{-2,6},{6,2},{-6,2},{4,5},{-4,5},{5,4},{-5,4},{3,6},{-3,6},{6,3},{-6,3},{0,7},{7,0},{1,7},{-1,7},{5,5},{-5,5},{7,1},{-7,1},{4,6},{-4,6},{6,4},{-6,4},{2,7},{-2,7},{7,2},{-7,2},{3,7},{-3,7},{7,3},{-7,3},{5,6},{-5,6},{6,5},{-6,5},{8,0},{4,7},{-4,7},{7,4},{-7,4},{8,1},{8,2},{6,6},{-6,6},{8,3},{5,7},{-5,7},{7,5},{-7,5},{8,4},{6,7},{-6,7},{7,6},{-7,6},{8,5},{7,7},{-7,7},{8,6},{8,7}};
    *dx = k[i][0]; *dy = k[i][1];
  }
  // ctx: caller context; dist_multiplier: channel width in Modular, 0 elsewhere (A.6 LZ77 [M])
  uint32_t Read(uint32_t ctx, uint32_t dist_multiplier = 0) {
    if (!c->lz77) { uint32_t cl = c->ctx_map[ctx]; return Hybrid(c->cfg[cl], ReadSym(cl)); }
    const uint32_t mask = (1u << 20) - 1;
    if (num_to_copy > 0) { uint32_t r = window[(copy_pos++) & mask]; num_to_copy--; window[(num_decoded++) & mask] = r; return r; }
    uint32_t cl = c->ctx_map[ctx]; uint32_t tok = ReadSym(cl);
    if (tok >= c->lz_min_symbol) {
      num_to_copy = Hybrid(c->lz_len_cfg, tok - c->lz_min_symbol) + c->lz_min_length;
      uint32_t dcl = c->ctx_map[c->num_ctx]; uint32_t dtok = ReadSym(dcl); uint32_t distance = Hybrid(c->cfg[dcl], dtok);
      if (dist_multiplier == 0) distance++;
      else if (distance < 120) { int dx, dy; SpecialDistance(distance, &dx, &dy); int off = dx + int(dist_multiplier) * dy; distance = off < 1 ? 1u : uint32_t(off); }
      else distance -= 119;
      distance = std::min(distance, num_decoded); distance = std::min(distance, 1u << 20);
      copy_pos = num_decoded - distance;
      if (distance == 0) { for (uint32_t i = 0; i < std::min(num_to_copy, 1u << 20); i++) window[i] = 0; }
      JXLG_CHECK(num_to_copy >= c->lz_min_length, "LZ77 length overflow");
      return Read(ctx, dist_multiplier);
    }
    uint32_t r = Hybrid(c->cfg[cl], tok); window[(num_decoded++) & mask] = r; return r;
  }
  bool CheckFinal() const { return c->use_prefix || state == kAnsSignature; }
};

inline std::vector<uint8_t> DecodeContextMap(BitReader& br, size_t n, size_t* num_clusters) {
  std::vector<uint8_t> map(n, 0);
  if (br.ReadBits(1)) { int b = br.ReadBits(2); if (b) for (auto& m : map) m = uint8_t(br.ReadBits(b)); }
  else {
    bool mtf = br.ReadBits(1);
    Code nested = DecodeCode(br, 1, n > 2);
    SymbolReader r(&nested, &br);
    for (auto& m : map) { uint32_t v = r.Read(0); JXLG_CHECK(v < 256, "context map entry"); m = uint8_t(v); }
    JXLG_CHECK(r.CheckFinal(), "context map ANS final state");
    if (mtf) { uint8_t t[256]; for (int i = 0; i < 256; i++) t[i] = uint8_t(i);
      for (auto& m : map) { uint8_t idx = m, v = t[idx]; m = v; for (; idx; idx--) t[idx] = t[idx - 1]; t[0] = v; } }
  }
  uint8_t mx = 0; for (auto m : map) mx = std::max(mx, m); *num_clusters = size_t(mx) + 1;
  return map;
}

inline Code DecodeCode(BitReader& br, size_t num_ctx, bool allow_lz77) {
  Code c; c.num_ctx = num_ctx; size_t nd = num_ctx;
  c.lz77 = br.Bool();
  if (c.lz77) {
    JXLG_CHECK(allow_lz77, "LZ77 not allowed here");
    c.lz_min_symbol = br.U32(Val(224), Val(512), Val(4096), BitsOffset(15, 8));
    c.lz_min_length = br.U32(Val(3), Val(4), BitsOffset(2, 5), BitsOffset(8, 9));
    c.lz_len_cfg = detail::ReadHybridCfg(br, 8); nd++;
  }
  size_t ncl = 1;
  if (nd > 1) c.ctx_map = DecodeContextMap(br, nd, &ncl); else c.ctx_map.assign(1, 0);
  c.use_prefix = br.Bool(); c.log_alpha = c.use_prefix ? 15 : 5 + int(br.ReadBits(2));
  c.cfg.resize(ncl); for (auto& h : c.cfg) h = detail::ReadHybridCfg(br, c.log_alpha);
  if (c.use_prefix) {
    std::vector<uint32_t> asz(ncl);
    for (auto& a : asz) { if (br.ReadBits(1)) { int n = br.ReadBits(4); a = 1 + (1u << n) + br.ReadBits(n); } else a = 1; JXLG_CHECK(a <= (1u << 15), "prefix alphabet size"); }
    c.prefix.resize(ncl); for (size_t i = 0; i < ncl; i++) c.prefix[i] = detail::ReadPrefixCode(br, asz[i]);
  } else {
    c.ans.resize(ncl); for (size_t i = 0; i < ncl; i++) { auto h = detail::ReadAnsHistogram(br); JXLG_CHECK(h.size() <= (size_t(1) << c.log_alpha), "ANS alphabet exceeds log_alpha"); c.ans[i].Init(h, c.log_alpha); }
  }
  JXLG_CHECK(!br.overrun, "entropy code header truncated");
  return c;
}

// ---------------------------------------------------------------- encode side
struct Token { uint32_t ctx; uint32_t value; };

struct EncCode {
  size_t num_ctx = 0; std::vector<uint8_t> ctx_map; bool use_prefix = false; int log_alpha = 5;
  std::vector<HybridCfg> cfg; std::vector<std::vector<uint16_t>> freq;    // normalised (ANS) per cluster
  std::vector<AnsTable> ans; std::vector<std::vector<uint16_t>> rev;       // rev[cluster][cum_offset_of_sym + k] = slot
  std::vector<std::vector<uint32_t>> sym_start;                            // per cluster, start index into rev
  std::vector<PrefixTable> prefix; std::vector<uint32_t> alphabet;         // prefix mode
};

struct EncOptions { int max_clusters = 24; bool use_prefix = false; HybridCfg cfg; bool cluster = true; };

namespace detail {
inline std::vector<uint16_t> Normalize(const std::vector<uint64_t>& h) {
  size_t n = h.size(); std::vector<uint16_t> out(n, 0); uint64_t total = 0; for (auto v : h) total += v;
  if (total == 0) { out.assign(1, uint16_t(kAnsTab)); return out; }
  int64_t rem = kAnsTab; size_t big = 0;
  for (size_t i = 0; i < n; i++) if (h[i]) {
    uint64_t v = h[i] * kAnsTab / total; if (v == 0) v = 1; out[i] = uint16_t(v); rem -= int64_t(v); if (h[i] > h[big] || !h[big]) big = i;
  }
  // distribute remainder on the largest bins
  while (rem != 0) {
    size_t best = big;
    if (rem < 0) { for (size_t i = 0; i < n; i++) if (out[i] > out[best]) best = i; JXLG_CHECK(out[best] > 1, "normalize"); int64_t d = std::min<int64_t>(-rem, out[best] - 1); out[best] = uint16_t(out[best] - d); rem += d; }
    else { out[best] = uint16_t(out[best] + rem); rem = 0; }
  }
  while (out.size() > 1 && out.back() == 0) out.pop_back();
  return out;
}
inline double HistoCost(const std::vector<uint64_t>& h) {
  uint64_t t = 0; for (auto v : h) t += v; if (!t) return 0; double c = 0;
  for (auto v : h) if (v) c -= double(v) * std::log2(double(v) / double(t)); return c;
}
// Huffman code lengths limited to `limit` bits.
inline std::vector<uint8_t> HuffmanLengths(std::vector<uint64_t> h, int limit) {
  size_t n = h.size(); std::vector<uint8_t> len(n, 0);
  for (int attempt = 0; attempt < 32; attempt++) {
    struct Node { uint64_t w; int l, r; }; std::vector<Node> nodes; std::vector<int> live;
    for (size_t i = 0; i < n; i++) if (h[i]) { nodes.push_back({h[i], -1, int(i)}); live.push_back(int(nodes.size()) - 1); }
    if (live.size() == 0) return len; if (live.size() == 1) { len[nodes[live[0]].r] = 1; return len; }
    while (live.size() > 1) {
      std::sort(live.begin(), live.end(), [&](int a, int b) { return nodes[a].w > nodes[b].w; });
      int a = live.back(); live.pop_back(); int b = live.back(); live.pop_back();
      nodes.push_back({nodes[a].w + nodes[b].w, a, b}); live.push_back(int(nodes.size()) - 1);
    }
    std::fill(len.begin(), len.end(), 0); int maxl = 0;
    std::vector<std::pair<int, int>> st; st.push_back({live[0], 0});
    while (!st.empty()) { auto [id, d] = st.back(); st.pop_back(); if (nodes[id].l < 0) { len[nodes[id].r] = uint8_t(d); maxl = std::max(maxl, d); } else { st.push_back({nodes[id].l, d + 1}); st.push_back({nodes[id].r, d + 1}); } }
    if (maxl <= limit) return len;
    uint64_t t = 0; for (auto v : h) t += v; uint64_t floor_ = (t >> std::max(1, limit - attempt)) + 1;
    for (auto& v : h) if (v) v = std::max(v, floor_);
  }
  throw Error("huffman length limit");
}
inline void WriteHybridCfg(BitWriter& bw, const HybridCfg& c, int log_alpha) {
  bw.Write(CeilLog2(log_alpha + 1), c.split_exp);
  if (int(c.split_exp) != log_alpha) { bw.Write(CeilLog2(c.split_exp + 1), c.msb); bw.Write(CeilLog2(c.split_exp - c.msb + 1), c.lsb); }
}
inline void WriteAnsHistogram(BitWriter& bw, const std::vector<uint16_t>& f) {
  std::vector<uint32_t> nz; for (size_t i = 0; i < f.size(); i++) if (f[i]) nz.push_back(uint32_t(i));
  if (nz.size() <= 2) {
    bw.Write(1, 1); bw.Write(1, nz.size() == 2);
    if (nz.empty()) { bw.U8(0); return; }
    for (auto s : nz) bw.U8(s);
    if (nz.size() == 2) bw.Write(12, f[nz[0]]);
    return;
  }
  bw.Write(1, 0); bw.Write(1, 0);
  // shift = 13 (full precision): log=3 -> three 1 bits, then 3 bits = 13+1-8 = 6
  bw.Write(3, 7); bw.Write(3, 6);
  uint32_t length = std::max<uint32_t>(3, uint32_t(f.size())); bw.U8(length - 3);
  static const uint16_t enc[14][2] = {{5,17},{4,11},{4,15},{4,3},{4,9},{4,7},{3,4},{3,2},{3,5},{3,6},{3,0},{6,33},{7,1},{7,65}};
  std::vector<int> lc(length, 0); int omit = -1, omit_log = -1;
  for (uint32_t i = 0; i < length; i++) { uint32_t v = i < f.size() ? f[i] : 0; lc[i] = v ? FloorLog2(v) + 1 : 0; if (lc[i] > omit_log) { omit_log = lc[i]; omit = int(i); } }
  for (uint32_t i = 0; i < length; i++) bw.Write(enc[lc[i]][0], enc[lc[i]][1]);
  for (uint32_t i = 0; i < length; i++) {
    if (int(i) == omit || lc[i] <= 1) continue;
    int bc = PopCountPrecision(lc[i] - 1, 13); uint32_t v = f[i] - (1u << (lc[i] - 1));
    bw.Write(bc, v >> (lc[i] - 1 - bc));
  }
}
inline void WritePrefixCode(BitWriter& bw, const std::vector<uint8_t>& len, uint32_t alphabet) {
  if (alphabet == 1) return;
  std::vector<uint32_t> used; for (uint32_t i = 0; i < alphabet; i++) if (len[i]) used.push_back(i);
  int max_bits = 0; for (uint32_t c = alphabet - 1; c; c >>= 1) max_bits++;
  if (used.size() <= 2) {   // simple code; a single symbol has an implicit zero-length code
    bw.Write(2, 1); bw.Write(2, used.empty() ? 0 : uint32_t(used.size()) - 1);
    if (used.empty()) { bw.Write(max_bits, 0); return; }
    for (auto s : used) bw.Write(max_bits, s);
    return;
  }
  std::vector<uint64_t> clh(18, 0); for (uint32_t i = 0; i < alphabet; i++) clh[len[i]]++;
  std::vector<uint8_t> clcl = HuffmanLengths(clh, 5);
  int nz = 0; for (auto v : clcl) nz += v != 0;
  if (nz == 1) { for (auto& v : clcl) if (v) v = 1; }   // decoder treats single code as zero-length
  bw.Write(2, 0);
  static const uint8_t vlc[6][2] = {{2,0},{4,7},{3,3},{2,2},{2,1},{4,15}};
  int space = 32;
  for (int i = 0; i < 18 && space > 0; i++) { uint8_t v = clcl[kCodeLengthOrder[i]]; bw.Write(vlc[v][0], vlc[v][1]); if (v) space -= 32 >> v; }
  PrefixTable cl; { std::vector<uint8_t> l = clcl; if (nz == 1) for (auto& v : l) v = 0; cl.Build(l); }
  int sp = 32768;
  for (uint32_t i = 0; i < alphabet && sp > 0; i++) { if (nz > 1) bw.Write(clcl[len[i]], cl.enc_code[len[i]]); if (len[i]) sp -= 32768 >> len[i]; }
}
}  // namespace detail

// Builds clustered histograms + tables from per-context token-symbol histograms h[ctx][symbol].
inline EncCode BuildCodeFromHist(std::vector<std::vector<uint64_t>> h, size_t num_ctx, const EncOptions& opt);
// Builds them from all token streams that will share this code.
inline EncCode BuildCode(const std::vector<const std::vector<Token>*>& streams, size_t num_ctx, const EncOptions& opt) {
  std::vector<std::vector<uint64_t>> h(num_ctx);
  for (auto* s : streams) for (const Token& t : *s) {
    uint32_t tok, nb, bits; HybridEncode(opt.cfg, t.value, &tok, &nb, &bits);
    JXLG_CHECK(t.ctx < num_ctx, "token ctx"); auto& hh = h[t.ctx]; if (hh.size() <= tok) hh.resize(tok + 1, 0); hh[tok]++;
  }
  return BuildCodeFromHist(std::move(h), num_ctx, opt);
}
inline EncCode BuildCodeFromHist(std::vector<std::vector<uint64_t>> h, size_t num_ctx, const EncOptions& opt) {
  EncCode e; e.num_ctx = num_ctx; e.use_prefix = opt.use_prefix; h.resize(num_ctx);
  uint32_t max_tok = 0; for (auto& hh : h) { while (!hh.empty() && hh.back() == 0) hh.pop_back(); if (!hh.empty()) max_tok = std::max<uint32_t>(max_tok, uint32_t(hh.size() - 1)); }
  // greedy clustering by entropy cost increase
  std::vector<std::vector<uint64_t>> ch; e.ctx_map.assign(num_ctx, 0);
  std::vector<size_t> order(num_ctx); for (size_t i = 0; i < num_ctx; i++) order[i] = i;
  std::vector<uint64_t> tot(num_ctx, 0); for (size_t i = 0; i < num_ctx; i++) for (auto v : h[i]) tot[i] += v;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tot[a] > tot[b]; });
  std::vector<double> ccost;
  for (size_t ci : order) {
    if (tot[ci] == 0) { if (ch.empty()) { ch.push_back({}); ccost.push_back(0); } e.ctx_map[ci] = 0; continue; }
    double own = detail::HistoCost(h[ci]); double best = 1e300; int bi = -1;
    if (opt.cluster) for (size_t k = 0; k < ch.size(); k++) {
      std::vector<uint64_t> m = ch[k]; if (m.size() < h[ci].size()) m.resize(h[ci].size(), 0); for (size_t j = 0; j < h[ci].size(); j++) m[j] += h[ci][j];
      double inc = detail::HistoCost(m) - ccost[k] - own; if (inc < best) { best = inc; bi = int(k); }
    }
    bool can_new = int(ch.size()) < opt.max_clusters;
    if (bi < 0 || (can_new && best > 0.002 * double(tot[ci]) + 40)) { ch.push_back(h[ci]); ccost.push_back(own); e.ctx_map[ci] = uint8_t(ch.size() - 1); }
    else { auto& m = ch[bi]; if (m.size() < h[ci].size()) m.resize(h[ci].size(), 0); for (size_t j = 0; j < h[ci].size(); j++) m[j] += h[ci][j]; ccost[bi] = detail::HistoCost(m); e.ctx_map[ci] = uint8_t(bi); }
  }
  if (ch.empty()) ch.push_back({});
  size_t ncl = ch.size(); e.cfg.assign(ncl, opt.cfg);
  if (e.use_prefix) {
    e.log_alpha = 15; e.prefix.resize(ncl); e.alphabet.resize(ncl);
    for (size_t k = 0; k < ncl; k++) {
      uint32_t a = std::max<uint32_t>(1, uint32_t(ch[k].size())); e.alphabet[k] = a; std::vector<uint64_t> hh = ch[k]; hh.resize(a, 0);
      std::vector<uint8_t> len = detail::HuffmanLengths(hh, 15); int used = 0; size_t last = 0; for (size_t i = 0; i < a; i++) if (len[i]) { used++; last = i; }
      if (used == 1) { len[last] = 0; e.prefix[k].single = last; }
      e.prefix[k].Build(len); if (used == 1) e.prefix[k].len[last] = 0;
    }
  } else {
    e.log_alpha = 5; while ((1u << e.log_alpha) <= max_tok) e.log_alpha++; JXLG_CHECK(e.log_alpha <= 8, "token alphabet too large for ANS");
    e.freq.resize(ncl); e.ans.resize(ncl); e.rev.resize(ncl); e.sym_start.resize(ncl);
    for (size_t k = 0; k < ncl; k++) {
      e.freq[k] = detail::Normalize(ch[k]); e.ans[k].Init(e.freq[k], e.log_alpha);
      size_t ts = size_t(1) << e.log_alpha; e.sym_start[k].assign(ts + 1, 0);
      for (size_t s = 0; s < ts; s++) e.sym_start[k][s + 1] = e.sym_start[k][s] + e.ans[k].freq[s];
      e.rev[k].assign(kAnsTab, 0);
      for (uint32_t v = 0; v < kAnsTab; v++) { uint32_t sym, off, f; e.ans[k].Lookup(v, &sym, &off, &f); e.rev[k][e.sym_start[k][sym] + off] = uint16_t(v); }
    }
  }
  return e;
}

inline void WriteCode(BitWriter& bw, const EncCode& e);
inline void WriteTokens(BitWriter& bw, const EncCode& e, const std::vector<Token>& toks);
inline void WriteContextMap(BitWriter& bw, const std::vector<uint8_t>& map, size_t ncl) {
  int bits = ncl <= 1 ? 0 : CeilLog2(ncl);
  if (bits <= 3 && map.size() * bits <= 256) { bw.Write(1, 1); bw.Write(2, bits); if (bits) for (auto m : map) bw.Write(bits, m); return; }
  bw.Write(1, 0); bw.Write(1, 0);  // not simple, no MTF
  std::vector<Token> toks; toks.reserve(map.size()); for (auto m : map) toks.push_back(Token{0, m});
  EncOptions o; o.cfg = HybridCfg{4, 2, 0}; o.cluster = false; EncCode nested = BuildCode({&toks}, 1, o);
  WriteCode(bw, nested); WriteTokens(bw, nested, toks);
}

inline void WriteCode(BitWriter& bw, const EncCode& e) {
  bw.Write(1, 0);  // lz77.enabled = 0
  if (e.num_ctx > 1) WriteContextMap(bw, e.ctx_map, e.cfg.size());
  bw.Write(1, e.use_prefix); if (!e.use_prefix) bw.Write(2, e.log_alpha - 5);
  for (auto& h : e.cfg) detail::WriteHybridCfg(bw, h, e.log_alpha);
  if (e.use_prefix) {
    for (auto a : e.alphabet) { if (a == 1) bw.Write(1, 0); else { bw.Write(1, 1); int n = FloorLog2(a - 1); bw.Write(4, n); bw.Write(n, a - 1 - (1u << n)); } }
    for (size_t k = 0; k < e.cfg.size(); k++) { std::vector<uint8_t> len = e.prefix[k].len; if (e.prefix[k].max_len == 0 && e.alphabet[k] > 1) { len.assign(e.alphabet[k], 0); len[e.prefix[k].single] = 1; } detail::WritePrefixCode(bw, len, e.alphabet[k]); }
  } else for (auto& f : e.freq) detail::WriteAnsHistogram(bw, f);
}

// Writes one token stream (ANS state word first; A.6 "ANS symbol read" mirrored).
inline void WriteTokens(BitWriter& bw, const EncCode& e, const std::vector<Token>& toks) {
  size_t n = toks.size();
  if (e.use_prefix) {
    for (const Token& t : toks) { uint32_t cl = e.ctx_map[t.ctx], tok, nb, bits; HybridEncode(e.cfg[cl], t.value, &tok, &nb, &bits);
      const PrefixTable& p = e.prefix[cl]; if (p.max_len) bw.Write(p.len[tok], p.enc_code[tok]); bw.Write(nb, bits); }
    return;
  }
  std::vector<uint32_t> flush(n); std::vector<uint8_t> has(n, 0); uint32_t state = kAnsSignature;
  for (size_t i = n; i-- > 0;) {
    uint32_t cl = e.ctx_map[toks[i].ctx], tok, nb, bits; HybridEncode(e.cfg[cl], toks[i].value, &tok, &nb, &bits);
    uint32_t f = e.ans[cl].freq[tok]; JXLG_CHECK(f > 0, "ANS encode: zero-frequency symbol");
    if ((state >> (32 - kAnsLogTab)) >= f) { flush[i] = state & 0xffff; has[i] = 1; state >>= 16; }
    state = ((state / f) << kAnsLogTab) + e.rev[cl][e.sym_start[cl][tok] + (state % f)];
  }
  bw.Write(32, state);
  for (size_t i = 0; i < n; i++) {
    if (has[i]) bw.Write(16, flush[i]);
    uint32_t cl = e.ctx_map[toks[i].ctx], tok, nb, bits; HybridEncode(e.cfg[cl], toks[i].value, &tok, &nb, &bits); bw.Write(nb, bits);
  }
}

}  // namespace jxlgpu
