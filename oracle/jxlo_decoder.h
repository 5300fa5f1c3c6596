// ORACLE — test infrastructure only (see jxlo_bits.h header). PARITY UNPINNED.
// Frame decode: LfGlobal / LfGroup / HfGlobal / PassGroup sections, dequantisation,
// chroma-from-luma, inverse transforms, adaptive LF smoothing, gaborish, EPF, XYB->RGB,
// transfer functions and sample conversion. Restates SURVEY.md Appendix A.5, A.8-A.10 —
// the work libjxl does inside JxlDecoderProcessInput (N/Decoder/JxlDecoder.cpp:252), with the
// output buffer contract of ReadFrameData (N/Decoder/JxlDecoder.cpp:289-323: interleaved,
// tightly packed, native endian, first frame only, straight alpha :233).
#pragma once
#include "jxlo_headers.h"
#include "jxlo_modular.h"
#include "jxlo_vardct.h"
#include <thread>
#include <atomic>
#include <functional>

namespace jxlo {

inline void ParallelFor(size_t n, int threads, const std::function<void(size_t)>& fn) {
  if (threads <= 1 || n <= 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
  std::atomic<size_t> next(0); std::vector<std::thread> pool; std::string err; std::mutex mu;
  int nt = int(std::min<size_t>(size_t(threads), n));
  for (int t = 0; t < nt; t++) pool.emplace_back([&]() { for (;;) { size_t i = next.fetch_add(1); if (i >= n) break; try { fn(i); } catch (const std::exception& e) { std::lock_guard<std::mutex> lk(mu); if (err.empty()) err = e.what(); } } });
  for (auto& th : pool) th.join();
  if (!err.empty()) throw Error(err);
}

struct Plane { int w = 0, h = 0; std::vector<float> d; Plane() {} Plane(int w_, int h_) : w(w_), h(h_), d(size_t(w_) * h_, 0.f) {} float* row(int y) { return d.data() + size_t(y) * w; } const float* row(int y) const { return d.data() + size_t(y) * w; } };

struct Quantizer { uint32_t global_scale = 1, quant_lf = 16; float InvGlobalScale() const { return 65536.0f / float(global_scale); } };
struct ColorCorrelation { uint32_t color_factor = 84; float base_x = 0.0f, base_b = 1.0f; int32_t x_factor_lf = 0, b_factor_lf = 0;   // *_lf already minus 128
  float YtoX(int f) const { return base_x + float(f) / float(color_factor); } float YtoB(int f) const { return base_b + float(f) / float(color_factor); } };

struct DecodeOptions { int threads = 1; bool keep_stages = false; };

struct FrameState {
  ImageMetadata meta; FrameHeader fh; Toc toc; const uint8_t* frame_data = nullptr; size_t frame_size = 0;   // frame_data points at the first section
  // LfGlobal
  float lf_dequant[3] = {1.0f / 4096, 1.0f / 512, 1.0f / 256}; Quantizer q; BlockCtxMap bctx; ColorCorrelation cfl;
  bool has_tree = false; Tree tree; Code tree_code; ModularImage gimg; GroupHeader gheader; size_t global_decoded = 0;   // channels fully decoded in the global section
  // VarDCT frame-wide images
  int xb = 0, yb = 0, xpad = 0, ypad = 0;
  std::vector<int32_t> lfq[3]; Plane lf[3]; std::vector<uint8_t> strategy; std::vector<uint8_t> is_first; std::vector<int32_t> hf_mul; std::vector<uint8_t> sharp; std::vector<uint8_t> lf_idx;
  int xt = 0, yt = 0; std::vector<int8_t> ytox, ytob;
  // HfGlobal
  std::vector<std::vector<float>> dequant;   // per quant table: 3 * rows*cols
  uint32_t num_hf_presets = 1; std::vector<std::vector<std::vector<uint32_t>>> orders;   // [pass][order*3+c] -> permutation
  std::vector<Code> ac_codes; std::vector<std::vector<uint32_t>> natural;
  Plane xyb[3];
  // debug stages
  std::vector<int32_t> dbg_coeffs;   // [3][ypad*xpad] quantised coefficients laid out per block in storage order at the block's pixel rect (row-major within block rect)
  Plane dbg_idct[3], dbg_gab[3], dbg_epf[3];
};

// A reader positioned at a logical section; with a single TOC entry all sections share one reader.
struct SectionReaders {
  FrameState* fs; BitReader shared; bool single;
  SectionReaders(FrameState* f) : fs(f), shared(f->frame_data, f->toc.size.empty() ? 0 : f->toc.size[0]), single(f->toc.size.size() == 1) {}
  BitReader Get(size_t logical) { if (single) return shared; JXLO_CHECK(fs->toc.offset[logical] + fs->toc.size[logical] <= fs->frame_size, "section beyond end of input (truncated file)"); return BitReader(fs->frame_data + fs->toc.offset[logical], fs->toc.size[logical]); }
  void Done(BitReader& br) { JXLO_CHECK(!br.overrun, "section truncated"); if (single) shared = br; }
};

inline uint32_t StreamIdLfCoeff(const FrameHeader& f, uint32_t g) { (void)f; return 1 + g; }
inline uint32_t StreamIdModularLf(const FrameHeader& f, uint32_t g) { return 1 + f.num_lf_groups + g; }
inline uint32_t StreamIdHfMeta(const FrameHeader& f, uint32_t g) { return 1 + 2 * f.num_lf_groups + g; }
inline uint32_t StreamIdQuantTable(const FrameHeader& f, uint32_t t) { return 1 + 3 * f.num_lf_groups + t; }
inline uint32_t StreamIdModularGroup(const FrameHeader& f, uint32_t pass, uint32_t g) { return 1 + 3 * f.num_lf_groups + 17 + pass * f.num_groups + g; }

// ------------------------------------------------------------------ LfGlobal
inline void DecodeLfGlobal(FrameState& fs, BitReader& br) {
  const FrameHeader& fh = fs.fh; const ImageMetadata& m = fs.meta;
  JXLO_CHECK(!(fh.flags & (kFlagPatches | kFlagSplines | kFlagNoise)), "patches/splines/noise are not supported");
  // LfChannelDequantization comes first for EVERY frame encoding (Modular frames carry it too, unused)
  if (!br.Bool()) for (int c = 0; c < 3; c++) { fs.lf_dequant[c] = br.F16() * (1.0f / 128.0f); JXLO_CHECK(fs.lf_dequant[c] >= 1e-8f, "lf dequant"); }
  if (fh.encoding == 0) {
    fs.q.global_scale = br.U32(BitsOffset(11, 1), BitsOffset(11, 2049), BitsOffset(12, 4097), BitsOffset(16, 8193));
    fs.q.quant_lf = br.U32(Val(16), BitsOffset(5, 1), BitsOffset(8, 1), BitsOffset(16, 1));
    if (!br.Bool()) {
      BlockCtxMap& b = fs.bctx; b.num_lf_ctxs = 1;
      for (int j = 0; j < 3; j++) { uint32_t n = br.ReadBits(4); b.lf_thr[j].resize(n); for (auto& t : b.lf_thr[j]) t = UnpackSigned(br.U32(Bits(4), BitsOffset(8, 16), BitsOffset(16, 272), BitsOffset(32, 65808))); b.num_lf_ctxs *= n + 1; }
      uint32_t nq = br.ReadBits(4); b.qf_thr.resize(nq); for (auto& t : b.qf_thr) t = br.U32(Bits(2), BitsOffset(3, 4), BitsOffset(5, 12), BitsOffset(8, 44)) + 1;
      JXLO_CHECK(b.num_lf_ctxs * (nq + 1) <= 64, "block context map too large");
      size_t ncl = 0; b.map = DecodeContextMap(br, size_t(3) * kNumOrders * b.num_lf_ctxs * (nq + 1), &ncl); JXLO_CHECK(ncl <= 16, "too many block contexts"); b.num_ctxs = uint32_t(ncl);
    }
    if (!br.Bool()) { ColorCorrelation& c = fs.cfl; c.color_factor = br.U32(Val(84), Val(256), BitsOffset(8, 2), BitsOffset(16, 258)); c.base_x = br.F16(); c.base_b = br.F16();
      c.x_factor_lf = int32_t(br.ReadBits(8)) - 128; c.b_factor_lf = int32_t(br.ReadBits(8)) - 128; }
  }
  fs.has_tree = br.Bool();
  if (fs.has_tree) {
    size_t nch = (fh.encoding == 1 ? 3 : 0) + m.ec.size();
    size_t limit = std::min<size_t>(size_t(1) << 22, 1024 + size_t(fh.xsize) * fh.ysize * std::max<size_t>(nch, 1) / 16);
    limit = std::max<size_t>(limit, 1 << 16);   // VarDCT frames: LF/HF-metadata streams also use this tree
    fs.tree = DecodeTree(br, limit); fs.tree_code = DecodeCode(br, NumLeaves(fs.tree));
  }
  // global modular image: colour channels (Modular frames) + extra channels
  ModularImage& g = fs.gimg; g.ch.clear(); g.nb_meta = 0; g.bitdepth = int(m.bd.bits);
  if (fh.encoding == 1) {
    JXLO_CHECK(!fh.do_ycbcr, "YCbCr modular frames not supported");
    int nc = (m.ce.color_space == kCsGray && !m.xyb_encoded) ? 1 : 3; for (int c = 0; c < nc; c++) g.ch.push_back(Channel(int(fh.xsize), int(fh.ysize)));
  }
  for (size_t i = 0; i < m.ec.size(); i++) { JXLO_CHECK(fh.ec_upsampling[i] == 1 && fh.upsampling == 1, "upsampling not supported"); uint32_t s = m.ec[i].dim_shift; g.ch.push_back(Channel(int(DivCeil(fh.xsize, 1u << s)), int(DivCeil(fh.ysize, 1u << s)), int(s), int(s))); }
  fs.global_decoded = 0;
  if (!g.ch.empty()) {
    // ModularDecode with undo_transforms=false; channels larger than group_dim are left for the group sections.
    ModularOptions opt; opt.max_chan_size = int(fh.group_dim);
    fs.gheader = ModularDecode(br, g, 0, fs.has_tree ? &fs.tree : nullptr, fs.has_tree ? &fs.tree_code : nullptr, opt, false);
    size_t c = size_t(g.nb_meta); for (; c < g.ch.size(); c++) if (g.ch[c].w > int(fh.group_dim) || g.ch[c].h > int(fh.group_dim)) break;
    fs.global_decoded = c;
  }
}

// Decodes the group-local part of the global modular image: channels with min(hshift,vshift) in [min_shift,max_shift].
inline void DecodeModularGroup(FrameState& fs, BitReader& br, int x0, int y0, int xs, int ys, int min_shift, int max_shift, uint32_t stream_id) {
  ModularImage& full = fs.gimg; ModularImage gi; gi.bitdepth = full.bitdepth; std::vector<size_t> idx; std::vector<std::array<int, 2>> origin;
  for (size_t c = fs.global_decoded; c < full.ch.size(); c++) {
    const Channel& fc = full.ch[c]; int shift = std::min(fc.hshift, fc.vshift); if (shift > max_shift || shift < min_shift) continue;
    int rx0 = x0 >> fc.hshift, ry0 = y0 >> fc.vshift, rxs = xs >> fc.hshift, rys = ys >> fc.vshift;
    if (rx0 >= fc.w || ry0 >= fc.h) continue; rxs = std::min(rxs, fc.w - rx0); rys = std::min(rys, fc.h - ry0); if (rxs <= 0 || rys <= 0) continue;
    gi.ch.push_back(Channel(rxs, rys, fc.hshift, fc.vshift)); idx.push_back(c); origin.push_back({rx0, ry0});
  }
  if (gi.ch.empty()) return;
  ModularDecode(br, gi, stream_id, fs.has_tree ? &fs.tree : nullptr, fs.has_tree ? &fs.tree_code : nullptr, ModularOptions(), true);
  JXLO_CHECK(gi.ch.size() == idx.size(), "group-local transforms changed the channel count");
  for (size_t k = 0; k < idx.size(); k++) { Channel& fc = full.ch[idx[k]]; const Channel& gc = gi.ch[k]; JXLO_CHECK(gc.w + origin[k][0] <= fc.w && gc.h + origin[k][1] <= fc.h, "group channel geometry");
    for (int y = 0; y < gc.h; y++) memcpy(fc.row(origin[k][1] + y) + origin[k][0], gc.row(y), sizeof(int32_t) * size_t(gc.w)); }
}

// ------------------------------------------------------------------ LfGroup
inline void DecodeLfGroup(FrameState& fs, BitReader& br, uint32_t g) {
  const FrameHeader& fh = fs.fh; int gx = int(g % fh.xlfgroups), gy = int(g / fh.xlfgroups);
  int cx0 = gx * 256, cy0 = gy * 256;   // in 8x8 cells (group_dim 256 for VarDCT)
  if (fh.encoding == 0) {
    JXLO_CHECK(!(fh.flags & kFlagUseLfFrame), "LF frames are not supported");
    int w = std::min(256, fs.xb - cx0), h = std::min(256, fs.yb - cy0);
    uint32_t extra_prec = br.ReadBits(2);
    ModularImage img; img.bitdepth = 16; for (int c = 0; c < 3; c++) img.ch.push_back(Channel(w, h));
    ModularDecode(br, img, StreamIdLfCoeff(fh, g), fs.has_tree ? &fs.tree : nullptr, fs.has_tree ? &fs.tree_code : nullptr, ModularOptions(), true);
    float inv = fs.q.InvGlobalScale() / float(fs.q.quant_lf), mul = 1.0f / float(1u << extra_prec);
    float fac[3]; for (int c = 0; c < 3; c++) fac[c] = fs.lf_dequant[c] * inv * mul;
    float cx = fs.cfl.YtoX(fs.cfl.x_factor_lf), cb = fs.cfl.YtoB(fs.cfl.b_factor_lf);
    const BlockCtxMap& b = fs.bctx;
    for (int y = 0; y < h; y++) {
      const int32_t* qy = img.ch[0].row(y); const int32_t* qx = img.ch[1].row(y); const int32_t* qb = img.ch[2].row(y);
      for (int x = 0; x < w; x++) {
        size_t o = size_t(cy0 + y) * fs.xb + cx0 + x; fs.lfq[0][o] = qx[x]; fs.lfq[1][o] = qy[x]; fs.lfq[2][o] = qb[x];
        float Y = float(qy[x]) * fac[1]; fs.lf[1].d[o] = Y; fs.lf[0].d[o] = float(qx[x]) * fac[0] + cx * Y; fs.lf[2].d[o] = float(qb[x]) * fac[2] + cb * Y;
        uint32_t bx_ = 0, by_ = 0, bb_ = 0; for (int32_t t : b.lf_thr[0]) if (qx[x] > t) bx_++; for (int32_t t : b.lf_thr[1]) if (qy[x] > t) by_++; for (int32_t t : b.lf_thr[2]) if (qb[x] > t) bb_++;
        fs.lf_idx[o] = uint8_t((bx_ * (b.lf_thr[2].size() + 1) + bb_) * (b.lf_thr[1].size() + 1) + by_);
      }
    }
  }
  { const int lf_dim = int(fh.group_dim) * 8;   // an LF group is 8 x 8 groups: 2048 px for VarDCT, 1024 .. 8192 px for Modular frames (group_size_shift)
    DecodeModularGroup(fs, br, gx * lf_dim, gy * lf_dim, lf_dim, lf_dim, 3, 1000, StreamIdModularLf(fh, g)); }
  if (fh.encoding == 0) {
    int w = std::min(256, fs.xb - cx0), h = std::min(256, fs.yb - cy0);
    uint32_t nb = br.ReadBits(CeilLog2(uint64_t(w) * h)) + 1;
    int tw = (w + 7) / 8, th = (h + 7) / 8;
    ModularImage img; img.bitdepth = 8; img.ch.push_back(Channel(tw, th, 3, 3)); img.ch.push_back(Channel(tw, th, 3, 3)); img.ch.push_back(Channel(int(nb), 2)); img.ch.push_back(Channel(w, h));
    ModularDecode(br, img, StreamIdHfMeta(fh, g), fs.has_tree ? &fs.tree : nullptr, fs.has_tree ? &fs.tree_code : nullptr, ModularOptions(), true);
    for (int y = 0; y < th; y++) for (int x = 0; x < tw; x++) { int32_t a = img.ch[0].row(y)[x], b2 = img.ch[1].row(y)[x]; JXLO_CHECK(a >= -128 && a <= 127 && b2 >= -128 && b2 <= 127, "CfL factor range");
      size_t o = size_t(cy0 / 8 + y) * fs.xt + cx0 / 8 + x; fs.ytox[o] = int8_t(a); fs.ytob[o] = int8_t(b2); }
    uint32_t num = 0;
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
      size_t o = size_t(cy0 + y) * fs.xb + cx0 + x; if (fs.strategy[o] != 255) continue;
      JXLO_CHECK(num < nb, "HF metadata: more blocks than announced");
      int32_t s = img.ch[2].row(0)[num]; JXLO_CHECK(s >= 0 && s < kNumStrategies, "invalid AC strategy");
      int bw = kCoveredX[s], bh = kCoveredY[s];
      JXLO_CHECK(x + bw <= w && y + bh <= h, "AC strategy block out of bounds");
      JXLO_CHECK((x % 32) + bw <= 32 && (y % 32) + bh <= 32, "AC strategy block crosses a group boundary");
      int32_t qf = 1 + std::max(0, std::min(255, img.ch[2].row(1)[num]));
      for (int iy = 0; iy < bh; iy++) for (int ix = 0; ix < bw; ix++) { size_t p = o + size_t(iy) * fs.xb + ix; JXLO_CHECK(fs.strategy[p] == 255, "overlapping AC strategy blocks"); fs.strategy[p] = uint8_t(s); fs.is_first[p] = 0; fs.hf_mul[p] = qf; }
      fs.is_first[o] = 1; num++;
    }
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { int32_t v = img.ch[3].row(y)[x]; JXLO_CHECK(v >= 0 && v < 8, "EPF sharpness range"); fs.sharp[size_t(cy0 + y) * fs.xb + cx0 + x] = uint8_t(v); }
  }
}

// ------------------------------------------------------------------ HfGlobal
inline QuantEncoding ReadQuantEncoding(BitReader& br, int t) {
  auto read_params = [&](DctParams& p) { p.num_bands = int(br.ReadBits(4)) + 1; for (int c = 0; c < 3; c++) { for (int i = 0; i < p.num_bands; i++) p.bands[c][i] = br.F16(); JXLO_CHECK(p.bands[c][0] >= 1e-8f, "distance band"); p.bands[c][0] *= 64.0f; } };
  QuantEncoding e; e.mode = int(br.ReadBits(3));
  switch (e.mode) {
    case kQModeLibrary: return LibraryEncoding(t);
    case kQModeId: JXLO_CHECK(kTableRows[t] == 1 && kTableCols[t] == 1, "quant mode/table mismatch"); for (int c = 0; c < 3; c++) for (int i = 0; i < 3; i++) e.idw[c][i] = br.F16() * 64.0f; break;
    case kQModeDCT2: JXLO_CHECK(kTableRows[t] == 1 && kTableCols[t] == 1, "quant mode/table mismatch"); for (int c = 0; c < 3; c++) for (int i = 0; i < 6; i++) e.dct2w[c][i] = br.F16() * 64.0f; break;
    case kQModeDCT4: JXLO_CHECK(kTableRows[t] == 1 && kTableCols[t] == 1, "quant mode/table mismatch"); for (int c = 0; c < 3; c++) for (int i = 0; i < 2; i++) e.dct4mul[c][i] = br.F16(); read_params(e.dct4); break;
    case kQModeDCT4x8: JXLO_CHECK(kTableRows[t] == 1 && kTableCols[t] == 1, "quant mode/table mismatch"); for (int c = 0; c < 3; c++) e.dct4x8mul[c] = br.F16(); read_params(e.dct4x8); break;
    case kQModeDCT: read_params(e.dct); break;
    default: throw Error("AFV / RAW quant table encodings are not supported");
  }
  return e;
}
inline void DecodeHfGlobal(FrameState& fs, BitReader& br) {
  const FrameHeader& fh = fs.fh;
  bool all_default = br.Bool(); fs.dequant.resize(kNumQuantTables);
  for (int t = 0; t < kNumQuantTables; t++) { QuantEncoding e = all_default ? LibraryEncoding(t) : ReadQuantEncoding(br, t); fs.dequant[t] = ComputeDequantTable(t, e); }
  fs.num_hf_presets = 1 + br.ReadBits(CeilLog2(fh.num_groups));
  fs.natural.resize(kNumOrders); for (int o = 0; o < kNumOrders; o++) { int s = kOrderStrategy[o]; int r = std::min(kCoveredX[s], kCoveredY[s]), c = std::max(kCoveredX[s], kCoveredY[s]); fs.natural[o] = NaturalOrder(r, c); }
  fs.orders.assign(fh.passes.num_passes, {}); fs.ac_codes.resize(fh.passes.num_passes);
  for (uint32_t p = 0; p < fh.passes.num_passes; p++) {
    uint32_t used = br.U32(Val(0x5F), Val(0x13), Val(0), Bits(13)); fs.orders[p].assign(kNumOrders * 3, {});
    if (used) {
      Code c = DecodeCode(br, 8); SymbolReader r(&c, &br);
      for (int o = 0; o < kNumOrders; o++) if (used >> o & 1) for (int ch = 0; ch < 3; ch++) {
        size_t size = fs.natural[o].size(); std::vector<uint32_t> perm = ReadPermutation(r, size / 64, size); auto& out = fs.orders[p][o * 3 + ch]; out.resize(size);
        for (size_t k = 0; k < size; k++) out[k] = fs.natural[o][perm[k]];
      }
      JXLO_CHECK(r.CheckFinal(), "coefficient order ANS final state");
    }
    fs.ac_codes[p] = DecodeCode(br, size_t(495) * fs.num_hf_presets * fs.bctx.num_ctxs);
  }
}

// ------------------------------------------------------------------ PassGroup: AC coefficients
// coeffs: 3 planes of 256x256 ints for this group in "cell-chunked" layout: storage position p of the varblock
// whose first cell is (by,bx) lives in the (p/64)-th covered cell (raster order inside the block), see CoefAddr.
inline size_t CoefAddr(int by, int bx, int bw, uint32_t p) { uint32_t j = p >> 6; return (size_t(by + int(j) / bw) * 32 + size_t(bx + int(j) % bw)) * 64 + (p & 63); }
inline void DecodeAcGroup(FrameState& fs, BitReader& br, uint32_t pass, uint32_t g, std::vector<int32_t>* coeffs /*[3]*/, std::vector<uint16_t>* nzeros /*[3] 32x32*/) {
  const FrameHeader& fh = fs.fh; int gx = int(g % fh.xgroups), gy = int(g / fh.xgroups); int cx0 = gx * 32, cy0 = gy * 32;
  int w = std::min(32, fs.xb - cx0), h = std::min(32, fs.yb - cy0);
  uint32_t preset = br.ReadBits(CeilLog2(fs.num_hf_presets)); JXLO_CHECK(preset < fs.num_hf_presets, "HF preset");
  const Code& code = fs.ac_codes[pass]; SymbolReader rd(&code, &br);
  uint32_t nbctx = fs.bctx.num_ctxs; uint32_t ctx_offset = 495 * nbctx * preset;
  uint32_t shift = pass + 1 < fh.passes.num_passes ? fh.passes.shift[pass] : 0;
  for (int c = 0; c < 3; c++) nzeros[c].assign(32 * 32, 0);
  for (int by = 0; by < h; by++) for (int bx = 0; bx < w; bx++) {
    size_t o = size_t(cy0 + by) * fs.xb + cx0 + bx; if (!fs.is_first[o]) continue;
    int s = fs.strategy[o]; int bw = kCoveredX[s], bh = kCoveredY[s]; uint32_t covered = uint32_t(bw * bh), log2c = uint32_t(FloorLog2(covered)), size = covered * 64; int ord = kStrategyOrder[s];
    for (int ci = 0; ci < 3; ci++) {
      int c = ci == 0 ? 1 : ci == 1 ? 0 : 2;
      uint32_t pred; { uint16_t* nz = nzeros[c].data(); if (bx == 0) pred = by == 0 ? 32 : nz[(by - 1) * 32 + bx]; else if (by == 0) pred = nz[by * 32 + bx - 1]; else pred = (uint32_t(nz[(by - 1) * 32 + bx]) + nz[by * 32 + bx - 1] + 1) / 2; }
      uint32_t bctx = fs.bctx.Context(fs.lf_idx[o], uint32_t(fs.hf_mul[o]), uint32_t(ord), uint32_t(c));
      uint32_t nz = rd.Read(ctx_offset + NonZeroCtxBucket(pred) * nbctx + bctx);
      JXLO_CHECK(nz + covered <= size, "too many non-zero coefficients");
      { uint16_t v = uint16_t((nz + covered - 1) >> log2c); for (int iy = 0; iy < bh; iy++) for (int ix = 0; ix < bw; ix++) nzeros[c][(by + iy) * 32 + bx + ix] = v; }
      const std::vector<uint32_t>& order = fs.orders[pass][ord * 3 + c].empty() ? fs.natural[ord] : fs.orders[pass][ord * 3 + c];
      uint32_t histo = ctx_offset + nbctx * kNonZeroBuckets + kZeroDensityContextCount * bctx; uint32_t prev = nz > size / 16 ? 0 : 1;
      for (uint32_t k = covered; k < size && nz != 0; k++) {
        uint32_t u = rd.Read(histo + ZeroDensityContext(nz, k, covered, log2c, prev));
        int32_t coef = int32_t(uint32_t(UnpackSigned(u)) << shift); coeffs[c][CoefAddr(by, bx, bw, order[k])] += coef; prev = u != 0; nz -= prev;
      }
      JXLO_CHECK(nz == 0, "non-zero count mismatch"); JXLO_CHECK(!br.overrun, "AC group truncated");
    }
  }
  JXLO_CHECK(rd.CheckFinal(), "AC group ANS final state");
}

}  // namespace jxlo
