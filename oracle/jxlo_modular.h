// ORACLE — test infrastructure only (see jxlo_bits.h header). PARITY UNPINNED.
// Modular sub-bitstream: MA tree, 14 predictors incl. the self-correcting Weighted
// predictor, RCT / palette / squeeze inverse transforms. Integer, bit-exact by
// construction. Restates SURVEY.md Appendix A.7; reached in the reference only through
// JxlDecoderProcessInput (N/Decoder/JxlDecoder.cpp:252) — LF image, HF metadata, alpha
// and whole-image lossless all go through this code.
#pragma once
#include "jxlo_entropy.h"

namespace jxlo {

struct Channel {
  int w = 0, h = 0, hshift = 0, vshift = 0; std::vector<int32_t> d;
  Channel() {}
  Channel(int w_, int h_, int hs = 0, int vs = 0) : w(w_), h(h_), hshift(hs), vshift(vs), d(size_t(w_) * h_, 0) {}
  int32_t* row(int y) { return d.data() + size_t(y) * w; }
  const int32_t* row(int y) const { return d.data() + size_t(y) * w; }
};

struct SqueezeParams { bool horizontal = false, in_place = false; uint32_t begin_c = 0, num_c = 1; };
struct Transform {
  int id = 0;                 // 0 RCT, 1 Palette, 2 Squeeze
  uint32_t begin_c = 0, rct_type = 6, num_c = 0, nb_colors = 0, nb_deltas = 0, predictor = 0;
  std::vector<SqueezeParams> squeezes;
};
struct WPHeader { int p1 = 16, p2 = 10, p3a = 7, p3b = 7, p3c = 7, p3d = 0, p3e = 0; int w[4] = {13, 12, 12, 12}; };
struct GroupHeader { bool use_global_tree = false; WPHeader wp; std::vector<Transform> transforms; };

struct TreeNode { int property = -1; int32_t splitval = 0; int lchild = 0, rchild = 0; int predictor = 0; int32_t offset = 0; uint32_t multiplier = 1; int leaf_id = 0; };
typedef std::vector<TreeNode> Tree;

struct ModularImage {
  std::vector<Channel> ch; int nb_meta = 0; int bitdepth = 8;
};

static const int kNumNonrefProps = 16;

// ------------------------------------------------------------------ headers
inline uint32_t ReadBeginC(BitReader& br) { return br.U32(Bits(3), BitsOffset(6, 8), BitsOffset(10, 72), BitsOffset(13, 1096)); }
inline void WriteBeginC(BitWriter& bw, uint32_t v) { bw.U32(Bits(3), BitsOffset(6, 8), BitsOffset(10, 72), BitsOffset(13, 1096), v); }

inline GroupHeader ReadGroupHeader(BitReader& br) {
  GroupHeader g; g.use_global_tree = br.Bool();
  if (!br.Bool()) { WPHeader& w = g.wp; w.p1 = br.ReadBits(5); w.p2 = br.ReadBits(5); w.p3a = br.ReadBits(5); w.p3b = br.ReadBits(5); w.p3c = br.ReadBits(5); w.p3d = br.ReadBits(5); w.p3e = br.ReadBits(5);
    for (int i = 0; i < 4; i++) w.w[i] = br.ReadBits(4); }
  uint32_t nt = br.U32(Val(0), Val(1), BitsOffset(4, 2), BitsOffset(8, 18));
  for (uint32_t i = 0; i < nt; i++) {
    Transform t; t.id = br.ReadBits(2); JXLO_CHECK(t.id < 3, "invalid transform id");
    if (t.id == 0) { t.begin_c = ReadBeginC(br); t.rct_type = br.U32(Val(6), Bits(2), BitsOffset(4, 2), BitsOffset(6, 10)); JXLO_CHECK(t.rct_type < 42, "rct type"); }
    else if (t.id == 1) { t.begin_c = ReadBeginC(br); t.num_c = br.U32(Val(1), Val(3), Val(4), BitsOffset(13, 1));
      t.nb_colors = br.U32(BitsOffset(8, 0), BitsOffset(10, 256), BitsOffset(12, 1280), BitsOffset(16, 5376));
      t.nb_deltas = br.U32(Val(0), BitsOffset(8, 1), BitsOffset(10, 257), BitsOffset(16, 1281)); t.predictor = br.ReadBits(4); JXLO_CHECK(t.predictor < 14, "palette predictor"); }
    else { uint32_t ns = br.U32(Val(0), BitsOffset(4, 1), BitsOffset(6, 9), BitsOffset(8, 41));
      for (uint32_t k = 0; k < ns; k++) { SqueezeParams s; s.horizontal = br.Bool(); s.in_place = br.Bool(); s.begin_c = ReadBeginC(br); s.num_c = br.U32(Val(1), Val(2), Val(3), BitsOffset(4, 4)); t.squeezes.push_back(s); } }
    g.transforms.push_back(t);
  }
  return g;
}
inline void WriteGroupHeader(BitWriter& bw, const GroupHeader& g) {
  bw.Bool(g.use_global_tree); bw.Bool(true);   // default WP header (the oracle encoder never customises it)
  bw.U32(Val(0), Val(1), BitsOffset(4, 2), BitsOffset(8, 18), uint32_t(g.transforms.size()));
  for (const Transform& t : g.transforms) {
    bw.Write(2, t.id);
    if (t.id == 0) { WriteBeginC(bw, t.begin_c); bw.U32(Val(6), Bits(2), BitsOffset(4, 2), BitsOffset(6, 10), t.rct_type); }
    else if (t.id == 1) { WriteBeginC(bw, t.begin_c); bw.U32(Val(1), Val(3), Val(4), BitsOffset(13, 1), t.num_c);
      bw.U32(BitsOffset(8, 0), BitsOffset(10, 256), BitsOffset(12, 1280), BitsOffset(16, 5376), t.nb_colors);
      bw.U32(Val(0), BitsOffset(8, 1), BitsOffset(10, 257), BitsOffset(16, 1281), t.nb_deltas); bw.Write(4, t.predictor); }
    else { bw.U32(Val(0), BitsOffset(4, 1), BitsOffset(6, 9), BitsOffset(8, 41), uint32_t(t.squeezes.size()));
      for (auto& s : t.squeezes) { bw.Bool(s.horizontal); bw.Bool(s.in_place); WriteBeginC(bw, s.begin_c); bw.U32(Val(1), Val(2), Val(3), BitsOffset(4, 4), s.num_c); } }
  }
}

// ------------------------------------------------------------------ MA tree
inline Tree DecodeTree(BitReader& br, size_t size_limit) {
  Code code = DecodeCode(br, 6); SymbolReader r(&code, &br); Tree tree; size_t to_decode = 1; int leaf = 0;
  while (to_decode > 0) {
    JXLO_CHECK(tree.size() < size_limit, "MA tree too large"); JXLO_CHECK(!br.overrun, "MA tree truncated");
    to_decode--;
    int prop = int(r.Read(1)) - 1; JXLO_CHECK(prop < 256, "MA tree property");
    TreeNode n;
    if (prop < 0) {
      n.property = -1; n.predictor = int(r.Read(2)); JXLO_CHECK(n.predictor < 14, "MA tree predictor");
      n.offset = UnpackSigned(r.Read(3)); uint32_t ml = r.Read(4); JXLO_CHECK(ml < 31, "MA multiplier log");
      uint32_t mb = r.Read(5); JXLO_CHECK(mb < (1u << (31 - ml)) - 1, "MA multiplier bits"); n.multiplier = (mb + 1) << ml; n.leaf_id = leaf++;
    } else {
      n.property = prop; n.splitval = UnpackSigned(r.Read(0));
      n.lchild = int(tree.size() + to_decode + 1); n.rchild = int(tree.size() + to_decode + 2); to_decode += 2;
    }
    tree.push_back(n);
  }
  JXLO_CHECK(r.CheckFinal(), "MA tree ANS final state");
  return tree;
}
inline size_t NumLeaves(const Tree& t) { return (t.size() + 1) / 2; }

// tree must already be in BFS order (children allocated as the decoder expects)
inline void TokenizeTree(const Tree& tree, std::vector<Token>* out) {
  for (const TreeNode& n : tree) {
    if (n.property < 0) {
      out->push_back({1, 0}); out->push_back({2, uint32_t(n.predictor)}); out->push_back({3, PackSigned(n.offset)});
      uint32_t ml = 0, m = n.multiplier; while ((m & 1) == 0) { m >>= 1; ml++; }
      out->push_back({4, ml}); out->push_back({5, m - 1});
    } else { out->push_back({1, uint32_t(n.property + 1)}); out->push_back({0, PackSigned(n.splitval)}); }
  }
}
// Builds a BFS-ordered tree from nested thresholds on a list of properties: level i splits on props[i]
// with thresholds[i] (ascending); every leaf uses `predictor`.
inline Tree MakeFixedTree(const std::vector<int>& props, const std::vector<std::vector<int32_t>>& thresholds, int predictor) {
  struct Tmp { int prop; int32_t split; int l, r; };
  std::vector<Tmp> tmp;
  struct Builder { const std::vector<int>& props; const std::vector<std::vector<int32_t>>& thr; std::vector<Tmp>& tmp;
    int Build(size_t level, int lo, int hi) {
      if (level >= props.size()) { tmp.push_back({-1, 0, -1, -1}); return int(tmp.size()) - 1; }
      if (lo >= hi) return Build(level + 1, 0, level + 1 < props.size() ? int(thr[level + 1].size()) : 0);
      int mid = (lo + hi) / 2; int id = int(tmp.size()); tmp.push_back({props[level], thr[level][mid], -1, -1});
      int l = Build(level, mid + 1, hi); int r = Build(level, lo, mid); tmp[id].l = l; tmp[id].r = r; return id;
    } } b{props, thresholds, tmp};
  int root = b.Build(0, 0, props.empty() ? 0 : int(thresholds[0].size()));
  Tree tree; std::vector<int> queue; queue.push_back(root); size_t head = 0; int leaf = 0;
  while (head < queue.size()) {   // BFS order == the order DecodeTree allocates children in
    const Tmp& t = tmp[queue[head++]]; TreeNode n;
    if (t.prop < 0) { n.property = -1; n.predictor = predictor; n.leaf_id = leaf++; }
    else { n.property = t.prop; n.splitval = t.split; n.lchild = int(queue.size()); n.rchild = int(queue.size()) + 1; queue.push_back(t.l); queue.push_back(t.r); }
    tree.push_back(n);
  }
  return tree;
}

// ------------------------------------------------------------------ weighted predictor (A.7 [L])
struct WPState {
  WPHeader h; int xsize; std::vector<uint32_t> pred_errors[4]; std::vector<int32_t> error; int64_t prediction[4]; int64_t pred = 0;
  uint32_t divlookup[64];
  WPState(const WPHeader& hdr, int xs) : h(hdr), xsize(xs) {
    for (auto& p : pred_errors) p.assign(size_t(xs + 2) * 2, 0); error.assign(size_t(xs + 2) * 2, 0);
    for (int i = 0; i < 64; i++) divlookup[i] = (1u << 24) / (i + 1);
    for (auto& p : prediction) p = 0;
  }
  uint32_t ErrorWeight(uint64_t x, uint32_t maxweight) const {
    int shift = FloorLog2(x + 1) - 5; if (shift < 0) shift = 0;
    return 4 + ((maxweight * divlookup[x >> shift]) >> shift);
  }
  // returns prediction; *max_err receives property 15
  int64_t Predict(int x, int y, int64_t N, int64_t W, int64_t NE, int64_t NW, int64_t NN, int32_t* max_err) {
    size_t cur = (y & 1) ? 0 : size_t(xsize + 2), prev = (y & 1) ? size_t(xsize + 2) : 0;
    size_t pos_N = prev + x, pos_NE = x < xsize - 1 ? pos_N + 1 : pos_N, pos_NW = x > 0 ? pos_N - 1 : pos_N;
    uint32_t weights[4];
    for (int i = 0; i < 4; i++) weights[i] = ErrorWeight(uint64_t(pred_errors[i][pos_N]) + pred_errors[i][pos_NE] + pred_errors[i][pos_NW], uint32_t(h.w[i]));
    N *= 8; W *= 8; NE *= 8; NW *= 8; NN *= 8;
    int64_t teW = x == 0 ? 0 : error[cur + x - 1], teN = error[pos_N], teNW = error[pos_NW], sumWN = teN + teW, teNE = error[pos_NE];
    if (max_err) { int64_t p = teW; if (std::llabs(teN) > std::llabs(p)) p = teN; if (std::llabs(teNW) > std::llabs(p)) p = teNW; if (std::llabs(teNE) > std::llabs(p)) p = teNE; *max_err = int32_t(p); }
    prediction[0] = W + NE - N;
    prediction[1] = N - (((sumWN + teNE) * h.p1) >> 5);
    prediction[2] = W - (((sumWN + teNW) * h.p2) >> 5);
    prediction[3] = N - ((teNW * h.p3a + teN * h.p3b + teNE * h.p3c + (NN - N) * h.p3d + (NW - W) * h.p3e) >> 5);
    uint32_t wsum = 0; for (int i = 0; i < 4; i++) wsum += weights[i];
    int lw = FloorLog2(wsum); wsum = 0; for (int i = 0; i < 4; i++) { weights[i] >>= lw - 4; wsum += weights[i]; }
    int64_t sum = (wsum >> 1) - 1; for (int i = 0; i < 4; i++) sum += prediction[i] * int64_t(weights[i]);
    pred = (sum * int64_t(divlookup[wsum - 1])) >> 24;
    if (((teN ^ teW) | (teN ^ teNW)) > 0) return (pred + 3) >> 3;
    int64_t mx = std::max(W, std::max(NE, N)), mn = std::min(W, std::min(NE, N));
    pred = std::max(mn, std::min(mx, pred));
    return (pred + 3) >> 3;
  }
  void Update(int64_t val, int x, int y) {
    size_t cur = (y & 1) ? 0 : size_t(xsize + 2), prev = (y & 1) ? size_t(xsize + 2) : 0;
    val *= 8; error[cur + x] = int32_t(pred - val);
    for (int i = 0; i < 4; i++) { uint32_t err = uint32_t((std::llabs(prediction[i] - val) + 3) >> 3); pred_errors[i][cur + x] = err; pred_errors[i][prev + x + 1] += err; }
  }
};

// ------------------------------------------------------------------ per-pixel prediction machinery
struct Neigh { int64_t W, N, NW, NE, NN, WW, NEE; };
inline Neigh GetNeigh(const int32_t* p, int x, int y, int w) {  // p = current row pointer; rows contiguous
  Neigh n; const int32_t* up = p - w; const int32_t* up2 = up - w;
  n.W = x ? p[x - 1] : (y ? up[x] : 0); n.N = y ? up[x] : n.W; n.NW = (x && y) ? up[x - 1] : n.W;
  n.NE = (x + 1 < w && y) ? up[x + 1] : n.N; n.NN = y > 1 ? up2[x] : n.N; n.WW = x > 1 ? p[x - 2] : n.W;
  n.NEE = (x + 2 < w && y) ? up[x + 2] : n.NE; return n;
}
inline int64_t ClampedGradient(int64_t W, int64_t N, int64_t NW) { int64_t lo = std::min(W, N), hi = std::max(W, N); return std::max(lo, std::min(hi, W + N - NW)); }
inline int64_t Predict(int pred, const Neigh& n, int64_t wp) {
  switch (pred) {
    case 0: return 0; case 1: return n.W; case 2: return n.N; case 3: return (n.W + n.N) / 2;
    case 4: { int64_t p = n.W + n.N - n.NW; return std::llabs(p - n.W) < std::llabs(p - n.N) ? n.W : n.N; }
    case 5: return ClampedGradient(n.W, n.N, n.NW); case 6: return wp; case 7: return n.NE; case 8: return n.NW; case 9: return n.WW;
    case 10: return (n.W + n.NW) / 2; case 11: return (n.N + n.NW) / 2; case 12: return (n.N + n.NE) / 2;
    case 13: return (6 * n.N - 2 * n.NN + 7 * n.W + n.WW + n.NEE + 3 * n.NE + 8) / 16;
  }
  throw Error("predictor");
}

struct TreeInfo { int max_prop = 0; bool uses_wp = false; };
inline TreeInfo Analyze(const Tree& t) { TreeInfo i; for (auto& n : t) { if (n.property >= 0) { i.max_prop = std::max(i.max_prop, n.property); if (n.property == 15) i.uses_wp = true; } else if (n.predictor == 6) i.uses_wp = true; } return i; }

// Walks every pixel of channel `ci` in raster order. CB(x, y, leaf, prediction) must return the pixel value.
template <class CB>
inline void ForEachPixel(const Tree& tree, const TreeInfo& ti, const WPHeader& wph, ModularImage& img, int ci, uint32_t stream_id, CB&& cb) {
  Channel& c = img.ch[ci]; if (!c.w || !c.h) return;
  std::vector<int> refs;   // earlier channels with identical geometry, nearest first
  int nref = ti.max_prop >= kNumNonrefProps ? (ti.max_prop - kNumNonrefProps) / 4 + 1 : 0;
  for (int j = ci - 1; j >= 0 && int(refs.size()) < nref; j--) { const Channel& r = img.ch[j]; if (r.w == c.w && r.h == c.h && r.hshift == c.hshift && r.vshift == c.vshift) refs.push_back(j); }
  std::vector<int32_t> props(size_t(kNumNonrefProps + 4 * nref), 0);
  std::unique_ptr<WPState> wp; if (ti.uses_wp) wp.reset(new WPState(wph, c.w));
  props[0] = ci; props[1] = int32_t(stream_id);
  for (int y = 0; y < c.h; y++) {
    int32_t* p = c.row(y); props[2] = y; props[9] = 0;
    for (int x = 0; x < c.w; x++) {
      Neigh n = GetNeigh(p, x, y, c.w);
      props[3] = x; props[4] = int32_t(std::llabs(n.N)); props[5] = int32_t(std::llabs(n.W)); props[6] = int32_t(n.N); props[7] = int32_t(n.W);
      props[8] = int32_t(n.W - props[9]); props[9] = int32_t(n.W + n.N - n.NW); props[10] = int32_t(n.W - n.NW); props[11] = int32_t(n.NW - n.N);
      props[12] = int32_t(n.N - n.NE); props[13] = int32_t(n.N - n.NN); props[14] = int32_t(n.W - n.WW);
      int64_t wpred = 0; if (wp) { int32_t me; wpred = wp->Predict(x, y, n.N, n.W, n.NE, n.NW, n.NN, &me); props[15] = me; }
      for (size_t k = 0; k < refs.size(); k++) {
        const Channel& r = img.ch[refs[k]]; const int32_t* rp = r.row(y); int64_t v = rp[x];
        int64_t rW = x ? rp[x - 1] : 0, rN = y ? rp[x - r.w] : rW, rNW = (x && y) ? rp[x - 1 - r.w] : rW; int64_t g = ClampedGradient(rW, rN, rNW);
        size_t o = kNumNonrefProps + 4 * k; props[o] = int32_t(std::llabs(v)); props[o + 1] = int32_t(v); props[o + 2] = int32_t(std::llabs(v - g)); props[o + 3] = int32_t(v - g);
      }
      int node = 0; while (tree[node].property >= 0) node = props[tree[node].property] > tree[node].splitval ? tree[node].lchild : tree[node].rchild;
      int64_t pred = Predict(tree[node].predictor, n, wpred);
      int32_t val = cb(x, y, tree[node], pred); p[x] = val;
      if (wp) wp->Update(val, x, y);
    }
  }
}

// ------------------------------------------------------------------ transforms
inline void DefaultSqueezeParams(std::vector<SqueezeParams>* out, const ModularImage& img) {
  int nb = int(img.ch.size()) - img.nb_meta; out->clear(); if (nb <= 0) return;
  int w = img.ch[img.nb_meta].w, h = img.ch[img.nb_meta].h;
  if (nb > 2 && img.ch[img.nb_meta + 1].w == w && img.ch[img.nb_meta + 1].h == h) {
    SqueezeParams p; p.horizontal = true; p.in_place = false; p.begin_c = img.nb_meta + 1; p.num_c = 2; out->push_back(p); p.horizontal = false; out->push_back(p);
  }
  SqueezeParams p; p.begin_c = img.nb_meta; p.num_c = nb; p.in_place = true;
  if (!(w > h)) { if (h > 8) { p.horizontal = false; out->push_back(p); h = (h + 1) / 2; } }
  while (w > 8 || h > 8) {
    if (w > 8) { p.horizontal = true; out->push_back(p); w = (w + 1) / 2; }
    if (h > 8) { p.horizontal = false; out->push_back(p); h = (h + 1) / 2; }
  }
}

// Applies the channel-list side of a transform (what the bitstream layout looks like after it).
inline void MetaApply(Transform& t, ModularImage& img) {
  if (t.id == 0) { JXLO_CHECK(t.begin_c + 3 <= img.ch.size(), "RCT channel range"); return; }
  if (t.id == 1) {
    uint32_t end_c = t.begin_c + t.num_c - 1; JXLO_CHECK(end_c < img.ch.size() && t.num_c >= 1, "palette channel range");
    if (int(t.begin_c) < img.nb_meta) { JXLO_CHECK(int(end_c) < img.nb_meta, "palette across meta boundary"); img.nb_meta += 2 - int(t.num_c); } else img.nb_meta += 1;
    img.ch.erase(img.ch.begin() + t.begin_c + 1, img.ch.begin() + end_c + 1);
    Channel pal(int(t.nb_colors + t.nb_deltas), int(t.num_c), -1, -1); img.ch.insert(img.ch.begin(), pal); return;
  }
  if (t.squeezes.empty()) DefaultSqueezeParams(&t.squeezes, img);
  for (const SqueezeParams& s : t.squeezes) {
    uint32_t end_c = s.begin_c + s.num_c - 1; JXLO_CHECK(end_c < img.ch.size(), "squeeze channel range");
    if (int(s.begin_c) < img.nb_meta) { JXLO_CHECK(s.in_place && int(end_c) < img.nb_meta, "squeeze on meta channels"); img.nb_meta += int(s.num_c); }
    size_t offset = s.in_place ? end_c + 1 : img.ch.size();
    for (uint32_t c = s.begin_c; c <= end_c; c++) {
      Channel& ch = img.ch[c]; Channel res;
      if (s.horizontal) { int w = ch.w; ch = Channel((w + 1) / 2, ch.h, ch.hshift + 1, ch.vshift); res = Channel(w - (w + 1) / 2, ch.h, ch.hshift, ch.vshift); }
      else { int h = ch.h; ch = Channel(ch.w, (h + 1) / 2, ch.hshift, ch.vshift + 1); res = Channel(ch.w, h - (h + 1) / 2, ch.hshift, ch.vshift); }
      img.ch.insert(img.ch.begin() + offset + (c - s.begin_c), res);
    }
  }
}

inline int64_t SmoothTendency(int64_t B, int64_t a, int64_t n) {
  int64_t diff = 0;
  if (B >= a && a >= n) { diff = (4 * B - 3 * n - a + 6) / 12; if (diff - (diff & 1) > 2 * (B - a)) diff = 2 * (B - a) + 1; if (diff + (diff & 1) > 2 * (a - n)) diff = 2 * (a - n); }
  else if (B <= a && a <= n) { diff = (4 * B - 3 * n - a - 6) / 12; if (diff + (diff & 1) < 2 * (B - a)) diff = 2 * (B - a) - 1; if (diff - (diff & 1) < 2 * (a - n)) diff = 2 * (a - n); }
  return diff;
}

inline void InverseRCT(ModularImage& img, uint32_t begin_c, uint32_t type) {
  uint32_t perm = type / 7, k = type % 7; Channel& c0 = img.ch[begin_c]; Channel& c1 = img.ch[begin_c + 1]; Channel& c2 = img.ch[begin_c + 2];
  JXLO_CHECK(c0.w == c1.w && c0.w == c2.w && c0.h == c1.h && c0.h == c2.h, "RCT channel sizes");
  size_t n = c0.d.size();
  for (size_t i = 0; i < n; i++) {
    int32_t A = c0.d[i], B = c1.d[i], C = c2.d[i], o0, o1, o2;
    if (k == 6) { int32_t t = A - (C >> 1); int32_t G = C + t; int32_t Bl = t - (B >> 1); int32_t R = Bl + B; o0 = R; o1 = G; o2 = Bl; }
    else { int32_t D = A, E = B, F = C; if (k & 1) F += A; if ((k >> 1) == 1) E += A; if ((k >> 1) == 2) E += (A + F) >> 1; o0 = D; o1 = E; o2 = F; }
    c0.d[i] = o0; c1.d[i] = o1; c2.d[i] = o2;
  }
  if (perm) {
    Channel t0 = std::move(img.ch[begin_c]), t1 = std::move(img.ch[begin_c + 1]), t2 = std::move(img.ch[begin_c + 2]);
    img.ch[begin_c + perm % 3] = std::move(t0); img.ch[begin_c + (perm + 1 + perm / 3) % 3] = std::move(t1); img.ch[begin_c + (perm + 2 - perm / 3) % 3] = std::move(t2);
  }
}
inline void ForwardRCT_YCgCo(const int32_t R, const int32_t G, const int32_t B, int32_t* Y, int32_t* Co, int32_t* Cg) {
  *Co = R - B; int32_t t = B + (*Co >> 1); *Cg = G - t; *Y = t + (*Cg >> 1);
}

inline void InversePalette(ModularImage& img, const Transform& t, const WPHeader& wph) {
  // [L] only explicit palette entries (0 <= index < nb_colors+nb_deltas) and the implicit colour cubes are
  // restated; the 72-entry delta palette for negative indices is not recalled and is rejected.
  int nb = int(t.num_c); Channel pal = img.ch[0]; uint32_t c0 = t.begin_c + 1; int psize = pal.w; int bitdepth = img.bitdepth;
  Channel idx = img.ch[c0]; JXLO_CHECK(pal.h == nb, "palette geometry");
  for (int i = 1; i < nb; i++) img.ch.insert(img.ch.begin() + c0 + 1, Channel(idx.w, idx.h, idx.hshift, idx.vshift));
  for (int c = 0; c < nb; c++) {
    Channel& out = img.ch[c0 + c];
    for (int y = 0; y < idx.h; y++) for (int x = 0; x < idx.w; x++) {
      int index = idx.row(y)[x]; int32_t v;
      JXLO_CHECK(index >= 0, "delta-palette (negative index) not supported by the oracle");
      if (index < psize) v = pal.row(c)[index];
      else if (index < psize + 64) { int i2 = index - psize; int div = c == 0 ? 1 : c == 1 ? 4 : 16; if (c > 2) v = 0; else v = int32_t(((int64_t((i2 / div) % 4) * ((int64_t(1) << bitdepth) - 1)) >> 2) + (int64_t(1) << std::max(0, bitdepth - 3))); }
      else { int i2 = index - psize - 64; int div = c == 0 ? 1 : c == 1 ? 5 : 25; if (c > 2) v = 0; else v = int32_t((int64_t((i2 / div) % 5) * ((int64_t(1) << bitdepth) - 1)) >> 2); }
      out.row(y)[x] = v;
    }
    if (t.nb_deltas > 0) {
      for (int y = 0; y < idx.h; y++) for (int x = 0; x < idx.w; x++) {
        int index = idx.row(y)[x]; if (index >= int(t.nb_deltas)) continue;
        Neigh n = GetNeigh(out.row(y), x, y, out.w); JXLO_CHECK(t.predictor != 6, "palette delta with weighted predictor not supported by the oracle");
        out.row(y)[x] = int32_t(out.row(y)[x] + Predict(int(t.predictor), n, 0));
      }
    }
  }
  img.ch.erase(img.ch.begin()); img.nb_meta--; (void)wph;
}

inline void InverseSqueeze(ModularImage& img, const Transform& t) {
  for (size_t i = t.squeezes.size(); i-- > 0;) {
    const SqueezeParams& s = t.squeezes[i]; uint32_t end_c = s.begin_c + s.num_c - 1;
    size_t offset = s.in_place ? end_c + 1 : img.ch.size() - s.num_c;
    if (int(s.begin_c) < img.nb_meta) img.nb_meta -= int(s.num_c);
    for (uint32_t c = s.begin_c; c <= end_c; c++) {
      size_t rc = offset + (c - s.begin_c); Channel& avg = img.ch[c]; Channel& res = img.ch[rc];
      if (s.horizontal) {
        Channel out(avg.w + res.w, avg.h, avg.hshift - 1, avg.vshift);
        for (int y = 0; y < out.h; y++) { const int32_t* pa = avg.row(y); const int32_t* pr = res.w ? res.row(y) : nullptr; int32_t* po = out.row(y);
          for (int x = 0; x < res.w; x++) { int64_t a = pa[x], nx = x + 1 < avg.w ? pa[x + 1] : a, left = x ? po[(x << 1) - 1] : a; int64_t diff = pr[x] + SmoothTendency(left, a, nx); int64_t A = a + diff / 2; po[x << 1] = int32_t(A); po[(x << 1) + 1] = int32_t(A - diff); }
          if (out.w & 1) po[out.w - 1] = pa[avg.w - 1]; }
        img.ch[c] = std::move(out);
      } else {
        Channel out(avg.w, avg.h + res.h, avg.hshift, avg.vshift - 1);
        for (int y = 0; y < res.h; y++) { const int32_t* pa = avg.row(y); const int32_t* pn = y + 1 < avg.h ? avg.row(y + 1) : pa; const int32_t* pr = res.row(y);
          int32_t* po = out.row(y << 1); int32_t* po1 = out.row((y << 1) + 1); const int32_t* pt = y ? out.row((y << 1) - 1) : nullptr;
          for (int x = 0; x < out.w; x++) { int64_t a = pa[x], nx = pn[x], top = y ? pt[x] : a; int64_t diff = pr[x] + SmoothTendency(top, a, nx); int64_t A = a + diff / 2; po[x] = int32_t(A); po1[x] = int32_t(A - diff); } }
        if (out.h & 1) memcpy(out.row(out.h - 1), avg.row(avg.h - 1), sizeof(int32_t) * size_t(out.w));
        img.ch[c] = std::move(out);
      }
    }
    img.ch.erase(img.ch.begin() + offset, img.ch.begin() + offset + s.num_c);
  }
}

inline void UndoTransforms(ModularImage& img, const GroupHeader& g) {
  for (size_t i = g.transforms.size(); i-- > 0;) {
    const Transform& t = g.transforms[i];
    if (t.id == 0) InverseRCT(img, t.begin_c, t.rct_type); else if (t.id == 1) InversePalette(img, t, g.wp); else InverseSqueeze(img, t);
  }
}

// ------------------------------------------------------------------ sub-bitstream decode / encode
struct ModularOptions { int max_chan_size = 1 << 30; };

// Decodes the channels of `img` (geometry preset by the caller). Global tree/code are used when the
// header says so. Returns the header; transforms are undone when undo_transforms is set.
inline GroupHeader ModularDecode(BitReader& br, ModularImage& img, uint32_t stream_id, const Tree* gtree, const Code* gcode,
                                 const ModularOptions& opt, bool undo_transforms) {
  GroupHeader g; if (img.ch.empty()) return g;
  g = ReadGroupHeader(br);
  for (Transform& t : g.transforms) MetaApply(t, img);
  {  // nothing to decode here (all channels belong to group sections): no tree, no ANS state word
    size_t num = 0; for (size_t i = 0; i < img.ch.size(); i++) { const Channel& c = img.ch[i]; if (int(i) >= img.nb_meta && (c.w > opt.max_chan_size || c.h > opt.max_chan_size)) break; if (c.w && c.h) num++; }
    if (num == 0) return g;
  }
  Tree ltree; Code lcode; const Tree* tree = gtree; const Code* code = gcode;
  if (!g.use_global_tree) {
    size_t px = 0; for (auto& c : img.ch) px += size_t(c.w) * c.h;
    ltree = DecodeTree(br, std::min<size_t>(size_t(1) << 22, 1024 + px)); lcode = DecodeCode(br, NumLeaves(ltree)); tree = &ltree; code = &lcode;
  } else JXLO_CHECK(gtree && gcode && !gtree->empty(), "global MA tree missing");
  TreeInfo ti = Analyze(*tree);
  uint32_t dist_mult = 0; size_t nch = 0;
  for (size_t i = 0; i < img.ch.size(); i++) { Channel& c = img.ch[i]; if (int(i) >= img.nb_meta && (c.w > opt.max_chan_size || c.h > opt.max_chan_size)) break; nch = i + 1; if (uint32_t(c.w) > dist_mult) dist_mult = uint32_t(c.w); }
  SymbolReader rd(code, &br);
  for (size_t i = 0; i < nch; i++) {
    ForEachPixel(*tree, ti, g.wp, img, int(i), stream_id, [&](int, int, const TreeNode& leaf, int64_t pred) -> int32_t {
      uint32_t v = rd.Read(uint32_t(leaf.leaf_id), dist_mult);
      return int32_t(int64_t(UnpackSigned(v)) * int64_t(leaf.multiplier) + leaf.offset + pred);
    });
    JXLO_CHECK(!br.overrun, "modular stream truncated");
  }
  JXLO_CHECK(rd.CheckFinal(), "modular ANS final state");
  if (undo_transforms) UndoTransforms(img, g);
  return g;
}

// Encoder side: tokenises channels [first, last) of img against a tree (offset 0 / multiplier 1 leaves).
inline void ModularTokenize(ModularImage& img, size_t first, size_t last, uint32_t stream_id, const Tree& tree, std::vector<Token>* out) {
  TreeInfo ti = Analyze(tree); WPHeader wph;
  for (size_t i = first; i < last; i++) {
    ForEachPixel(tree, ti, wph, img, int(i), stream_id, [&](int x, int y, const TreeNode& leaf, int64_t pred) -> int32_t {
      int32_t v = img.ch[i].row(y)[x]; int64_t r = int64_t(v) - pred - leaf.offset; JXLO_CHECK(leaf.multiplier == 1, "encoder trees use multiplier 1");
      out->push_back({uint32_t(leaf.leaf_id), PackSigned(int32_t(r))}); return v;
    });
  }
}

}  // namespace jxlo
