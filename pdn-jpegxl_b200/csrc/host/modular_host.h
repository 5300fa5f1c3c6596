// pdn-jpegxl_b200 engine — host-side bitstream front-end (Modular group headers, MA tree decode/tokenise).
// Part of the product: parses what must be parsed serially on the CPU and hands flat tables
// to the sm_100a kernels. Replaces libjxl work reached from N/Decoder/JxlDecoder.cpp:252,454
// and N/Encoder/JxlEncoder.cpp:128,367 of the reference. Per ISO/IEC 18181-1 as digested in
// SURVEY.md Appendix A.7. Constants tagged [M]/[L] there are unverified against libjxl.
#pragma once
#include "entropy.h"

namespace jxlgpu {

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
    Transform t; t.id = br.ReadBits(2); JXLG_CHECK(t.id < 3, "invalid transform id");
    if (t.id == 0) { t.begin_c = ReadBeginC(br); t.rct_type = br.U32(Val(6), Bits(2), BitsOffset(4, 2), BitsOffset(6, 10)); JXLG_CHECK(t.rct_type < 42, "rct type"); }
    else if (t.id == 1) { t.begin_c = ReadBeginC(br); t.num_c = br.U32(Val(1), Val(3), Val(4), BitsOffset(13, 1));
      t.nb_colors = br.U32(BitsOffset(8, 0), BitsOffset(10, 256), BitsOffset(12, 1280), BitsOffset(16, 5376));
      t.nb_deltas = br.U32(Val(0), BitsOffset(8, 1), BitsOffset(10, 257), BitsOffset(16, 1281)); t.predictor = br.ReadBits(4); JXLG_CHECK(t.predictor < 14, "palette predictor"); }
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
    JXLG_CHECK(tree.size() < size_limit, "MA tree too large"); JXLG_CHECK(!br.overrun, "MA tree truncated");
    to_decode--;
    int prop = int(r.Read(1)) - 1; JXLG_CHECK(prop < 256, "MA tree property");
    TreeNode n;
    if (prop < 0) {
      n.property = -1; n.predictor = int(r.Read(2)); JXLG_CHECK(n.predictor < 14, "MA tree predictor");
      n.offset = UnpackSigned(r.Read(3)); uint32_t ml = r.Read(4); JXLG_CHECK(ml < 31, "MA multiplier log");
      uint32_t mb = r.Read(5); JXLG_CHECK(mb < (1u << (31 - ml)) - 1, "MA multiplier bits"); n.multiplier = (mb + 1) << ml; n.leaf_id = leaf++;
    } else {
      n.property = prop; n.splitval = UnpackSigned(r.Read(0));
      n.lchild = int(tree.size() + to_decode + 1); n.rchild = int(tree.size() + to_decode + 2); to_decode += 2;
    }
    tree.push_back(n);
  }
  JXLG_CHECK(r.CheckFinal(), "MA tree ANS final state");
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


}  // namespace jxlgpu
