// pdn-jpegxl_b200 engine — device-side Modular channel decoder (MA-tree walk, 14 predictors
// incl. the self-correcting Weighted predictor, residual reconstruction). Integer, bit-exact.
// One thread decodes one sub-bitstream; the neighbourhood W/WW/N/NW/NE/NEE rolls through
// registers so each pixel costs two global loads (NEE, NN) plus the entropy-coded symbol.
// Replaces libjxl's DecodeModularChannelMAANS reached from N/Decoder/JxlDecoder.cpp:252
// (SURVEY.md A.7). Used for the LF image, HF metadata, extra channels and lossless frames.
#pragma once
#include "common.cuh"

namespace jxlgpu {

static const uint32_t kMaxWpWidth = 2048;   // widest channel the weighted-predictor scratch is sized for
struct DWPHeader { int32_t p1, p2, p3a, p3b, p3c, p3d, p3e; int32_t w[4]; };
// tree node: decision {x=property>=0, y=splitval, z=lchild, w=rchild}; leaf {x=-1, y=(ctx<<4)|predictor, z=offset, w=multiplier}
typedef int4 DTreeNode;

#ifdef __CUDACC__
struct WPScratch {   // global memory, 5 arrays of 2*(w+2) ints each (A.7 weighted predictor state)
  uint32_t* pe[4]; int32_t* error; int xs2;
  __device__ void Bind(int32_t* base, int w) { xs2 = w + 2; for (int i = 0; i < 4; i++) pe[i] = reinterpret_cast<uint32_t*>(base) + size_t(i) * 2 * xs2; error = base + size_t(4) * 2 * xs2; }
  __device__ void Clear() { for (int i = 0; i < 2 * xs2; i++) { pe[0][i] = pe[1][i] = pe[2][i] = pe[3][i] = 0; error[i] = 0; } }
};
__device__ __forceinline__ size_t WPScratchInts(int w) { return size_t(5) * 2 * (w + 2); }

__device__ __forceinline__ uint32_t WPErrorWeight(uint64_t x, uint32_t maxweight) {
  int shift = (63 - __clzll((long long)(x + 1))) - 5; if (shift < 0) shift = 0;
  uint32_t div = (1u << 24) / (uint32_t(x >> shift) + 1u);
  return 4 + ((maxweight * div) >> shift);
}

struct WPPred { long long prediction[4]; long long pred; int32_t max_err; };

__device__ __forceinline__ long long WPPredict(const DWPHeader& h, WPScratch& s, WPPred& o, int x, int y, int xsize, long long N, long long W, long long NE, long long NW, long long NN) {
  size_t cur = (y & 1) ? 0 : size_t(s.xs2), prev = (y & 1) ? size_t(s.xs2) : 0;
  size_t pos_N = prev + x, pos_NE = x < xsize - 1 ? pos_N + 1 : pos_N, pos_NW = x > 0 ? pos_N - 1 : pos_N;
  uint32_t weights[4];
#pragma unroll
  for (int i = 0; i < 4; i++) weights[i] = WPErrorWeight(uint64_t(s.pe[i][pos_N]) + s.pe[i][pos_NE] + s.pe[i][pos_NW], uint32_t(h.w[i]));
  N *= 8; W *= 8; NE *= 8; NW *= 8; NN *= 8;
  long long teW = x == 0 ? 0 : s.error[cur + x - 1], teN = s.error[pos_N], teNW = s.error[pos_NW], sumWN = teN + teW, teNE = s.error[pos_NE];
  long long p = teW; if (llabs(teN) > llabs(p)) p = teN; if (llabs(teNW) > llabs(p)) p = teNW; if (llabs(teNE) > llabs(p)) p = teNE; o.max_err = int32_t(p);
  o.prediction[0] = W + NE - N;
  o.prediction[1] = N - (((sumWN + teNE) * h.p1) >> 5);
  o.prediction[2] = W - (((sumWN + teNW) * h.p2) >> 5);
  o.prediction[3] = N - ((teNW * h.p3a + teN * h.p3b + teNE * h.p3c + (NN - N) * h.p3d + (NW - W) * h.p3e) >> 5);
  uint32_t wsum = weights[0] + weights[1] + weights[2] + weights[3]; int lw = 31 - __clz(wsum); wsum = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) { weights[i] >>= lw - 4; wsum += weights[i]; }
  long long sum = (long long)(wsum >> 1) - 1;
#pragma unroll
  for (int i = 0; i < 4; i++) sum += o.prediction[i] * (long long)weights[i];
  o.pred = (sum * (long long)((1u << 24) / wsum)) >> 24;
  if (((teN ^ teW) | (teN ^ teNW)) > 0) return (o.pred + 3) >> 3;
  long long mx = max(W, max(NE, N)), mn = min(W, min(NE, N));
  o.pred = max(mn, min(mx, o.pred));
  return (o.pred + 3) >> 3;
}
__device__ __forceinline__ void WPUpdate(WPScratch& s, const WPPred& o, long long val, int x, int y) {
  size_t cur = (y & 1) ? 0 : size_t(s.xs2), prev = (y & 1) ? size_t(s.xs2) : 0;
  val *= 8; s.error[cur + x] = int32_t(o.pred - val);
#pragma unroll
  for (int i = 0; i < 4; i++) { uint32_t err = uint32_t((llabs(o.prediction[i] - val) + 3) >> 3); s.pe[i][cur + x] = err; s.pe[i][prev + x + 1] += err; }
}

// Per-channel lookup table for the common case of an MA subtree that tests a single dynamic property with uniform
// leaves (predictor fixed, offset 0, multiplier 1): context cluster = lut[#thresholds below the property value].
// When every threshold lies in [-128, 126] the lookup is a direct 256-entry table on the clamped property value.
struct ChanLut { int32_t thr[32]; uint16_t cluster[33]; uint8_t direct[256]; int32_t n, prop, predictor, ok, has_direct;
  uint2 dinfo[256]; };   // dinfo[v+128] = {info word of the cluster for property value v, byte offset of its alias rows}: one LDS.64 replaces direct[] -> info[] -> address math

// T = int32_t when all sums of four samples fit 32 bits (bit depth <= 20), else int64_t (libjxl's pixel_type_w).
template <typename T> struct ModMath {
  static __device__ __forceinline__ T Abs(T v) { return v < 0 ? -v : v; }
  static __device__ __forceinline__ T PropValue(int p, int chan, int stream_id, int x, int y, T N, T W, T NW, T NE, T NN, T WW, int32_t prev_grad, int32_t wp_err, uint32_t* err) {
    switch (p) {
      case 0: return chan; case 1: return stream_id; case 2: return y; case 3: return x; case 4: return Abs(N); case 5: return Abs(W); case 6: return N; case 7: return W;
      case 8: return W - prev_grad; case 9: return W + N - NW; case 10: return W - NW; case 11: return NW - N; case 12: return N - NE; case 13: return N - NN; case 14: return W - WW; case 15: return wp_err;
      default: *err = kErrRefProps; return 0;
    }
  }
  static __device__ __forceinline__ T Gradient(T N, T W, T NW) { T lo = min(W, N), hi = max(W, N); return max(lo, min(hi, W + N - NW)); }
  static __device__ __forceinline__ T Prediction(int pr, T N, T W, T NW, T NE, T NN, T WW, T NEE, T wpred) {
    switch (pr) {
      case 0: return 0; case 1: return W; case 2: return N; case 3: return (W + N) / 2;
      case 4: { T p = W + N - NW; return Abs(p - W) < Abs(p - N) ? W : N; }
      case 5: return Gradient(N, W, NW);
      case 6: return wpred; case 7: return NE; case 8: return NW; case 9: return WW; case 10: return (W + NW) / 2; case 11: return (N + NW) / 2; case 12: return (N + NE) / 2;
      default: return (6 * N - 2 * NN + 7 * W + WW + NEE + 3 * NE + 8) / 16;
    }
  }
};

// Hand-over between the lane that parses (lane 0) and the warp-collective speculative loop (DecodeRowsLeanSpec): reader state, the
// distinct clusters ("slots", at most 32: one per lane) this channel's contexts can select, and the context value -> slot map.
struct LeanSpecPrep { BitRd br; uint32_t state, err, K, ok; uint32_t slot_info[32], slot_alias_off[32]; uint8_t slot_of[256]; };

// Transforms listed in a sub-bitstream's own header (group sections): RCTs and palettes without delta entries are undone on the device (what
// libjxl's lossless encoder picks per group). kind 0 RCT: a = rct_type. kind 1 Palette: a = num_c, b = number of colours.
struct GroupTransforms { static const int kMax = 4; uint32_t n = 0, kind[kMax], begin[kMax], a[kMax], b[kMax]; };

struct ModDecoder {
  SymReader rd; CodeView cv; const DTreeNode* tree; DWPHeader wp; bool uses_wp; bool wide; uint32_t dist_mult = 0; /* LZ77: widest channel of the sub-bitstream */ ChanLut* lut;
  // Earlier channels of the same sub-bitstream with the geometry of the channel being decoded, nearest first (MA-tree properties 16 + 4k .. 19 + 4k:
  // |v|, v, |v - g|, v - g of that channel's sample at the same position, g its clamped gradient). The caller keeps the list (NoteChannel).
  GroupTransforms* gt = nullptr;   // where ReadGroupHeaderDev leaves the transforms of the stream's own header (null: the caller cannot undo any)
  static const int kMaxRefs = 4; const int32_t* ref_p[kMaxRefs]; size_t ref_stride[kMaxRefs]; int ref_n = 0;
  struct Seen { const int32_t* p; size_t stride; int w, h, hs, vs; }; static const int kMaxSeen = 8; Seen seen[kMaxSeen]; int num_seen = 0;   // the most recent channels (a squeezed image has dozens)
  __device__ void ResetChannels() { num_seen = 0; ref_n = 0; }
  // call before decoding a channel: selects its reference channels among those noted so far, then notes the channel itself
  __device__ void NoteChannel(const int32_t* p, size_t stride, int w, int h, int hs, int vs) {
    ref_n = 0; for (int j = num_seen - 1; j >= 0 && ref_n < kMaxRefs; j--) if (seen[j].w == w && seen[j].h == h && seen[j].hs == hs && seen[j].vs == vs) { ref_p[ref_n] = seen[j].p; ref_stride[ref_n] = seen[j].stride; ref_n++; }
    if (num_seen == kMaxSeen) { for (int j = 1; j < kMaxSeen; j++) seen[j - 1] = seen[j]; num_seen--; }
    if (num_seen < kMaxSeen) { seen[num_seen].p = p; seen[num_seen].stride = stride; seen[num_seen].w = w; seen[num_seen].h = h; seen[num_seen].hs = hs; seen[num_seen].vs = vs; num_seen++; }
  }
  __device__ __forceinline__ int32_t RefProp(int p, int x, int y) const {
    const int k = (p - 16) >> 2, which = (p - 16) & 3; if (k >= ref_n) return 0;   // fewer matching channels than the tree asks for: the property reads 0
    const int32_t* rp = ref_p[k] + size_t(y) * ref_stride[k]; const long long v = rp[x];
    const long long rW = x ? rp[x - 1] : 0, rN = y ? rp[x - ptrdiff_t(ref_stride[k])] : rW, rNW = (x && y) ? rp[x - 1 - ptrdiff_t(ref_stride[k])] : rW;
    const long long lo = min(rW, rN), hi = max(rW, rN), g = max(lo, min(hi, rW + rN - rNW)), d = v - g;
    return int32_t(which == 0 ? (v < 0 ? -v : v) : which == 1 ? v : which == 2 ? (d < 0 ? -d : d) : d);
  }   // lut: per-decoding-thread scratch (shared memory)

  // In-order walk of the subtree under `root` (<= branch first): ascending thresholds, leaves per interval.
  __device__ void BuildLut(int root) {
    ChanLut& L = *lut; L.ok = 0; L.n = 0; L.prop = -1; L.predictor = -1; L.has_direct = 0; int nleaf = 0;
    int stack[40]; uint8_t state[40]; int sp = 0; stack[0] = root; state[0] = 0;
    while (sp >= 0) {
      DTreeNode nd = tree[stack[sp]];
      if (nd.x < 0) {   // leaf
        if (nd.z != 0 || nd.w != 1) return; int pr = nd.y & 15; if (L.predictor < 0) L.predictor = pr; else if (L.predictor != pr) return;
        if (nleaf > 32) return; L.cluster[nleaf++] = cv.ctx_map[uint32_t(nd.y) >> 4]; sp--; continue;
      }
      if (nd.x < 2 || nd.x > 15) return; if (L.prop < 0) L.prop = nd.x; else if (L.prop != nd.x) return;
      if (state[sp] == 0) { state[sp] = 1; if (sp >= 38) return; stack[sp + 1] = nd.w; state[sp + 1] = 0; sp++; }
      else if (state[sp] == 1) { state[sp] = 2; if (L.n >= 32) return; L.thr[L.n++] = nd.y; stack[sp + 1] = nd.z; state[sp + 1] = 0; sp++; }
      else sp--;
    }
    if (nleaf != L.n + 1) return;
    for (int i = 1; i < L.n; i++) if (L.thr[i] < L.thr[i - 1]) return;
    if (L.prop < 0) L.prop = 2;   // single leaf: any property, zero thresholds
    bool direct = true; for (int i = 0; i < L.n; i++) if (L.thr[i] < -128 || L.thr[i] > 126) direct = false;
    if (direct) { for (int v = -128; v <= 127; v++) { int cnt = 0; for (int i = 0; i < L.n; i++) cnt += v > L.thr[i]; const uint32_t cl = L.cluster[cnt]; L.direct[v + 128] = uint8_t(cl);
        L.dinfo[v + 128] = make_uint2(cv.info[cl], (cl << cv.log_alpha) * 8u); } L.has_direct = 1; }
    L.ok = 1;
  }

  // kMode 0: generic tree walk; 1: LUT fast path (any property/predictor); 2: LUT fast path specialised for property 8 + gradient predictor
  // kWp: the weighted predictor may be in use (only the wide instantiations carry its code and registers)
  template <typename T, int kMode, bool kSmem, bool kWp>
  __device__ __noinline__ void DecodeRows(int root, DTreeNode n, int chan, int stream_id, int32_t* out, size_t stride, int w, int h, int32_t* wp_base) {
    typedef ModMath<T> M; WPScratch ws; WPPred wo; wo.max_err = 0; if (kWp && this->uses_wp) { ws.Bind(wp_base, w); ws.Clear(); }
    SymReader rd = this->rd; CodeView cv = this->cv; const DTreeNode* tree = this->tree; const DWPHeader wp = this->wp; const bool uses_wp = kWp && this->uses_wp; ChanLut* lut = this->lut;   // registers, not *this
    struct WriteBack { SymReader& dst; SymReader& src; __device__ ~WriteBack() { dst = src; } } wb{this->rd, rd};
    __builtin_assume(__isShared(lut)); if (kSmem) cv.AssumeShared();
    const ChanLut& L = *lut; const int fprop = L.prop, fpred = L.predictor, fn = L.n; const bool direct = L.has_direct != 0;
    const bool need_nn = uses_wp || kMode == 0 || (kMode == 1 && (fprop == 13 || fpred == 13));
    for (int y = 0; y < h; y++) {
      int32_t* cur = out + size_t(y) * stride; const int32_t* up = cur - stride; const int32_t* up2 = up - stride;
      T W = 0, WW = 0, N, NW, NE, NEE; int32_t prev_grad = 0;
      if (y) { N = up[0]; NE = w > 1 ? up[1] : N; NEE = w > 2 ? up[2] : NE; W = N; NW = W; WW = W; } else { N = NW = NE = NEE = 0; W = WW = 0; }
      if (kMode == 1 && fpred == 0 && !uses_wp && (fn == 0 || fprop == 2)) {   // zero-entropy row: the context cluster has a single symbol and the predictor is Zero
        int cnt = 0; for (int i = 0; i < fn; i++) cnt += y > L.thr[i]; const uint32_t info = cv.info[L.cluster[cnt]]; const uint32_t t = info >> 16;
        if (t != 0xffffu && t < (1u << (info & 0xff))) { const int32_t cval = UnpackSignedDev(t); for (int x = 0; x < w; x++) cur[x] = cval; continue; }
        // one cluster for the whole row and no prediction: the samples do not depend on each other, only the ANS state chain is serial
        // (HF metadata of frames with variable blocks: strategies and quantiser multipliers are coded this way)
        { const uint32_t cl = L.cluster[cnt]; for (int x = 0; x < w; x++) cur[x] = UnpackSignedDev(kSmem ? rd.ReadClusterAns(cv, cl) : rd.ReadCluster(cv, cl)); continue; }
      }
      for (int x = 0; x < w; x++) {
        const T nee_next = (y && x + 3 < w) ? T(up[x + 3]) : T(0);   // issued early: independent of the symbol being decoded
        T NN = (need_nn && y > 1) ? T(up2[x]) : N;
        T wpred = 0; if (uses_wp) wpred = T(WPPredict(wp, ws, wo, x, y, w, N, W, NE, NW, NN));
        uint32_t tok; T pred; int32_t offset = 0; uint32_t mult = 1;
        if (kMode != 0) {
          const int32_t v = kMode == 2 ? int32_t(W) - prev_grad : int32_t(M::PropValue(fprop, chan, stream_id, x, y, N, W, NW, NE, NN, WW, prev_grad, wo.max_err, &rd.err));
          uint32_t cl;
          if (direct) cl = L.direct[min(max(v, -128), 127) + 128]; else { int cnt = 0; for (int i = 0; i < fn; i++) cnt += v > L.thr[i]; cl = L.cluster[cnt]; }
          tok = rd.ReadCluster(cv, cl); pred = kMode == 2 ? M::Gradient(N, W, NW) : M::Prediction(fpred, N, W, NW, NE, NN, WW, NEE, wpred);
        } else {
          DTreeNode nd = n;
          while (nd.x >= 0) { int32_t v = nd.x >= 16 ? this->RefProp(nd.x, x, y) : int32_t(M::PropValue(nd.x, chan, stream_id, x, y, N, W, NW, NE, NN, WW, prev_grad, wo.max_err, &rd.err)); nd = tree[v > nd.y ? nd.z : nd.w]; }
          tok = cv.lz77 ? rd.ReadLz(cv, uint32_t(nd.y) >> 4, this->dist_mult) : rd.Read(cv, uint32_t(nd.y) >> 4); pred = M::Prediction(nd.y & 15, N, W, NW, NE, NN, WW, NEE, wpred); offset = nd.z; mult = uint32_t(nd.w);
        }
        const int32_t val = kMode != 0 ? int32_t(T(UnpackSignedDev(tok)) + pred) : int32_t((long long)UnpackSignedDev(tok) * (long long)mult + offset + (long long)pred);
        cur[x] = val;
        if (uses_wp) WPUpdate(ws, wo, val, x, y);
        prev_grad = int32_t(W + N - NW);
        { T oldW = W; W = val; WW = x >= 1 ? oldW : W; }
        if (y) { NW = N; N = NE; NE = NEE; NEE = (x + 3 < w) ? nee_next : NE; }
        else { N = W; NW = W; NE = W; NEE = W; }
      }
    }
  }

  // Lean path: ANS code with every table in shared memory, no weighted predictor, int32 samples, gradient predictor,
  // context = direct LUT of property 8 (previous residual). This is the LF-coefficient / alpha / lossless hot loop.
  __device__ __noinline__ void DecodeRowsLean(int32_t* out, size_t stride, int w, int h) {
    CodeView cv = this->cv; ChanLut* lut = this->lut; __builtin_assume(__isShared(lut)); cv.AssumeShared();
    // reader state in plain locals (nothing may take the reader's address here, or it all moves to local memory)
    BitRd br = this->rd.br; uint32_t state = this->rd.state, err = this->rd.err;
    const uint2* dinfo = lut->dinfo; const uint8_t* alias_bytes = reinterpret_cast<const uint8_t*>(cv.alias);
    const uint32_t log_entry = 12 - cv.log_alpha, pos_mask = (1u << log_entry) - 1;
    // One symbol: context word di = dinfo[clamp(W - prev_grad)] -> token -> residual. A single warp issues at best one instruction
    // every other cycle, so the loop is written for instruction count: the gradient predictor and property 8 only need W, N and NW,
    // i.e. ONE load per pixel (the next N, requested a full iteration early), and row 0 (no row above) has its own loop.
    auto symbol = [&](int32_t ctxv) -> int32_t {
      const uint2 di = dinfo[min(max(ctxv, -128), 127) + 128]; const uint32_t info = di.x;
      uint32_t tok = info >> 16;
      if (tok == 0xffffu) {   // ANS step (clusters with a single symbol carry it in the info word instead)
        const uint32_t idx = state & 0xfff, i = idx >> log_entry, pos = idx & pos_mask;
        const DAlias e = *reinterpret_cast<const DAlias*>(alias_bytes + di.y + i * 8u);
        const bool g = pos >= (e.x & 0xffu); tok = g ? ((e.x >> 8) & 0xffu) : i;
        const uint32_t s1 = (g ? (e.y >> 16) : (e.y & 0xffffu)) * (state >> 12) + (g ? (e.x >> 16) : 0u) + pos;
        const bool refill = s1 < 65536u; state = refill ? ((s1 << 16) | (br.Peek32() & 0xffffu)) : s1;   // branch-free renormalisation
        if (refill) br.Skip(16);
      }
      if (tok >= (1u << (info & 0xff))) {   // hybrid-uint tail (rare for LF residuals)
        DHybrid hc; hc.split_exp = uint8_t(info & 0xff); hc.msb = uint8_t((info >> 8) & 15); hc.lsb = uint8_t((info >> 12) & 15);
        const uint32_t split = 1u << hc.split_exp, n = hc.split_exp - (hc.msb + hc.lsb) + ((tok - split) >> (hc.msb + hc.lsb));
        if (n >= 32) { err = err ? err : kErrHybrid; tok = 0; }
        else { const uint32_t low = tok & ((1u << hc.lsb) - 1); const uint32_t t2 = tok >> hc.lsb; const uint32_t hi = (t2 & ((1u << hc.msb) - 1)) | (1u << hc.msb); tok = (((hi << n) | br.Read(int(n))) << hc.lsb) | low; }
      }
      return UnpackSignedDev(tok);
    };
    { int32_t W = 0, prev_grad = 0;   // row 0: N = NW = W, so the gradient is W itself
      for (int x = 0; x < w; x++) { const int32_t val = symbol(W - prev_grad) + W; out[x] = val; prev_grad = W; W = val; } }
    for (int y = 1; y < h; y++) {
      int32_t* __restrict__ cur = out + size_t(y) * stride; const int32_t* __restrict__ up = cur - stride;
      int32_t N = up[0], NW = N, W = N, prev_grad = 0; const int last = w - 1;
#pragma unroll 2
      for (int x = 0; x < w; x++) {
        const int32_t n_next = up[min(x + 1, last)];   // independent of the symbol being decoded: its latency hides behind the ANS step
        const int32_t res = symbol(W - prev_grad);
        const int32_t g = W + N - NW, val = res + max(min(W, N), min(max(W, N), g));
        cur[x] = val; prev_grad = g; NW = N; N = n_next; W = val;
      }
    }
    this->rd.br = br; this->rd.state = state; this->rd.err = err;
  }

  // Lane 0: decides whether channel `chan` can take the warp-collective loop (same conditions as DecodeRowsLean plus: at most 32
  // distinct clusters, transposed alias table fits `spec_bytes`), and if so fills P. Returns P.ok.
  __device__ bool PrepareLeanSpec(int chan, int stream_id, LeanSpecPrep& P, uint32_t spec_bytes) {
    P.ok = 0;
    int root = 0; DTreeNode n = tree[0];
    while (n.x == 0 || n.x == 1) { int v = n.x == 0 ? chan : stream_id; root = v > n.y ? n.z : n.w; n = tree[root]; }
    if (wide || uses_wp || cv.use_prefix || cv.lz77 || !cv.AllShared() || spec_bytes < (256u << cv.log_alpha)) return false;
    BuildLut(root); const ChanLut& L = *lut;
    if (!(L.ok && L.prop == 8 && L.predictor == 5 && L.has_direct)) return false;
    uint32_t K = 0;
    for (int v = 0; v < 256; v++) {
      const uint2 di = L.dinfo[v]; uint32_t k = 0; while (k < K && P.slot_alias_off[k] != di.y) k++;
      if (k == K) { if (K == 32) return false; P.slot_alias_off[K] = di.y; P.slot_info[K] = di.x; K++; }
      P.slot_of[v] = uint8_t(k);
    }
    P.K = K; P.br = rd.br; P.state = rd.state; P.err = rd.err; P.ok = 1; return true;
  }
  __device__ void FinishLeanSpec(const LeanSpecPrep& P) { rd.br = P.br; rd.state = P.state; rd.err = P.err; }

  // Decodes one channel in raster order into out[y*stride + x]. wp_base: scratch for the weighted predictor (may be null when !uses_wp).
  // kNarrow: the caller guarantees !uses_wp && !wide (checked on the host), so the 64-bit / weighted-predictor code is not instantiated
  template <bool kNarrow = false>
  __device__ void DecodeChannel(int chan, int stream_id, int32_t* out, size_t stride, int w, int h, int32_t* wp_base) {
    if (w <= 0 || h <= 0) return;
    // resolve static decisions (properties 0 = channel, 1 = stream id) at the top of the tree once per channel
    int root = 0; DTreeNode n = tree[0];
    while (n.x == 0 || n.x == 1) { int v = n.x == 0 ? chan : stream_id; root = v > n.y ? n.z : n.w; n = tree[root]; }
    if (uses_wp && w > int(kMaxWpWidth)) { rd.err = kErrUnsupportedStream; return; }
    BuildLut(root); const ChanLut& L = *lut;
    if (cv.lz77) {   // LZ77 streams: every value goes through the window, so only the generic tree walk reads them
      if (!kNarrow && (wide || uses_wp)) DecodeRows<long long, 0, false, true>(root, n, chan, stream_id, out, stride, w, h, wp_base);
      else if (wide || uses_wp) rd.err = kErrUnsupportedStream;
      else DecodeRows<int32_t, 0, false, false>(root, n, chan, stream_id, out, stride, w, h, wp_base);
      return;
    }
    const bool sm = cv.AllShared() && !cv.use_prefix;
    if (!kNarrow && (wide || uses_wp)) { if (L.ok) DecodeRows<long long, 1, false, true>(root, n, chan, stream_id, out, stride, w, h, wp_base); else DecodeRows<long long, 0, false, true>(root, n, chan, stream_id, out, stride, w, h, wp_base); }
    else if (wide || uses_wp) rd.err = kErrUnsupportedStream;
    else if (L.ok && sm && L.prop == 8 && L.predictor == 5 && L.has_direct) DecodeRowsLean(out, stride, w, h);
    else if (L.ok && L.prop == 8 && L.predictor == 5) DecodeRows<int32_t, 2, false, false>(root, n, chan, stream_id, out, stride, w, h, wp_base);
    else if (L.ok && sm) DecodeRows<int32_t, 1, true, false>(root, n, chan, stream_id, out, stride, w, h, wp_base);
    else if (L.ok) DecodeRows<int32_t, 1, false, false>(root, n, chan, stream_id, out, stride, w, h, wp_base);
    else DecodeRows<int32_t, 0, false, false>(root, n, chan, stream_id, out, stride, w, h, wp_base);
  }
};

// Warp-collective form of ModDecoder::DecodeRowsLean (ANS, gradient predictor, context = property 8 through a direct LUT). The serial
// chain of one symbol is  previous sample -> context -> cluster -> alias entry of (cluster, state) -> token + next state -> sample.
// Here every lane k < K decodes the CURRENT state under the assumption that the cluster is slot k (its own row of a transposed alias
// table: conflict-free LDS.64), while the context of the symbol is still being derived from the previous sample; two shuffles then pick
// the lane whose assumption was right. The alias lookup and the state arithmetic leave the critical path, which drops from ~280 to
// ~110 cycles per symbol. All lanes carry identical copies of the bit reader and of the neighbourhood; lane 0 stores the samples.
// T: [1 << log_alpha][32] uint2 in shared memory (built here from the staged alias rows).
static __device__ __noinline__ void DecodeRowsLeanSpec(LeanSpecPrep& P, const uint8_t* alias_bytes, uint32_t log_alpha, uint2* T, int32_t* out, size_t stride, int w, int h, int lane) {
  __builtin_assume(__isShared(T)); __builtin_assume(__isShared(&P));
  const uint32_t K = P.K, nent = 1u << log_alpha;
  for (uint32_t idx = uint32_t(lane); idx < nent * 32u; idx += 32u) { const uint32_t i = idx >> 5, k = idx & 31u;
    T[idx] = k < K ? *reinterpret_cast<const uint2*>(alias_bytes + P.slot_alias_off[k] + i * 8u) : make_uint2(0u, 0u); }
  const uint32_t my_info = P.slot_info[uint32_t(lane) < K ? lane : 0]; const bool my_const = (my_info >> 16) != 0xffffu; const uint32_t my_split = 1u << (my_info & 0xffu);
  BitRd br = P.br; uint32_t state = P.state, err = P.err;
  __syncwarp();
  const uint32_t log_entry = 12 - log_alpha, pos_mask = (1u << log_entry) - 1;
  // Software pipeline across symbols: the warp issues in order and the loop has a back edge, so the alias entry of the NEXT symbol is
  // requested as soon as the new ANS state exists (right after the shuffle), not at the top of the next iteration: its shared-memory
  // latency then runs beside the unpack / predict / store tail of the current symbol.
  const uint8_t* slot_of = P.slot_of; const uint2* Tl = T + lane;
  uint2 e = Tl[((state & 0xfff) >> log_entry) << 5];
  auto symbol = [&](int32_t ctxv) -> int32_t {
    const uint32_t slot = slot_of[min(max(ctxv, -128), 127) + 128];
    const uint32_t idx = state & 0xfff, i = idx >> log_entry, pos = idx & pos_mask;
    const bool g = pos >= (e.x & 0xffu);
    const uint32_t s1 = (g ? (e.y >> 16) : (e.y & 0xffffu)) * (state >> 12) + (g ? (e.x >> 16) : 0u) + pos;
    const bool refill = s1 < 65536u;
    uint32_t cand_state = refill ? ((s1 << 16) | (br.Peek32() & 0xffffu)) : s1;
    uint32_t cand = (g ? ((e.x >> 8) & 0xffu) : i) | (refill ? 0x100u : 0u);
    if (my_const) { cand = my_info >> 16; cand_state = state; }   // single-symbol cluster: no bits read, state unchanged
    if ((cand & 0xffu) >= my_split) cand |= 0x200u;                // token has a hybrid-uint tail
    const uint32_t sel = __shfl_sync(0xffffffffu, cand, int(slot));
    state = __shfl_sync(0xffffffffu, cand_state, int(slot));
    e = Tl[((state & 0xfff) >> log_entry) << 5];
    if (sel & 0x100u) br.Skip(16);
    uint32_t tok = sel & 0xffu;
    if (sel & 0x200u) {   // rare for LF residuals; uniform over the warp
      const uint32_t info = __shfl_sync(0xffffffffu, my_info, int(slot));
      const uint32_t se = info & 0xff, msb = (info >> 8) & 15, lsb = (info >> 12) & 15, split = 1u << se, n = se - (msb + lsb) + ((tok - split) >> (msb + lsb));
      if (n >= 32) { err = err ? err : kErrHybrid; tok = 0; }
      else { const uint32_t low = tok & ((1u << lsb) - 1); const uint32_t t2 = tok >> lsb; const uint32_t hi = (t2 & ((1u << msb) - 1)) | (1u << msb); tok = (((hi << n) | br.Read(int(n))) << lsb) | low; }
    }
    return UnpackSignedDev(tok);
  };
  // Rows move through registers in chunks of 32 samples: lane j keeps sample x0 + j of the row being decoded (one coalesced 128-byte
  // store per chunk) and sample x0 + j of the row above (one coalesced load per chunk, issued a whole chunk ahead of its first use:
  // at ~120 cycles per symbol a load issued one symbol ahead would sit on the critical path). Neighbours come from shuffles that do not
  // depend on the symbol being decoded.
  { int32_t W = 0, prev_grad = 0;   // row 0: N = NW = W, so the gradient is W itself
    for (int x0 = 0; x0 < w; x0 += 32) { const int cnt = min(32, w - x0); int32_t mine = 0;
      for (int j = 0; j < cnt; j++) { const int32_t val = symbol(W - prev_grad) + W; if (lane == j) mine = val; prev_grad = W; W = val; }
      if (lane < cnt) out[x0 + lane] = mine; } }
  for (int y = 1; y < h; y++) {
    int32_t* cur = out + size_t(y) * stride; const int32_t* up = cur - stride; const int last = w - 1;
    __syncwarp();   // row y-1 is stored
    int32_t upc = up[min(lane, last)];
    int32_t N = __shfl_sync(0xffffffffu, upc, 0), NW = N, W = N, prev_grad = 0;
    for (int x0 = 0; x0 < w; x0 += 32) {
      const int32_t upn = up[min(x0 + 32 + lane, last)]; const int cnt = min(32, w - x0); int32_t mine = 0;
      for (int j = 0; j < cnt; j++) {
        const int32_t n_next = __shfl_sync(0xffffffffu, j + 1 < 32 ? upc : upn, (j + 1) & 31);   // row above at min(x + 1, last)
        const int32_t res = symbol(W - prev_grad);
        const int32_t g = W + N - NW, val = res + max(min(W, N), min(max(W, N), g));
        if (lane == j) mine = val;
        prev_grad = g; NW = N; N = n_next; W = val;
      }
      if (lane < cnt) cur[x0 + lane] = mine;
      upc = upn;
    }
  }
  __syncwarp();
  if (lane == 0) { P.br = br; P.state = state; P.err = err; }
  __syncwarp();
}
#endif

// Palette index -> sample of colour channel c (SURVEY.md A.7 "Palette"): explicit entries [0, pal_w), then the implicit 4x4x4 cube (64 entries,
// offset by 2^(bitdepth-3)) and the implicit 5x5x5 cube. Negative indices address the 72-entry delta palette, whose table is not available
// offline: *bad is set for them.
__device__ __forceinline__ int32_t PaletteLookup(const int32_t* pal_row, int index, int c, int pal_w, int bitdepth, bool* bad) {
  if (index < 0) { *bad = true; return 0; }
  if (index < pal_w) return pal_row[index];
  if (c > 2) return 0;
  const long long maxv = (1ll << bitdepth) - 1;
  if (index < pal_w + 64) { const int i2 = index - pal_w, div = c == 0 ? 1 : c == 1 ? 4 : 16; return int32_t((((long long)((i2 / div) % 4) * maxv) >> 2) + (1ll << max(0, bitdepth - 3))); }
  const int i2 = index - pal_w - 64, div = c == 0 ? 1 : c == 1 ? 5 : 25; return int32_t(((long long)((i2 / div) % 5) * maxv) >> 2);
}

}  // namespace jxlgpu
