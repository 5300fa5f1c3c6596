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

struct ModDecoder {
  SymReader rd; CodeView cv; const DTreeNode* tree; DWPHeader wp; bool uses_wp; uint32_t max_prop;
  // Decodes one channel in raster order into out[y*stride + x]. wp_base: scratch for the weighted predictor (may be null when !uses_wp).
  __device__ void DecodeChannel(int chan, int stream_id, int32_t* out, size_t stride, int w, int h, int32_t* wp_base) {
    if (w <= 0 || h <= 0) return;
    // resolve static decisions (properties 0 = channel, 1 = stream id) at the top of the tree once per channel
    int root = 0; DTreeNode n = tree[0];
    while (n.x == 0 || n.x == 1) { int v = n.x == 0 ? chan : stream_id; root = v > n.y ? n.z : n.w; n = tree[root]; }
    if (uses_wp && w > int(kMaxWpWidth)) { rd.err = kErrUnsupportedStream; return; }
    WPScratch ws; WPPred wo; if (uses_wp) { ws.Bind(wp_base, w); ws.Clear(); }
    const bool single_leaf = n.x < 0;
    for (int y = 0; y < h; y++) {
      int32_t* cur = out + size_t(y) * stride; const int32_t* up = cur - stride; const int32_t* up2 = up - stride;
      long long W = 0, WW = 0, N, NW, NE, NEE; int32_t prev_grad = 0;
      // prime the rolling window for x = 0
      if (y) { N = up[0]; NE = w > 1 ? up[1] : N; NEE = w > 2 ? up[2] : NE; W = N; NW = W; WW = W; } else { N = NW = NE = NEE = 0; W = WW = 0; }
      for (int x = 0; x < w; x++) {
        long long NN = y > 1 ? (long long)up2[x] : N;
        long long wpred = 0; if (uses_wp) wpred = WPPredict(wp, ws, wo, x, y, w, N, W, NE, NW, NN);
        DTreeNode nd = n;
        if (!single_leaf) {
          int idx = root;
          while (nd.x >= 0) {
            long long v;
            switch (nd.x) {
              case 0: v = chan; break; case 1: v = stream_id; break; case 2: v = y; break; case 3: v = x; break;
              case 4: v = llabs(N); break; case 5: v = llabs(W); break; case 6: v = N; break; case 7: v = W; break;
              case 8: v = W - prev_grad; break; case 9: v = W + N - NW; break; case 10: v = W - NW; break; case 11: v = NW - N; break;
              case 12: v = N - NE; break; case 13: v = N - NN; break; case 14: v = W - WW; break; case 15: v = wo.max_err; break;
              default: v = 0; rd.err = kErrRefProps; break;
            }
            idx = int32_t(v) > nd.y ? nd.z : nd.w; nd = tree[idx];
          }
        }
        uint32_t tok = rd.Read(cv, uint32_t(nd.y) >> 4);
        long long pred;
        switch (nd.y & 15) {
          case 0: pred = 0; break; case 1: pred = W; break; case 2: pred = N; break; case 3: pred = (W + N) / 2; break;
          case 4: { long long p = W + N - NW; pred = llabs(p - W) < llabs(p - N) ? W : N; break; }
          case 5: { long long lo = min(W, N), hi = max(W, N); pred = max(lo, min(hi, W + N - NW)); break; }
          case 6: pred = wpred; break; case 7: pred = NE; break; case 8: pred = NW; break; case 9: pred = WW; break;
          case 10: pred = (W + NW) / 2; break; case 11: pred = (N + NW) / 2; break; case 12: pred = (N + NE) / 2; break;
          default: pred = (6 * N - 2 * NN + 7 * W + WW + NEE + 3 * NE + 8) / 16; break;
        }
        int32_t val = int32_t((long long)UnpackSignedDev(tok) * (long long)uint32_t(nd.w) + nd.z + pred);
        cur[x] = val;
        if (uses_wp) WPUpdate(ws, wo, val, x, y);
        prev_grad = int32_t(W + N - NW);
        // roll the window to x+1
        { long long oldW = W; W = val; WW = x >= 1 ? oldW : W; }
        if (y) { NW = N; N = NE; NE = NEE; NEE = (x + 3 < w) ? (long long)up[x + 3] : NE; }
        else { N = W; NW = W; NE = W; NEE = W; }
      }
    }
  }
};
#endif

}  // namespace jxlgpu
