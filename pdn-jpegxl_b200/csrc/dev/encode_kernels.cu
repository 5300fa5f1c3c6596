// pdn-jpegxl_b200 engine — encode kernels (sm_100a) behind SaveImage, low-effort pipeline:
//   k_enc_scan        BGRA scan: isGray / hasTransparency            (N/Encoder/JxlEncoder.cpp:33-77), 4 B/px read
//   k_enc_to_xyb      BGRA -> linear -> XYB planes (+ alpha plane)    (BgraTo* of N/Encoder/PixelFormatConversion.cpp:16-121 fused in), 4 + 12 B/px
//   k_enc_to_planes   BGRA -> integer planes + YCgCo RCT (lossless)
//   k_enc_sharpen     Van Cittert step of the inverse gaborish
//   k_enc_dct8        forward DCT8 + quantise + LF extraction          (SURVEY.md A.9/A.11), 12 + 6 B/px
//   k_enc_lf_quant    LF quantisation with chroma-from-luma
//   k_enc_mod_tokens  Modular tokenisation (gradient predictor, fixed MA tree), one thread per sample
//   k_enc_ac_tokens   AC tokenisation with the A.8 context model, one thread per 256x256 group
//   k_enc_histogram   token histograms; k_enc_ans: ANS stream writer, one thread per section stream
// Replaces the libjxl work behind JxlEncoderAddImageFrame / JxlEncoderFlushInput (N/Encoder/JxlEncoder.cpp:128,367).
#include "enc_frame.cuh"
#include "kernels.h"

namespace jxlgpu {

__device__ __constant__ float kSrgbLut[256];
void UploadSrgbLut(const float* lut) { cudaMemcpyToSymbol(kSrgbLut, lut, 256 * sizeof(float)); }

__global__ void k_enc_scan(const uint8_t* __restrict__ bgra, uint32_t w, uint32_t h, uint32_t stride, uint32_t* flags) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y; uint32_t f = 0;
  if (x < w && y < h) { uchar4 p = *reinterpret_cast<const uchar4*>(bgra + size_t(y) * stride + size_t(x) * 4); if (!(p.z == p.y && p.y == p.x)) f |= 1; if (p.w < 255) f |= 2; }
  f = __reduce_or_sync(0xffffffffu, f); if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

__global__ void k_enc_to_xyb(const DEncFrame* ep, const uint8_t* __restrict__ bgra) {
  // yy counts rows of the extended plane (halo rows first); `bgra` points at the first row of the band itself, halo rows lie before it
  const DEncFrame& e = *ep; int xx = blockIdx.x * blockDim.x + threadIdx.x, ye = blockIdx.y * blockDim.y + threadIdx.y; if (xx >= int(e.xpad) || ye >= int(e.ext_rows)) return;
  const int yy = ye - int(e.ext_top);
  int x = min(xx, int(e.xsize) - 1), y = max(e.src_row_min, min(yy, e.src_row_max)); uchar4 p = *reinterpret_cast<const uchar4*>(bgra + ptrdiff_t(y) * ptrdiff_t(e.stride) + size_t(x) * 4);
  float r, g, b;
  if (e.has_src_profile) {   // matrix/TRC ICC source: tone curves from the profile, then its colorants -> linear sRGB (what libjxl's CMS step does)
    const float lr = e.src_lut[p.z], lg = e.src_lut[256 + p.y], lb = e.src_lut[512 + p.x]; const float* m = e.src_matrix;
    r = m[0] * lr + m[1] * lg + m[2] * lb; g = m[3] * lr + m[4] * lg + m[5] * lb; b = m[6] * lr + m[7] * lg + m[8] * lb;
  } else { r = kSrgbLut[e.gray ? p.x : p.z]; g = kSrgbLut[e.gray ? p.x : p.y]; b = kSrgbLut[p.x]; }   // gray takes the B channel (N/Encoder/PixelFormatConversion.cpp:34)
  const float bias = 0.0037930732552754493f, cb = cbrtf(bias);
  float m0 = 0.30f * r + 0.622f * g + 0.078f * b + bias, m1 = 0.23f * r + 0.692f * g + 0.078f * b + bias, m2 = 0.24342268924547819f * r + 0.20476744424496821f * g + 0.55180986650955360f * b + bias;
  float g0 = cbrtf(fmaxf(m0, 0.f)) - cb, g1 = cbrtf(fmaxf(m1, 0.f)) - cb, g2 = cbrtf(fmaxf(m2, 0.f)) - cb;
  size_t plane = size_t(e.xpad) * e.ext_rows, at = size_t(ye) * e.xpad + xx; e.xyb[at] = 0.5f * (g0 - g1); e.xyb[plane + at] = 0.5f * (g0 + g1); e.xyb[2 * plane + at] = g2;
  if (e.alpha && xx < int(e.xsize) && yy >= 0 && yy < int(e.ysize)) e.planes[size_t(e.alpha_plane) * e.xsize * e.ysize + size_t(yy) * e.xsize + xx] = p.w;
}

__global__ void k_enc_to_planes(const DEncFrame* ep, const uint8_t* __restrict__ bgra) {
  const DEncFrame& e = *ep; int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y; if (x >= int(e.xsize) || y >= int(e.ysize)) return;
  uchar4 p = *reinterpret_cast<const uchar4*>(bgra + size_t(y) * e.stride + size_t(x) * 4); size_t n = size_t(e.xsize) * e.ysize, at = size_t(y) * e.xsize + x;
  if (e.gray) e.planes[at] = p.x;
  else { int32_t R = p.z, G = p.y, B = p.x; int32_t Co = R - B; int32_t t = B + (Co >> 1); int32_t Cg = G - t; int32_t Y = t + (Cg >> 1); e.planes[at] = Y; e.planes[n + at] = Co; e.planes[2 * n + at] = Cg; }
  if (e.alpha) e.planes[size_t(e.alpha_plane) * n + at] = p.w;
}

__global__ void k_enc_sharpen(float* __restrict__ cur, const float* __restrict__ orig, const float* __restrict__ blur, size_t n) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; if (i < n) cur[i] += orig[i] - blur[i];
}

// One thread per 8x8 block; channels in order Y, X, B because X/B quantise against the dequantised Y (A.8 chroma from luma).
__global__ void __launch_bounds__(64) k_enc_dct8(const DEncFrame* ep) {
  const DEncFrame& e = *ep; int cell = blockIdx.x * blockDim.x + threadIdx.x; if (cell >= int(e.xb * e.yb)) return;
  int cy = cell / int(e.xb), cx = cell % int(e.xb); int g = (cy >> 5) * int(e.xgroups) + (cx >> 5); int by = cy & 31, bx = cx & 31;
  const float* cos8 = e.tables->cosines + CosOff(3); const float* dq = e.dequant8; size_t plane = size_t(e.xpad) * e.ext_rows, lfplane = size_t(e.xb) * e.yb;
  float scale = e.inv_gs / float(e.hf_mul); float ydq[64]; uint32_t nzc[3];
#pragma unroll 1
  for (int ci = 0; ci < 3; ci++) {
    const int c = ci == 0 ? 1 : ci == 1 ? 0 : 2; const float* px = e.xyb + c * plane + (size_t(e.ext_top) + size_t(cy) * 8) * e.xpad + size_t(cx) * 8; float t[64], S[64];
    // rows: t[y][hf] = (1/8) sum_x px[y][x] cos[hf][x]
    for (int y = 0; y < 8; y++) { float v[8]; for (int x = 0; x < 8; x++) v[x] = px[size_t(y) * e.xpad + x]; for (int k = 0; k < 8; k++) { float a = 0; for (int x = 0; x < 8; x++) a += v[x] * cos8[k * 8 + x]; t[y * 8 + k] = a * 0.125f; } }
    // columns: F[vf][hf]; storage (square block) S[hf][vf]
    for (int hf = 0; hf < 8; hf++) for (int vf = 0; vf < 8; vf++) { float a = 0; for (int y = 0; y < 8; y++) a += t[y * 8 + hf] * cos8[vf * 8 + y]; S[hf * 8 + vf] = a * 0.125f; }
    e.lf[c * lfplane + cell] = S[0];
    int16_t* out = e.coeffs + (size_t(g) * 3 + c) * 65536 + (size_t(by) * 32 + bx) * 64; uint32_t nz = 0;
    const float mulc = c == 1 ? scale : c == 0 ? scale * e.xm : scale * e.bm, kc = c == 0 ? e.kx : e.kb;
    for (int p = 0; p < 64; p++) {
      int q = 0;
      if (p) { float step = dq[c * 64 + p] * mulc; float v = S[p]; if (c != 1) v -= kc * ydq[p]; float qf = v / step; if (fabsf(qf) >= 0.56f) q = __float2int_rn(qf); q = max(-32768, min(32767, q));
        if (c == 1) { float a = q == 0 ? 0.f : q == 1 ? e.quant_bias[1] : q == -1 ? -e.quant_bias[1] : float(q) - e.quant_bias[3] / float(q); ydq[p] = a * step; } nz += q != 0; }
      out[p] = int16_t(q);
    }
    nzc[c] = nz;
  }
  e.nz[cell] = uint8_t(nzc[0]); e.nz[lfplane + cell] = uint8_t(nzc[1]); e.nz[2 * lfplane + cell] = uint8_t(nzc[2]);
}

__global__ void k_enc_lf_quant(const DEncFrame* ep) {
  const DEncFrame& e = *ep; size_t n = size_t(e.xb) * e.yb, i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; if (i >= n) return;
  float fy = e.lf_fac[1]; int32_t qy = __float2int_rn(e.lf[n + i] / fy); float ydq = float(qy) * fy;
  int32_t qx = __float2int_rn((e.lf[i] - e.cfl_x_lf * ydq) / e.lf_fac[0]), qb = __float2int_rn((e.lf[2 * n + i] - e.cfl_b_lf * ydq) / e.lf_fac[2]);
  e.lfq[i] = qy; e.lfq[n + i] = qx; e.lfq[2 * n + i] = qb;   // stream channel order Y, X, B
}

__device__ __forceinline__ uint32_t PackSignedDev(int32_t v) { return (uint32_t(v) << 1) ^ uint32_t(v >> 31); }
// token word: ctx[0,16) | symbol[16,24) | nbits[24,30); second word: extra bits
__device__ __forceinline__ void MakeToken(uint32_t ctx, uint32_t value, uint32_t split_exp, uint32_t msb, uint2* out) {
  uint32_t split = 1u << split_exp, tok, nb = 0, bits = 0;
  if (value < split) tok = value; else { uint32_t n = 31 - __clz(value), m = value - (1u << n); tok = split + ((n - split_exp) << msb) + (m >> (n - msb)); nb = n - msb; bits = m & ((1u << nb) - 1); }
  *out = make_uint2(ctx | (tok << 16) | (nb << 24), bits);
}

// Modular tokenisation: stream s covers rect (x0,y0,w,h) of `nch` planes of size (pw x ph); tokens in channel-major raster order.
__global__ void k_enc_mod_tokens(const DEncFrame* ep, const DEncModStream* streams, const int32_t* __restrict__ planes, uint32_t pw, uint32_t ph, uint32_t nch, const uint16_t* __restrict__ leaf_lut /*[kind][8 ch][11]*/) {
  const DEncFrame& e = *ep; const DEncModStream s = streams[blockIdx.y]; uint32_t per = s.w * s.h, total = per * nch, i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= total) return;
  uint32_t c = i / per, r = i % per; int y = int(r / s.w), x = int(r % s.w); const int32_t* p = planes + size_t(c) * pw * ph + size_t(s.y0) * pw + s.x0; const int W = int(pw);
  auto at = [&](int yy, int xx) -> int32_t { return p[size_t(yy) * W + xx]; };
  int32_t Wv = x ? at(y, x - 1) : (y ? at(y - 1, x) : 0), N = y ? at(y - 1, x) : Wv, NW = (x && y) ? at(y - 1, x - 1) : Wv;
  int32_t lo = min(Wv, N), hi = max(Wv, N), pred = max(lo, min(hi, Wv + N - NW));
  int32_t prev_grad = 0; if (x) { int xp = x - 1; int32_t Wp = xp ? at(y, xp - 1) : (y ? at(y - 1, xp) : 0), Np = y ? at(y - 1, xp) : Wp, NWp = (xp && y) ? at(y - 1, xp - 1) : Wp; prev_grad = Wp + Np - NWp; }
  int32_t prop8 = Wv - prev_grad; const int32_t thr[10] = {-64, -24, -8, -3, -1, 0, 2, 7, 23, 63}; int bucket = 0; for (int k = 0; k < 10; k++) bucket += prop8 > thr[k];
  uint32_t ctx = leaf_lut[(s.kind * 8 + c) * 11 + bucket];
  MakeToken(ctx, PackSignedDev(at(y, x) - pred), 4, 1, e.tokens + s.token_off + i);
}

__device__ __constant__ uint8_t kFreqCtxE[64] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22,
  23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 26, 27, 27, 27, 27, 28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30};
__device__ __constant__ uint8_t kNumNzCtxE[64] = {0, 0, 31, 62, 62, 93, 93, 93, 93, 123, 123, 123, 123, 152, 152, 152, 152, 152, 152, 152, 152, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180,
  206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206};

// AC tokenisation for DCT8-only frames (default block-context map: 15 contexts, order id 0).
__global__ void __launch_bounds__(32) k_enc_ac_tokens(const DEncFrame* ep) {
  const DEncFrame& e = *ep; int g = blockIdx.x * blockDim.x + threadIdx.x; if (g >= int(e.num_groups)) return;
  int gx = g % int(e.xgroups), gy = g / int(e.xgroups), cx0 = gx * 32, cy0 = gy * 32, w = min(32, int(e.xb) - cx0), h = min(32, int(e.yb) - cy0); size_t lfplane = size_t(e.xb) * e.yb;
  uint2* out = e.tokens + e.ac_token_off + size_t(g) * kMaxAcTokensPerGroup; uint32_t n = 0; const uint32_t nbctx = 15; const uint8_t bctx_of[3] = {7, 0, 7};   // default map, order 0: Y->0, X->7, B->7 (A.8)
  const uint16_t* order = e.order8;
  for (int by = 0; by < h; by++) for (int bx = 0; bx < w; bx++) {
    size_t cell = size_t(cy0 + by) * e.xb + cx0 + bx;
    for (int ci = 0; ci < 3; ci++) {
      int c = ci == 0 ? 1 : ci == 1 ? 0 : 2; const uint8_t* nzp = e.nz + c * lfplane; uint32_t nz = nzp[cell];
      uint32_t pred; if (bx == 0) pred = by == 0 ? 32 : nzp[cell - e.xb]; else if (by == 0) pred = nzp[cell - 1]; else pred = (uint32_t(nzp[cell - e.xb]) + nzp[cell - 1] + 1) >> 1;
      uint32_t nzb = pred < 8 ? pred : (pred >= 64 ? 36 : 4 + pred / 2); uint32_t bc = bctx_of[c];
      MakeToken(nzb * nbctx + bc, nz, 4, 2, out + n++);
      const int16_t* co = e.coeffs + (size_t(g) * 3 + c) * 65536 + (size_t(by) * 32 + bx) * 64; uint32_t histo = nbctx * 37 + 458 * bc, prev = nz > 4 ? 0 : 1;
      for (uint32_t k = 1; k < 64 && nz != 0; k++) { int32_t v = co[order[k]]; uint32_t zctx = (uint32_t(kNumNzCtxE[nz]) + kFreqCtxE[k]) * 2 + prev; MakeToken(histo + zctx, PackSignedDev(v), 4, 2, out + n++); prev = v != 0; nz -= prev; }
    }
  }
  e.ac_token_count[g] = n;
}

__global__ void k_enc_histogram(const uint2* __restrict__ tokens, const DEncStream* streams, uint32_t* hist) {
  const DEncStream s = streams[blockIdx.y]; uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= s.count) return;
  uint32_t t = tokens[s.token_off + i].x; atomicAdd(hist + size_t(t & 0xffff) * kEncAlphabet + ((t >> 16) & 0xff), 1u);
}

// ANS writer (A.6 "ANS symbol read" mirrored): reverse pass computes the state chain, forward pass packs the bits.
__global__ void __launch_bounds__(32) k_enc_ans(const DEncFrame* ep, const DEncStream* streams, uint32_t nstreams, const DEncCode* code) {
  const DEncFrame& e = *ep; uint32_t si = blockIdx.x * blockDim.x + threadIdx.x; if (si >= nstreams) return; const DEncStream s = streams[si];
  uint2* tk = e.tokens + s.token_off; uint32_t state = 0x130000u;
  const uint8_t* ctx_map = code->ctx_map; const uint16_t* freq = code->freq; const uint16_t* start = code->start; const uint16_t* rev = code->rev;
  for (uint32_t i = s.count; i-- > 0;) {
    uint32_t t = tk[i].x, cl = ctx_map[t & 0xffff], sym = (t >> 16) & 0xff; uint32_t f = freq[cl * kEncAlphabet + sym], flush = 0;
    if ((state >> 20) >= f) { flush = 0x10000u | (state & 0xffff); state >>= 16; }
    state = ((state / f) << 12) + rev[cl * 4096 + start[cl * kEncAlphabet + sym] + (state % f)];
    tk[i].x = (t & 0x3f000000u) | flush;   // keep nbits, replace ctx/symbol by the flush word
  }
  uint8_t* out = e.stream_bytes + s.byte_off; uint64_t acc = state; int nb = 32; size_t pos = 0;
  auto put = [&](uint32_t v, int n) { acc |= uint64_t(v) << nb; nb += n; while (nb >= 8) { out[pos++] = uint8_t(acc); acc >>= 8; nb -= 8; } };
  put(0, 0);
  for (uint32_t i = 0; i < s.count; i++) { uint2 t = tk[i]; if (t.x & 0x10000u) put(t.x & 0xffff, 16); int n = int((t.x >> 24) & 0x3f); if (n) put(t.y, n); }
  uint64_t bits = uint64_t(pos) * 8 + nb; if (nb) out[pos++] = uint8_t(acc);
  e.stream_bits[si] = bits;
}

void EncLaunchScan(const uint8_t* bgra, uint32_t w, uint32_t h, uint32_t stride, uint32_t* flags, cudaStream_t st) { dim3 grid((w + 255) / 256, h); k_enc_scan<<<grid, 256, 0, st>>>(bgra, w, h, stride, flags); CountLaunch(); }
void EncLaunchToXyb(const DEncFrame* d, const DEncFrame& h, const uint8_t* bgra, cudaStream_t st) { dim3 blk(32, 8), grid((h.xpad + 31) / 32, (h.ext_rows + 7) / 8); k_enc_to_xyb<<<grid, blk, 0, st>>>(d, bgra); CountLaunch(); }
void EncLaunchToPlanes(const DEncFrame* d, const DEncFrame& h, const uint8_t* bgra, cudaStream_t st) { dim3 blk(32, 8), grid((h.xsize + 31) / 32, (h.ysize + 7) / 8); k_enc_to_planes<<<grid, blk, 0, st>>>(d, bgra); CountLaunch(); }
void EncLaunchSharpen(float* cur, const float* orig, const float* blur, size_t n, cudaStream_t st) { k_enc_sharpen<<<unsigned((n + 255) / 256), 256, 0, st>>>(cur, orig, blur, n); CountLaunch(); }
void EncLaunchDct8(const DEncFrame* d, const DEncFrame& h, cudaStream_t st) { uint32_t cells = h.xb * h.yb; k_enc_dct8<<<(cells + 63) / 64, 64, 0, st>>>(d); k_enc_lf_quant<<<(cells + 255) / 256, 256, 0, st>>>(d); CountLaunch(2); }
void EncLaunchModTokens(const DEncFrame* d, const DEncModStream* streams, uint32_t nstreams, uint32_t max_tokens, const int32_t* planes, uint32_t pw, uint32_t ph, uint32_t nch, const uint16_t* leaf_lut, cudaStream_t st) {
  if (!nstreams || !max_tokens) return; dim3 grid((max_tokens + 255) / 256, nstreams); k_enc_mod_tokens<<<grid, 256, 0, st>>>(d, streams, planes, pw, ph, nch, leaf_lut); CountLaunch();
}
void EncLaunchAcTokens(const DEncFrame* d, const DEncFrame& h, cudaStream_t st) { k_enc_ac_tokens<<<(h.num_groups + 31) / 32, 32, 0, st>>>(d); CountLaunch(); }
void EncLaunchHistogram(const uint2* tokens, const DEncStream* streams, uint32_t nstreams, uint32_t max_count, uint32_t* hist, cudaStream_t st) {
  if (!nstreams || !max_count) return; dim3 grid((max_count + 255) / 256, nstreams); k_enc_histogram<<<grid, 256, 0, st>>>(tokens, streams, hist); CountLaunch();
}
void EncLaunchAns(const DEncFrame* d, const DEncStream* streams, uint32_t nstreams, const DEncCode* code, cudaStream_t st) { if (!nstreams) return; k_enc_ans<<<(nstreams + 31) / 32, 32, 0, st>>>(d, streams, nstreams, code); CountLaunch(); }

}  // namespace jxlgpu
