// pdn-jpegxl_b200 engine — host-side bitstream front-end (AC strategy tables, natural coefficient order, dequantisation matrices, AC contexts).
// Part of the product: parses what must be parsed serially on the CPU and hands flat tables
// to the sm_100a kernels. Replaces libjxl work reached from N/Decoder/JxlDecoder.cpp:252,454
// and N/Encoder/JxlEncoder.cpp:128,367 of the reference. Per ISO/IEC 18181-1 as digested in
// SURVEY.md Appendix A.8/A.9. Constants tagged [M]/[L] there are unverified against libjxl.
#pragma once
#include "bits.h"
#include <map>
#include <mutex>

namespace jxlgpu {

static const int kNumStrategies = 27, kNumOrders = 13, kNumQuantTables = 17;
enum { kDCT8 = 0, kIdentity = 1, kDCT2x2 = 2, kDCT4x4 = 3, kDCT16 = 4, kDCT32 = 5, kDCT16x8 = 6, kDCT8x16 = 7, kDCT32x8 = 8, kDCT8x32 = 9, kDCT32x16 = 10, kDCT16x32 = 11,
       kDCT4x8 = 12, kDCT8x4 = 13, kAFV0 = 14, kAFV3 = 17, kDCT64 = 18, kDCT64x32 = 19, kDCT32x64 = 20, kDCT128 = 21, kDCT128x64 = 22, kDCT64x128 = 23, kDCT256 = 24, kDCT256x128 = 25, kDCT128x256 = 26 };
// covered 8x8 cells: rows (cy) and cols (cx) per strategy
static const uint8_t kCoveredY[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16};
static const uint8_t kCoveredX[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32};
static const uint8_t kStrategyOrder[27] = {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12};
static const uint8_t kQuantTableOf[27] = {0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10, 11, 12, 12, 13, 14, 14, 15, 16, 16};
// quant table storage dims in cells (rows <= cols)
static const uint8_t kTableRows[17] = {1, 1, 1, 1, 2, 4, 1, 1, 2, 1, 1, 8, 4, 16, 8, 32, 16};
static const uint8_t kTableCols[17] = {1, 1, 1, 1, 2, 4, 2, 4, 4, 1, 1, 8, 8, 16, 16, 32, 32};
// representative strategy for each coefficient-order id
static const uint8_t kOrderStrategy[13] = {kDCT8, kIdentity, kDCT16, kDCT32, kDCT8x16, kDCT8x32, kDCT16x32, kDCT64, kDCT32x64, kDCT128, kDCT64x128, kDCT256, kDCT128x256};
inline bool IsPlainDCT(int s) { return s == kDCT8 || (s >= kDCT16 && s <= kDCT16x32) || s >= kDCT64; }

// Natural order for a storage layout of `rs` x `cs` cells (rs <= cs): order[k] = storage index (A.9 [M]).
inline std::vector<uint32_t> NaturalOrder(int rs, int cs) {
  int xs = std::max(rs, cs), ys = std::min(rs, cs); int ratio_log2 = CeilLog2(uint64_t(xs / ys)); uint32_t mask = (1u << ratio_log2) - 1;
  int W = xs * 8; std::vector<uint32_t> order(size_t(xs) * ys * 64); size_t cur = size_t(xs) * ys;
  for (int i = 0; i < W; i++) for (int j = 0; j <= i; j++) {
    int x = j, y = i - j; if (i & 1) std::swap(x, y); if (y & mask) continue; y >>= ratio_log2;
    size_t val; if (x < xs && y < ys) val = size_t(y) * xs + x; else val = cur++;
    order[val] = uint32_t(y * W + x);
  }
  for (int ip = W - 1; ip > 0; ip--) { int i = ip - 1;
    for (int j = 0; j <= i; j++) { int x = W - 1 - (i - j), y = W - 1 - j; if (i & 1) std::swap(x, y); if (y & mask) continue; y >>= ratio_log2; order[cur++] = uint32_t(y * W + x); } }
  return order;
}

// ------------------------------------------------------------------ dequant matrices
struct DctParams { int num_bands; float bands[3][17]; };
struct QuantEncoding { int mode = 0; DctParams dct; float idw[3][3]; float dct2w[3][6]; float dct4mul[3][2]; float dct4x8mul[3]; DctParams dct4x8, dct4; float afvw[3][9]; };
enum { kQModeLibrary = 0, kQModeId = 1, kQModeDCT2 = 2, kQModeDCT4 = 3, kQModeDCT4x8 = 4, kQModeAFV = 5, kQModeDCT = 6, kQModeRAW = 7 };

inline DctParams MakeParams(int n, std::initializer_list<double> x, std::initializer_list<double> y, std::initializer_list<double> b) {
  DctParams p; p.num_bands = n; int i = 0; for (double v : x) p.bands[0][i++] = float(v); i = 0; for (double v : y) p.bands[1][i++] = float(v); i = 0; for (double v : b) p.bands[2][i++] = float(v); return p;
}
// Library (default) encodings, table ids 0..16. DCT8 values [H/M]; the rest [M/L].
inline QuantEncoding LibraryEncoding(int t) {
  QuantEncoding e; e.mode = kQModeDCT;
  auto big = [&](double m) { return MakeParams(8, {m * 26629.073922049845, -1.025, -0.78, -0.65012, -0.19041574084286472, -0.20819395464, -0.421064, -0.32733845535848671},
                                               {m * 9311.3238710010046, -0.3041958212306401, -0.3633036457487539, -0.35660379990111464, -0.3443074455424403, -0.33699592683512467, -0.30180866526242109, -0.27321683125358037},
                                               {m * 4992.2486445538634, -1.2, -1.2, -0.8, -0.7, -0.7, -0.4, -0.5}); };
  auto bigr = [&](double m) { return MakeParams(8, {m * 23629.073922049845, -1.025, -0.78, -0.65012, -0.19041574084286472, -0.20819395464, -0.421064, -0.32733845535848671},
                                                {m * 8611.3238710010046, -0.3041958212306401, -0.3633036457487539, -0.35660379990111464, -0.3443074455424403, -0.33699592683512467, -0.30180866526242109, -0.27321683125358037},
                                                {m * 4492.2486445538634, -1.2, -1.2, -0.8, -0.7, -0.7, -0.4, -0.5}); };
  switch (t) {
    case 0: e.dct = MakeParams(6, {3150.0, 0.0, -0.4, -0.4, -0.4, -2.0}, {560.0, 0.0, -0.3, -0.3, -0.3, -0.3}, {512.0, -2.0, -1.0, 0.0, -1.0, -2.0}); break;
    case 1: { e.mode = kQModeId; float w[3][3] = {{280.0f, 3160.0f, 3160.0f}, {60.0f, 864.0f, 864.0f}, {18.0f, 200.0f, 200.0f}}; memcpy(e.idw, w, sizeof(w)); break; }
    case 2: { e.mode = kQModeDCT2; float w[3][6] = {{3840.0f, 2560.0f, 1280.0f, 640.0f, 480.0f, 300.0f}, {960.0f, 640.0f, 320.0f, 180.0f, 140.0f, 120.0f}, {640.0f, 320.0f, 128.0f, 64.0f, 32.0f, 16.0f}}; memcpy(e.dct2w, w, sizeof(w)); break; }
    case 3: { e.mode = kQModeDCT4; e.dct4 = MakeParams(4, {2200.0, 0.0, 0.0, 0.0}, {392.0, 0.0, 0.0, 0.0}, {112.0, -0.25, -0.25, -0.5}); for (auto& m : e.dct4mul) m[0] = m[1] = 1.0f; break; }
    case 4: e.dct = MakeParams(7, {8996.8725711814115328, -1.3000777393353804, -0.49424529824571225, -0.439093774457103443, -0.6350101832695744, -0.90177264050827612, -1.6162099239887414},
                               {3191.48366296844234752, -0.67424582104194355, -0.80745813428471001, -0.44925837484843441, -0.35865440981033403, -0.31322389111877305, -0.37615025315725483},
                               {1157.50408145487200256, -2.0531423165804414, -1.4, -0.50687130033378396, -0.42708730624733904, -1.4856834539296244, -4.9209142884401604}); break;
    case 5: e.dct = MakeParams(8, {15718.40830982518931456, -1.025, -0.98, -0.9012, -0.4, -0.48819395464, -0.421064, -0.27},
                               {7305.7636810695983104, -0.8041958212306401, -0.7633036457487539, -0.55660379990111464, -0.49785304658857626, -0.43699592683512467, -0.40180866526242109, -0.27321683125358037},
                               {3803.53173721215041536, -3.060733579805728, -2.0413270132490346, -2.0235650159727417, -0.5495389509954993, -0.4, -0.4, -0.3}); break;
    case 6: e.dct = MakeParams(7, {7240.7734393502, -0.7, -0.7, -0.2, -0.2, -0.2, -0.5}, {1448.15468787004, -0.5, -0.5, -0.5, -0.2, -0.2, -0.2}, {506.854140754517, -1.4, -0.2, -0.5, -0.5, -1.5, -3.6}); break;
    case 7: e.dct = MakeParams(8, {16283.2494710648897, -1.7812845336559429, -1.6309059012653515, -1.0382179034313539, -0.85, -0.7, -0.9, -1.2360638576849587},
                               {5089.15750884921511936, -0.320049391452786891, -0.35362849922161446, -0.30340000000000003, -0.61, -0.5, -0.5, -0.6},
                               {3397.77603275308720128, -0.321327362693153371, -0.34507619223117997, -0.70340000000000003, -0.9, -1.0, -1.0, -1.1754605576265209}); break;
    case 8: e.dct = MakeParams(8, {13844.97076442300573, -0.97113799999999995, -0.658, -0.42026, -0.22712, -0.2206, -0.226, -0.6},
                               {4798.964084220744293, -0.61125308982767057, -0.83770786552491361, -0.79014862079498627, -0.2692727459704829, -0.38272769465388551, -0.22924222653091453, -0.20719098826199578},
                               {1807.236946760964614, -1.2, -1.2, -0.7, -0.7, -0.7, -0.4, -0.5}); break;
    case 9: e.mode = kQModeDCT4x8; e.dct4x8 = MakeParams(4, {2198.050556016380522, -0.96269623020744692, -0.76194253026666783, -0.6551140670773547},
                               {764.3655248643528689, -0.92630200888366945, -0.9675229603596517, -0.27845290869168118}, {527.107573587542228, -1.4594385811273854, -1.450082094097871593, -1.5843722511996204});
            for (auto& m : e.dct4x8mul) m = 1.0f; break;
    case 10: e.mode = kQModeAFV; break;   // AFV basis not recalled: decoding AFV blocks is rejected
    case 11: e.dct = big(0.9); break; case 12: e.dct = bigr(0.65); break; case 13: e.dct = big(1.8); break; case 14: e.dct = bigr(1.3); break;
    case 15: e.dct = big(3.6); break; case 16: e.dct = bigr(2.6); break;
  }
  return e;
}
inline void QuantWeights(int rows, int cols, const DctParams& p, int c, float* out) {
  float bands[17]; bands[0] = p.bands[c][0]; JXLG_CHECK(bands[0] >= 1e-8f, "distance band");
  for (int i = 1; i < p.num_bands; i++) { float v = p.bands[c][i]; bands[i] = bands[i - 1] * (v > 0 ? 1.0f + v : 1.0f / (1.0f - v)); JXLG_CHECK(bands[i] >= 1e-8f, "distance band"); }
  float scale = float(p.num_bands - 1) / (float(std::sqrt(2.0)) + 1e-6f), rcpcol = scale / float(cols - 1), rcprow = scale / float(rows - 1);
  for (int y = 0; y < rows; y++) { float dy = y * rcprow; for (int x = 0; x < cols; x++) { float dx = x * rcpcol; float d = std::sqrt(dx * dx + dy * dy);
      float w; if (p.num_bands == 1) w = bands[0]; else { int idx = int(d); if (idx > p.num_bands - 2) idx = p.num_bands - 2; float frac = d - idx; w = bands[idx] * std::pow(bands[idx + 1] / bands[idx], frac); }
      out[y * cols + x] = w; } }
}
// Fills dequant multipliers (1/weight) for table `t`, 3 channels, storage layout rows x cols coefficients.
inline std::vector<float> ComputeDequantTable(int t, const QuantEncoding& e) {
  int rows = kTableRows[t] * 8, cols = kTableCols[t] * 8; size_t n = size_t(rows) * cols; std::vector<float> w(3 * n, 0.f);
  for (int c = 0; c < 3; c++) {
    float* o = w.data() + c * n;
    switch (e.mode) {
      case kQModeDCT: QuantWeights(rows, cols, e.dct, c, o); break;
      case kQModeId: for (size_t i = 0; i < 64; i++) o[i] = e.idw[c][0]; o[1] = e.idw[c][1]; o[8] = e.idw[c][1]; o[9] = e.idw[c][2]; break;
      case kQModeDCT2: { const float* d = e.dct2w[c]; o[0] = 1.0f; o[1] = o[8] = d[0]; o[9] = d[1];
        for (int y = 0; y < 2; y++) for (int x = 0; x < 2; x++) { o[y * 8 + x + 2] = d[2]; o[(y + 2) * 8 + x] = d[2]; o[(y + 2) * 8 + x + 2] = d[3]; }
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) { o[y * 8 + x + 4] = d[4]; o[(y + 4) * 8 + x] = d[4]; o[(y + 4) * 8 + x + 4] = d[5]; } break; }
      case kQModeDCT4: { float w4[16]; QuantWeights(4, 4, e.dct4, c, w4); for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) o[y * 8 + x] = w4[(y / 2) * 4 + x / 2];
        o[1] /= e.dct4mul[c][0]; o[8] /= e.dct4mul[c][0]; o[9] /= e.dct4mul[c][1]; break; }
      case kQModeDCT4x8: { float w48[32]; QuantWeights(4, 8, e.dct4x8, c, w48); for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) o[y * 8 + x] = w48[(y / 2) * 8 + x]; o[8] /= e.dct4x8mul[c]; break; }
      default: for (size_t i = 0; i < n; i++) o[i] = 1.0f; break;   // AFV / RAW: unsupported, never used (decoder rejects)
    }
  }
  for (auto& v : w) { JXLG_CHECK(v > 1e-8f && v < 1e8f, "quant weight out of range"); v = 1.0f / v; }
  return w;
}

// resample scale for an N-point LF DCT reinterpreted inside an 8N-point DCT (A.9 closed form)
inline double ResampleScale(int N, int u) { double p = 1; for (int k = 0; k < 3; k++) p *= std::cos(u * M_PI * double(1 << k) / (16.0 * N)); return 1.0 / p; }

static const uint8_t kCoeffFreqContext[64] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22,
  23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 26, 27, 27, 27, 27, 28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30};
static const uint8_t kCoeffNumNonzeroContext[64] = {0, 0, 31, 62, 62, 93, 93, 93, 93, 123, 123, 123, 123, 152, 152, 152, 152, 152, 152, 152, 152, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180,
  206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206};
static const int kNonZeroBuckets = 37, kZeroDensityContextCount = 458;
inline uint32_t ZeroDensityContext(uint32_t nz_left, uint32_t k, uint32_t covered, uint32_t log2_covered, uint32_t prev) {
  nz_left = (nz_left + covered - 1) >> log2_covered; k >>= log2_covered; return (kCoeffNumNonzeroContext[nz_left] + kCoeffFreqContext[k]) * 2 + prev;
}
inline uint32_t NonZeroCtxBucket(uint32_t n) { if (n > 64) n = 64; return n < 8 ? n : (n >= 64 ? 36 : 4 + n / 2); }

struct BlockCtxMap {
  std::vector<int32_t> lf_thr[3]; std::vector<uint32_t> qf_thr; std::vector<uint8_t> map; uint32_t num_ctxs = 15, num_lf_ctxs = 1;
  BlockCtxMap() { static const uint8_t d[39] = {0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14}; map.assign(d, d + 39); }
  uint32_t Context(uint32_t lf_idx, uint32_t qf, uint32_t ord, uint32_t c) const {
    uint32_t qf_idx = 0; for (uint32_t t : qf_thr) if (qf > t) qf_idx++;
    uint32_t idx = c < 2 ? (c ^ 1) : 2; idx = idx * kNumOrders + ord; idx = idx * (uint32_t(qf_thr.size()) + 1) + qf_idx; idx = idx * num_lf_ctxs + lf_idx; return map[idx];
  }
};

}  // namespace jxlgpu
