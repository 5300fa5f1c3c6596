// pdn-jpegxl_b200 engine — encode orchestration behind SaveImage / JxlB200EncodeToMemory.
// Restates what EncoderWriteImage configures and drives through libjxl
// (N/Encoder/JxlEncoder.cpp:147-392): pixel-format decision from a full scan (:33-77, gray only
// without ICC :67), 8-bit samples, XYB VarDCT for lossy / Modular with the original profile
// for lossless (:214,325), container always with uncompressed Exif / xml boxes (:201,284-310).
// The codec arithmetic runs in dev/encode_kernels.cu; the host writes headers, builds the
// entropy codes from device histograms and assembles the sections. No CPU encode path exists.
#include "engine.h"
#include "dev/enc_frame.cuh"
#include "dev/kernels.h"
#include "host/headers.h"
#include "host/modular_host.h"
#include "host/vardct_tables.h"
#include "host/icc.h"
#include <map>
#include <mutex>
#include <cmath>

namespace jxlgpu {

#define CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) throw Error(std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #expr); } while (0)

namespace {

struct Buf { void* p = nullptr; size_t n = 0; void Alloc(size_t b) { Free(); n = b ? b : 1; if (cudaMalloc(&p, n) != cudaSuccess) { cudaGetLastError(); p = nullptr; throw std::bad_alloc(); } } void Free() { if (p) cudaFree(p); p = nullptr; } ~Buf() { Free(); } template <class T> T* as() const { return static_cast<T*>(p); } };

// ---- fixed MA trees (same shape the decoder kernels walk; BFS layout as DecodeTree allocates it)
struct TreeBuilder {
  struct Tmp { int prop; int32_t split; int l, r; int pred; }; std::vector<Tmp> n;
  int Leaf(int pred) { n.push_back({-1, 0, -1, -1, pred}); return int(n.size()) - 1; }
  int Split(int prop, int32_t val, int gt, int le) { n.push_back({prop, val, gt, le, 0}); return int(n.size()) - 1; }
  int Range(int prop, const std::vector<int32_t>& thr, int lo, int hi, int pred) { if (lo >= hi) return Leaf(pred); int mid = (lo + hi) / 2; int gt = Range(prop, thr, mid + 1, hi, pred); int le = Range(prop, thr, lo, mid, pred); return Split(prop, thr[mid], gt, le); }
  int ResidualCtx(int pred) { static const std::vector<int32_t> thr = {-64, -24, -8, -3, -1, 0, 2, 7, 23, 63}; return Range(8, thr, 0, int(thr.size()), pred); }
  int Channels(int nch, int first, int pred) { if (nch == 1) return ResidualCtx(pred); int mid = first + nch / 2 - 1; int gt = Channels(nch - nch / 2, mid + 1, pred); int le = Channels(nch / 2, first, pred); return Split(0, mid, gt, le); }
  Tree Flatten(int root) {
    Tree tree; std::vector<int> queue{root}; size_t head = 0; int leaf = 0;
    while (head < queue.size()) { const Tmp t = n[queue[head++]]; TreeNode o;
      if (t.prop < 0) { o.property = -1; o.predictor = t.pred; o.leaf_id = leaf++; } else { o.property = t.prop; o.splitval = t.split; o.lchild = int(queue.size()); o.rchild = int(queue.size()) + 1; queue.push_back(t.l); queue.push_back(t.r); }
      tree.push_back(o); }
    return tree;
  }
};
Tree MakeVarDctTree(uint32_t nlf, int num_ec) {
  TreeBuilder b; int sharp = b.Leaf(0), hfmul = b.Leaf(0), strat = b.Leaf(0), cflc = b.Leaf(0);   // Zero predictor: constant maps decode as zero-entropy rows
  int blockinfo = b.Split(2, 0, hfmul, strat); int hfmeta = b.Split(0, 1, b.Split(0, 2, sharp, blockinfo), cflc);
  int groups = num_ec > 0 ? b.Channels(num_ec, 0, 5) : b.Leaf(5); int upper = b.Split(1, int32_t(3 * nlf + 17), groups, hfmeta);
  int lfc = b.Channels(3, 0, 5); int global = b.Leaf(5); int lower = b.Split(1, 0, lfc, global);
  return b.Flatten(b.Split(1, int32_t(2 * nlf), upper, lower));
}
Tree MakeLosslessTree(int nch) { TreeBuilder b; return b.Flatten(b.Channels(nch, 0, 5)); }
// leaf reached for the given properties (only the properties the fixed trees test: 0 channel, 1 stream, 2 y, 8 previous residual)
const TreeNode& LeafFor(const Tree& t, int chan, int stream, int y, int prop8) {
  int i = 0; while (t[i].property >= 0) { int p = t[i].property; int v = p == 0 ? chan : p == 1 ? stream : p == 2 ? y : prop8; i = v > t[i].splitval ? t[i].lchild : t[i].rchild; } return t[i];
}
const int32_t kProp8Rep[11] = {-100, -40, -10, -5, -2, 0, 1, 5, 10, 40, 100};   // one representative per bucket of ResidualCtx

// tiny host-side Modular tokeniser for the HF-metadata image (a few thousand samples): predictors W (1) and gradient (5)
void TokenizeSmallChannel(const std::vector<int32_t>& px, int w, int h, int chan, int stream, const Tree& tree, std::vector<Token>* out) {
  for (int y = 0; y < h; y++) { int32_t prev_grad = 0; for (int x = 0; x < w; x++) {
    auto at = [&](int yy, int xx) { return px[size_t(yy) * w + xx]; };
    int32_t W = x ? at(y, x - 1) : (y ? at(y - 1, x) : 0), N = y ? at(y - 1, x) : W, NW = (x && y) ? at(y - 1, x - 1) : W;
    const TreeNode& leaf = LeafFor(tree, chan, stream, y, W - prev_grad);
    int32_t pred = leaf.predictor == 0 ? 0 : leaf.predictor == 1 ? W : std::max(std::min(W, N), std::min(std::max(W, N), W + N - NW)); JXLG_CHECK(leaf.predictor == 0 || leaf.predictor == 1 || leaf.predictor == 5, "fixed tree predictor");
    out->push_back({uint32_t(leaf.leaf_id), PackSigned(at(y, x) - pred)}); prev_grad = W + N - NW; } }
}

void QuantizerFromDistance(float d, uint32_t* global_scale, uint32_t* quant_lf, float* q_ac) {
  d = std::max(d, 0.01f); float qac = 0.79f / d; float eff = 0.3f * std::pow(d / 0.3f, 0.83f); eff = std::min(d, std::max(0.5f * d, eff)); float qdc = std::min(50.0f, 1.0959f / eff);
  float scale = 65536.0f * qac / 5.0f; scale = std::min(32768.0f, std::max(1.0f, scale)); int gs = int(scale); int sdc = int(qdc * 4096.0f * 1.6f); if (gs > sdc) gs = std::max(1, sdc);
  *global_scale = uint32_t(gs); float v = qdc * (65536.0f / float(gs)) + 0.5f; *quant_lf = uint32_t(std::max(1.0f, std::min(65536.0f, v))); *q_ac = qac;
}

// appends `nbits` bits of a byte buffer to a bit writer at any alignment
void AppendBits(BitWriter& bw, const uint8_t* p, uint64_t nbits) {
  uint64_t full = nbits / 8; size_t i = 0; for (; i + 7 <= full; i += 7) { uint64_t v = 0; memcpy(&v, p + i, 7); bw.Write(56, v); } for (; i < full; i++) bw.Write(8, p[i]);
  int rem = int(nbits & 7); if (rem) bw.Write(rem, p[full] & ((1u << rem) - 1));
}

struct DeviceEncCode { Buf ctx_map, freq, start, rev, desc; };
void UploadEncCode(const EncCode& c, DeviceEncCode* d, cudaStream_t st) {
  size_t ncl = c.cfg.size(); std::vector<uint16_t> freq(ncl * kEncAlphabet, 0), start(ncl * kEncAlphabet, 0), rev(ncl * 4096, 0);
  for (size_t k = 0; k < ncl; k++) { JXLG_CHECK((size_t(1) << c.log_alpha) <= kEncAlphabet, "encoder alphabet"); for (size_t s = 0; s < (size_t(1) << c.log_alpha); s++) { freq[k * kEncAlphabet + s] = c.ans[k].freq[s]; start[k * kEncAlphabet + s] = uint16_t(c.sym_start[k][s]); } memcpy(&rev[k * 4096], c.rev[k].data(), 4096 * 2); }
  d->ctx_map.Alloc(c.ctx_map.size()); d->freq.Alloc(freq.size() * 2); d->start.Alloc(start.size() * 2); d->rev.Alloc(rev.size() * 2); d->desc.Alloc(sizeof(DEncCode));
  CUDA_OK(cudaMemcpyAsync(d->ctx_map.p, c.ctx_map.data(), c.ctx_map.size(), cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(d->freq.p, freq.data(), freq.size() * 2, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(d->start.p, start.data(), start.size() * 2, cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(d->rev.p, rev.data(), rev.size() * 2, cudaMemcpyHostToDevice, st));
  DEncCode h{d->ctx_map.as<uint8_t>(), d->freq.as<uint16_t>(), d->start.as<uint16_t>(), d->rev.as<uint16_t>()}; CUDA_OK(cudaMemcpyAsync(d->desc.p, &h, sizeof(h), cudaMemcpyHostToDevice, st)); CUDA_OK(cudaStreamSynchronize(st));
}

std::vector<std::vector<uint64_t>> HistFromDevice(const uint32_t* d_hist, size_t num_ctx, cudaStream_t st) {
  std::vector<uint32_t> raw(num_ctx * kEncAlphabet); CUDA_OK(cudaMemcpyAsync(raw.data(), d_hist, raw.size() * 4, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  std::vector<std::vector<uint64_t>> h(num_ctx); for (size_t c = 0; c < num_ctx; c++) { size_t last = 0; bool any = false; for (size_t s = 0; s < kEncAlphabet; s++) if (raw[c * kEncAlphabet + s]) { last = s; any = true; } if (any) { h[c].resize(last + 1); for (size_t s = 0; s <= last; s++) h[c][s] = raw[c * kEncAlphabet + s]; } }
  return h;
}
void AddTokensToHist(const std::vector<Token>& toks, const HybridCfg& cfg, std::vector<std::vector<uint64_t>>* h) { for (const Token& t : toks) { uint32_t tok, nb, bits; HybridEncode(cfg, t.value, &tok, &nb, &bits); auto& hh = (*h)[t.ctx]; if (hh.size() <= tok) hh.resize(tok + 1, 0); hh[tok]++; } }

const DTables* EncDeviceTables() {
  static std::mutex mu; static std::map<int, DTables*> per_dev; std::lock_guard<std::mutex> lk(mu); int dev = 0; CUDA_OK(cudaGetDevice(&dev)); auto it = per_dev.find(dev); if (it != per_dev.end()) return it->second;
  std::unique_ptr<DTables> h(new DTables); FillDeviceTables(h.get()); DTables* d = nullptr; CUDA_OK(cudaMalloc(&d, sizeof(DTables))); CUDA_OK(cudaMemcpy(d, h.get(), sizeof(DTables), cudaMemcpyHostToDevice));
  float lut[256]; for (int i = 0; i < 256; i++) { float v = float(i) * (1.0f / 255.0f); lut[i] = v <= 0.04045f ? v / 12.92f : std::pow((v + 0.055f) / 1.055f, 2.4f); } UploadSrgbLut(lut); per_dev[dev] = d; return d;
}

}  // namespace

EncodeResult EncodeOnGpu(const EncodeRequest& req) {
  EncodeResult res;
  if (!req.bgra) { res.status = EncStatus::NullParameter; return res; }
  cudaStream_t st = nullptr;
  try {
    std::string why; if (!CudaAvailable(&why)) { res.status = EncStatus::EncodeError; res.message = why; return res; }
    JXLG_CHECK(req.width > 0 && req.height > 0 && req.stride >= req.width * 4, "invalid bitmap");
    CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t ev0, ev1; cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, st);
    const uint32_t xs = req.width, ys = req.height; const bool lossless = req.lossless;
    // ---- input + scan (GetOutputPixelFormat, N/Encoder/JxlEncoder.cpp:33-77)
    Buf d_in, d_flags; const uint8_t* d_bgra = req.bgra; size_t in_bytes = size_t(req.stride) * ys;
    if (!req.device_input) { d_in.Alloc(in_bytes); CUDA_OK(cudaMemcpyAsync(d_in.p, req.bgra, in_bytes, cudaMemcpyHostToDevice, st)); d_bgra = d_in.as<uint8_t>(); }
    d_flags.Alloc(16); CUDA_OK(cudaMemsetAsync(d_flags.p, 0, 16, st)); EncLaunchScan(d_bgra, xs, ys, req.stride, d_flags.as<uint32_t>(), st);
    uint32_t flags = 0; CUDA_OK(cudaMemcpyAsync(&flags, d_flags.p, 4, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
    const bool is_gray = !(flags & 1) && req.icc_size == 0, has_alpha = (flags & 2) != 0; res.pixel_format = is_gray ? (has_alpha ? 1 : 0) : (has_alpha ? 3 : 2);
    // ---- headers
    ImageMetadata m; m.xsize = xs; m.ysize = ys; m.xyb_encoded = !lossless; m.ce.intent = 0; if (is_gray) m.ce.color_space = kCsGray;   // sRGB, perceptual intent (N/Encoder/JxlEncoder.cpp:269-282)
    // ICC profile given (N/Encoder/JxlEncoder.cpp:258-262): the codestream carries it as an ICC stream. Lossless keeps the samples as they are;
    // lossy needs the source -> XYB mapping, which libjxl gets from a CMS and the engine reads from a matrix/TRC profile directly.
    IccMatrixTrc src_profile; bool has_src_profile = false;
    if (req.icc_size) {
      m.ce.want_icc = true; m.icc.assign(req.icc, req.icc + req.icc_size);
      if (!lossless) { std::string why; JXLG_CHECK(ParseMatrixTrcIcc(req.icc, req.icc_size, &src_profile, &why), why + " — lossy encoding of this profile needs a CMS, which the engine does not have"); has_src_profile = true; }
    }
    if (has_alpha) { ExtraChannelInfo a; a.type = kEcAlpha; m.ec.push_back(a); }
    const int num_ec = has_alpha ? 1 : 0, ncolor = is_gray ? 1 : 3;
    FrameHeader fh; fh.encoding = lossless ? 1 : 0; fh.ec_upsampling.assign(num_ec, 1); fh.ec_blending.assign(num_ec, BlendingInfo());
    if (lossless) { fh.group_size_shift = 1; fh.lf.gab = false; fh.lf.epf_iters = 0; }
    else { bool hi = req.effort >= 5; int epf = 0; if (hi) { const float thr[3] = {0.7f, 1.5f, 4.0f}; for (float t : thr) if (req.distance >= t) epf++; } fh.lf.gab = hi; fh.lf.epf_iters = uint32_t(epf); }
    DeriveFrameDims(fh, m); const uint32_t nlf = fh.num_lf_groups, ng = fh.num_groups; const size_t nsec = NumTocEntries(fh); const bool single = nsec == 1; const uint32_t gd = fh.group_dim;
    Tree tree = lossless ? MakeLosslessTree(ncolor + num_ec) : MakeVarDctTree(nlf, num_ec); std::vector<Token> tree_tokens; TokenizeTree(tree, &tree_tokens);
    EncOptions topt; topt.cfg = HybridCfg{4, 1, 0}; EncOptions mopt; mopt.cfg = HybridCfg{4, 1, 0}; mopt.max_clusters = 48; EncOptions aopt; aopt.cfg = HybridCfg{4, 2, 0}; aopt.max_clusters = 64;
    EncCode tree_code = BuildCode({&tree_tokens}, 6, topt); const size_t nleaves = NumLeaves(tree);
    // leaf LUT for the device tokeniser: kind 0 = LF coefficient streams, 1 = pass-group streams, 2 = global stream
    std::vector<uint16_t> leaf_lut(3 * 8 * 11, 0);
    for (int kind = 0; kind < 3; kind++) for (int c = 0; c < 8; c++) for (int b = 0; b < 11; b++) { int stream = kind == 0 ? 1 : kind == 1 ? int(1 + 3 * nlf + 17) : 0; leaf_lut[(kind * 8 + c) * 11 + b] = uint16_t(LeafFor(tree, c, stream, 1, kProp8Rep[b]).leaf_id); }
    // ---- device frame
    DEncFrame e; memset(&e, 0, sizeof(e)); e.xsize = xs; e.ysize = ys; e.stride = req.stride; e.xb = fh.xblocks; e.yb = fh.yblocks; e.xpad = e.xb * 8; e.ypad = e.yb * 8; e.xgroups = fh.xgroups; e.ygroups = fh.ygroups; e.num_groups = ng; e.gray = is_gray; e.alpha = has_alpha;
    Buf d_srclut; if (has_src_profile) { d_srclut.Alloc(sizeof(src_profile.lut)); CUDA_OK(cudaMemcpyAsync(d_srclut.p, src_profile.lut, sizeof(src_profile.lut), cudaMemcpyHostToDevice, st)); e.src_lut = d_srclut.as<float>(); e.has_src_profile = 1; for (int i = 0; i < 9; i++) e.src_matrix[i] = float(src_profile.to_linear_srgb[i]); }
    e.tables = EncDeviceTables(); const size_t npx = size_t(xs) * ys, ppx = size_t(e.xpad) * e.ypad, cells = size_t(e.xb) * e.yb;
    const int nplanes = lossless ? ncolor + num_ec : num_ec; e.alpha_plane = lossless ? uint32_t(ncolor) : 0;
    uint32_t global_scale = 1, quant_lf = 16; float q_ac = 1;
    Buf d_e, d_xyb, d_tmp1, d_tmp2, d_planes, d_lf, d_lfq, d_coeffs, d_nz, d_dq, d_order, d_tokens, d_account, d_lut, d_modstreams, d_streams, d_hist_m, d_hist_a, d_bytes, d_bits, d_gabframe;
    d_e.Alloc(sizeof(DEncFrame)); if (nplanes) d_planes.Alloc(npx * nplanes * 4); e.planes = d_planes.as<int32_t>();
    d_lut.Alloc(leaf_lut.size() * 2); CUDA_OK(cudaMemcpyAsync(d_lut.p, leaf_lut.data(), leaf_lut.size() * 2, cudaMemcpyHostToDevice, st));
    // modular group streams (alpha for VarDCT, everything for lossless): all planes are full-size here
    const bool groups_have_modular = nplanes > 0 && (xs > gd || ys > gd); const bool global_has_modular = nplanes > 0 && !groups_have_modular;
    std::vector<DEncModStream> mod_streams; std::vector<DEncStream> m_streams; uint64_t token_cursor = 0, byte_cursor = 0; uint32_t max_mod_tokens = 0;
    auto add_mod_stream = [&](uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, uint32_t kind, uint32_t nch) { DEncModStream s{x0, y0, w, h, kind, 0, token_cursor}; mod_streams.push_back(s); uint32_t cnt = w * h * nch; DEncStream es{token_cursor, byte_cursor, cnt, 0}; m_streams.push_back(es); token_cursor += cnt; byte_cursor += (size_t(cnt) * 6 + 16 + 15) / 16 * 16; max_mod_tokens = std::max(max_mod_tokens, cnt); };
    // stream index bookkeeping: [LF groups (VarDCT)] [modular groups or global]
    const uint32_t first_lf_stream = 0; if (!lossless) for (uint32_t g = 0; g < nlf; g++) { uint32_t gx = g % fh.xlfgroups, gy = g / fh.xlfgroups; add_mod_stream(gx * 256, gy * 256, std::min<uint32_t>(256, e.xb - gx * 256), std::min<uint32_t>(256, e.yb - gy * 256), 0, 3); }
    const uint32_t first_group_stream = uint32_t(mod_streams.size());
    if (groups_have_modular) for (uint32_t g = 0; g < ng; g++) { uint32_t gx = g % fh.xgroups, gy = g / fh.xgroups; add_mod_stream(gx * gd, gy * gd, std::min(gd, xs - gx * gd), std::min(gd, ys - gy * gd), 1, uint32_t(nplanes)); }
    else if (global_has_modular) add_mod_stream(0, 0, xs, ys, 2, uint32_t(nplanes));
    const uint64_t ac_token_off = token_cursor; if (!lossless) token_cursor += uint64_t(ng) * kMaxAcTokensPerGroup; e.ac_token_off = ac_token_off;
    d_tokens.Alloc(std::max<uint64_t>(token_cursor, 1) * 8); e.tokens = d_tokens.as<uint2>();
    if (!lossless) {
      QuantizerFromDistance(req.distance, &global_scale, &quant_lf, &q_ac); const float inv_gs = 65536.0f / float(global_scale); OpsinInverse op;
      e.hf_mul = uint32_t(std::max(1, std::min(255, int(std::lrintf(q_ac * 65536.0f / float(global_scale)))))); e.inv_gs = inv_gs; e.xm = std::pow(0.8f, float(fh.x_qm_scale) - 2.0f); e.bm = std::pow(0.8f, float(fh.b_qm_scale) - 2.0f); e.kx = 0.f; e.kb = 1.f;
      const float lfd[3] = {1.0f / 4096, 1.0f / 512, 1.0f / 256}; for (int c = 0; c < 3; c++) e.lf_fac[c] = lfd[c] * inv_gs / float(quant_lf); e.cfl_x_lf = 0.f; e.cfl_b_lf = 1.f; for (int i = 0; i < 4; i++) e.quant_bias[i] = op.quant_bias[i];
      d_xyb.Alloc(ppx * 12); d_lf.Alloc(cells * 12); d_lfq.Alloc(cells * 12); d_coeffs.Alloc(size_t(ng) * 3 * 65536 * 2); d_nz.Alloc(cells * 3); d_account.Alloc(size_t(ng) * 4);
      std::vector<float> dq = ComputeDequantTable(0, LibraryEncoding(0)); d_dq.Alloc(dq.size() * 4); CUDA_OK(cudaMemcpyAsync(d_dq.p, dq.data(), dq.size() * 4, cudaMemcpyHostToDevice, st));
      std::vector<uint32_t> nat = NaturalOrder(1, 1); std::vector<uint16_t> o16(nat.begin(), nat.end()); d_order.Alloc(128); CUDA_OK(cudaMemcpyAsync(d_order.p, o16.data(), 128, cudaMemcpyHostToDevice, st));
      e.xyb = d_xyb.as<float>(); e.lf = d_lf.as<float>(); e.lfq = d_lfq.as<int32_t>(); e.coeffs = d_coeffs.as<int16_t>(); e.nz = d_nz.as<uint8_t>(); e.dequant8 = d_dq.as<float>(); e.order8 = d_order.as<uint16_t>(); e.ac_token_count = d_account.as<uint32_t>();
      CUDA_OK(cudaMemsetAsync(d_coeffs.p, 0, d_coeffs.n, st));
    }
    d_bytes.Alloc(std::max<uint64_t>(byte_cursor + (lossless ? 0 : uint64_t(ng) * (size_t(kMaxAcTokensPerGroup) * 6 + 16)), 16)); d_bits.Alloc((mod_streams.size() + ng + 1) * 8); e.stream_bytes = d_bytes.as<uint8_t>(); e.stream_bits = d_bits.as<uint64_t>();
    CUDA_OK(cudaMemcpyAsync(d_e.p, &e, sizeof(e), cudaMemcpyHostToDevice, st)); const DEncFrame* de = d_e.as<DEncFrame>();
    // ---- pixels -> planes / XYB -> DCT + quant
    if (lossless) EncLaunchToPlanes(de, e, d_bgra, st);
    else {
      EncLaunchToXyb(de, e, d_bgra, st);   // e.src_lut / e.src_matrix were filled above when an ICC source profile is in use
      if (fh.lf.gab) {   // approximate inverse gaborish: two Van Cittert iterations against the decoder's own kernel (padded frame, mirrored edges)
        d_tmp1.Alloc(ppx * 12); d_tmp2.Alloc(ppx * 12); d_gabframe.Alloc(sizeof(DFrame)); DFrame gf; memset(&gf, 0, sizeof(gf)); gf.xsize = e.xpad; gf.ysize = e.ypad; gf.xpad = e.xpad; gf.ypad = e.ypad; gf.lpf.gab = 1; memcpy(gf.lpf.gab_w, fh.lf.gab_w, sizeof(gf.lpf.gab_w));
        CUDA_OK(cudaMemcpyAsync(d_gabframe.p, &gf, sizeof(gf), cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(d_tmp1.p, d_xyb.p, ppx * 12, cudaMemcpyDeviceToDevice, st));
        for (int it = 0; it < 2; it++) { LaunchGaborishPlanes(d_gabframe.as<DFrame>(), gf, e.xyb, d_tmp2.as<float>(), st); EncLaunchSharpen(e.xyb, d_tmp1.as<float>(), d_tmp2.as<float>(), ppx * 3, st); }
      }
      EncLaunchDct8(de, e, st); EncLaunchAcTokens(de, e, st);
    }
    // ---- modular tokens on the device
    d_modstreams.Alloc(std::max<size_t>(mod_streams.size(), 1) * sizeof(DEncModStream)); if (!mod_streams.empty()) CUDA_OK(cudaMemcpyAsync(d_modstreams.p, mod_streams.data(), mod_streams.size() * sizeof(DEncModStream), cudaMemcpyHostToDevice, st));
    if (!lossless && nlf) { uint32_t mx = 0; for (uint32_t g = 0; g < nlf; g++) mx = std::max(mx, m_streams[first_lf_stream + g].count); EncLaunchModTokens(de, d_modstreams.as<DEncModStream>() + first_lf_stream, nlf, mx, e.lfq, e.xb, e.yb, 3, d_lut.as<uint16_t>(), st); }
    if (mod_streams.size() > first_group_stream) { uint32_t n = uint32_t(mod_streams.size()) - first_group_stream, mx = 0; for (uint32_t i = 0; i < n; i++) mx = std::max(mx, m_streams[first_group_stream + i].count);
      EncLaunchModTokens(de, d_modstreams.as<DEncModStream>() + first_group_stream, n, mx, e.planes, xs, ys, uint32_t(nplanes), d_lut.as<uint16_t>(), st); }
    // ---- AC stream descriptors need the per-group token counts
    std::vector<DEncStream> a_streams; std::vector<uint32_t> ac_counts(ng, 0);
    if (!lossless) { CUDA_OK(cudaMemcpyAsync(ac_counts.data(), d_account.p, size_t(ng) * 4, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
      for (uint32_t g = 0; g < ng; g++) { DEncStream s{ac_token_off + uint64_t(g) * kMaxAcTokensPerGroup, byte_cursor, ac_counts[g], 0}; a_streams.push_back(s); byte_cursor += (size_t(ac_counts[g]) * 6 + 16 + 15) / 16 * 16; } }
    std::vector<DEncStream> all = m_streams; all.insert(all.end(), a_streams.begin(), a_streams.end());
    d_streams.Alloc(std::max<size_t>(all.size(), 1) * sizeof(DEncStream)); if (!all.empty()) CUDA_OK(cudaMemcpyAsync(d_streams.p, all.data(), all.size() * sizeof(DEncStream), cudaMemcpyHostToDevice, st));
    const DEncStream* d_m = d_streams.as<DEncStream>(); const DEncStream* d_a = d_m + m_streams.size();
    // ---- histograms -> entropy codes (host: clustering, normalisation, alias tables)
    d_hist_m.Alloc(nleaves * kEncAlphabet * 4); CUDA_OK(cudaMemsetAsync(d_hist_m.p, 0, d_hist_m.n, st)); EncLaunchHistogram(e.tokens, d_m, uint32_t(m_streams.size()), max_mod_tokens, d_hist_m.as<uint32_t>(), st);
    const size_t n_ac_ctx = size_t(495) * 15; uint32_t max_ac = 0; for (uint32_t c : ac_counts) max_ac = std::max(max_ac, c);
    if (!lossless) { d_hist_a.Alloc(n_ac_ctx * kEncAlphabet * 4); CUDA_OK(cudaMemsetAsync(d_hist_a.p, 0, d_hist_a.n, st)); EncLaunchHistogram(e.tokens, d_a, ng, max_ac, d_hist_a.as<uint32_t>(), st); }
    std::vector<std::vector<uint64_t>> hm = HistFromDevice(d_hist_m.as<uint32_t>(), nleaves, st);
    // HF metadata (tiny) is tokenised on the host: CfL maps all zero, every block DCT8 with one hf multiplier, constant EPF sharpness
    std::vector<std::vector<Token>> hfmeta_tokens(nlf); std::vector<uint32_t> hfmeta_nb(nlf, 0);
    if (!lossless) for (uint32_t g = 0; g < nlf; g++) { uint32_t gx = g % fh.xlfgroups, gy = g / fh.xlfgroups; int w = int(std::min<uint32_t>(256, e.xb - gx * 256)), h = int(std::min<uint32_t>(256, e.yb - gy * 256)), tw = (w + 7) / 8, th = (h + 7) / 8, nb = w * h; hfmeta_nb[g] = uint32_t(nb);
      int sid = int(1 + 2 * nlf + g); std::vector<int32_t> zeros(size_t(tw) * th, 0), info(size_t(nb) * 2, 0), sharp(size_t(w) * h, fh.lf.epf_iters ? 4 : 0); for (int i = 0; i < nb; i++) info[nb + i] = int32_t(e.hf_mul) - 1;
      TokenizeSmallChannel(zeros, tw, th, 0, sid, tree, &hfmeta_tokens[g]); TokenizeSmallChannel(zeros, tw, th, 1, sid, tree, &hfmeta_tokens[g]); TokenizeSmallChannel(info, nb, 2, 2, sid, tree, &hfmeta_tokens[g]); TokenizeSmallChannel(sharp, w, h, 3, sid, tree, &hfmeta_tokens[g]);
      AddTokensToHist(hfmeta_tokens[g], mopt.cfg, &hm); }
    EncCode mcode = BuildCodeFromHist(hm, nleaves, mopt); EncCode acode; DeviceEncCode dm, da; UploadEncCode(mcode, &dm, st);
    if (!lossless) { acode = BuildCodeFromHist(HistFromDevice(d_hist_a.as<uint32_t>(), n_ac_ctx, st), n_ac_ctx, aopt); UploadEncCode(acode, &da, st); }
    // ---- ANS streams on the device
    EncLaunchAns(de, d_m, uint32_t(m_streams.size()), dm.desc.as<DEncCode>(), st); if (!lossless) { /* stream_bits index continues after the modular streams */ }
    std::vector<uint64_t> bits_m(m_streams.size(), 0), bits_a(ng, 0);
    if (!m_streams.empty()) CUDA_OK(cudaMemcpyAsync(bits_m.data(), e.stream_bits, bits_m.size() * 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    if (!lossless) { EncLaunchAns(de, d_a, ng, da.desc.as<DEncCode>(), st); CUDA_OK(cudaMemcpyAsync(bits_a.data(), e.stream_bits, size_t(ng) * 8, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st)); }
    std::vector<uint8_t> bytes(byte_cursor); if (byte_cursor) CUDA_OK(cudaMemcpyAsync(bytes.data(), d_bytes.p, byte_cursor, cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
    cudaEventRecord(ev1, st); cudaEventSynchronize(ev1); cudaEventElapsedTime(&res.times.total, ev0, ev1); cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    // ---- sections (host)
    std::vector<BitWriter> secw(single ? 1 : nsec); auto W = [&](size_t i) -> BitWriter& { return single ? secw[0] : secw[i]; };
    GroupHeader plain; plain.use_global_tree = true; GroupHeader gheader = plain; if (lossless && ncolor == 3) { Transform t; t.id = 0; t.begin_c = 0; t.rct_type = 6; gheader.transforms.push_back(t); }
    { BitWriter& bw = W(0);
      bw.Bool(true);   /* LfChannelDequantization all_default: present for Modular frames too */
      if (!lossless) { bw.U32(BitsOffset(11, 1), BitsOffset(11, 2049), BitsOffset(12, 4097), BitsOffset(16, 8193), global_scale); bw.U32(Val(16), BitsOffset(5, 1), BitsOffset(8, 1), BitsOffset(16, 1), quant_lf); bw.Bool(true); bw.Bool(true); }
      bw.Bool(true); WriteCode(bw, tree_code); WriteTokens(bw, tree_code, tree_tokens); WriteCode(bw, mcode);
      if (nplanes > 0) { WriteGroupHeader(bw, gheader); if (global_has_modular) { const DEncStream& s = m_streams[first_group_stream]; AppendBits(bw, &bytes[s.byte_off], bits_m[first_group_stream]); } } }
    for (uint32_t g = 0; g < nlf; g++) { BitWriter& bw = W(1 + g); if (lossless) continue;
      bw.Write(2, 0); WriteGroupHeader(bw, plain); AppendBits(bw, &bytes[m_streams[first_lf_stream + g].byte_off], bits_m[first_lf_stream + g]);
      uint32_t gx = g % fh.xlfgroups, gy = g / fh.xlfgroups; uint64_t wh = uint64_t(std::min<uint32_t>(256, e.xb - gx * 256)) * std::min<uint32_t>(256, e.yb - gy * 256);
      bw.Write(CeilLog2(wh), hfmeta_nb[g] - 1); WriteGroupHeader(bw, plain); WriteTokens(bw, mcode, hfmeta_tokens[g]); }
    { BitWriter& bw = W(1 + nlf); if (!lossless) { bw.Bool(true); bw.Write(CeilLog2(ng), 0); bw.U32(Val(0x5F), Val(0x13), Val(0), Bits(13), 0); WriteCode(bw, acode); } }
    for (uint32_t g = 0; g < ng; g++) { BitWriter& bw = W(2 + nlf + g);
      if (!lossless) AppendBits(bw, &bytes[a_streams[g].byte_off], bits_a[g]);
      if (groups_have_modular) { WriteGroupHeader(bw, plain); AppendBits(bw, &bytes[m_streams[first_group_stream + g].byte_off], bits_m[first_group_stream + g]); } }
    BitWriter cs; cs.Write(16, 0x0AFF); WriteImageHeaders(cs, m); WriteFrameHeader(cs, fh, m);
    std::vector<std::vector<uint8_t>> secs; std::vector<size_t> sizes; for (auto& w : secw) secs.push_back(w.Finish()); for (auto& s : secs) sizes.push_back(s.size());
    WriteToc(cs, sizes); std::vector<uint8_t> code = cs.Finish(); for (auto& s : secs) code.insert(code.end(), s.begin(), s.end());
    // container always (JxlEncoderUseBoxes, N/Encoder/JxlEncoder.cpp:201); Exif / xml boxes uncompressed, blobs passed through (:284-310)
    res.file = ContainerPrologue(); if (req.exif_size) AppendBox(res.file, "Exif", req.exif, req.exif_size); if (req.xmp_size) AppendBox(res.file, "xml ", req.xmp, req.xmp_size); AppendBox(res.file, "jxlc", code.data(), code.size());
  } catch (const std::bad_alloc&) { res.status = EncStatus::OutOfMemory; }
  catch (const std::exception& ex) { res.status = EncStatus::EncodeError; res.message = ex.what(); }
  if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  return res;
}

}  // namespace jxlgpu
