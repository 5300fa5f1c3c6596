// pdn-jpegxl_b200 engine — per-frame device descriptors (host fills, kernels read).
// One DFrame describes one JPEG XL frame being decoded on one GPU: geometry,
// quantiser / chroma-from-luma / block-context parameters from LfGlobal (A.8), entropy codes
// and the MA tree (A.6/A.7), loop-filter and colour parameters (A.10), plus the device
// buffers the stage kernels pass data through. Layout in HBM is documented in DESIGN.md.
#pragma once
#include "common.cuh"
#include "modular.cuh"

namespace jxlgpu {

static const int kMaxPasses = 11;

struct DLoopFilter {
  uint32_t gab, epf_iters; float gab_w[6]; float epf_sharp_lut[8]; float epf_channel_scale[3]; float epf_quant_mul, pass0_sigma_scale, pass2_sigma_scale, border_sad_mul, sigma_for_modular;
};
struct DColor {
  float opsin_inv[9]; float opsin_bias[3]; float opsin_bias_cbrt[3]; float itscale; float intensity_target; float mix_to_target[9];   // to_target * opsin_inv * itscale, folded on the host: one 3x3 product per pixel
  float to_target[9];   // linear sRGB -> target primaries
  uint32_t tf; float gamma; uint32_t xyb_encoded; uint32_t num_color;   // tf: host ColorEncoding tf enum, 0 = pure gamma (gamma = exponent)
};
struct DOutput {
  uint32_t sample_type;    // 0 u8, 1 u16, 2 f16, 3 f32
  uint32_t num_channels;   // colour + alpha as the reference counts them (N/Decoder/JxlDecoder.cpp:494)
  uint32_t color_channels; int32_t alpha_plane; int32_t black_plane; uint32_t premultiplied; uint32_t orientation; uint32_t bgra;   // bgra: fused BGRA32 surface pack (S/JpegXLLoad.cs:219-249)
  uint32_t out_w, out_h;   // post-orientation dimensions
  uint32_t alpha_bits, alpha_exp_bits, black_bits, bits, exp_bits;
};
struct DModChannel { uint32_t w, h, hshift, vshift; uint64_t plane_off; };   // plane_off: int32 element offset into DFrame::mod_planes
// One inverse transform of the global Modular image, in the order the kernels run them (the reverse of the order the file lists them).
// kind 0 RCT: planes p[0..2] of n samples, in place. kind 1 Palette: index plane p[0] (w x h) and palette plane p[1] (pal_w entries per row,
// one row per output channel) expand into num_c new planes out[0..num_c) — out of place, the index plane is read by every output channel.
// A Modular sub-bitstream that brings its own MA tree and entropy code (use_global_tree = 0): parsed on the host, tables in the blob. One entry per
// group section of a Modular frame (index g), one for the global stream (index num_groups), one per LF-group section (num_groups + 1 + g);
// data_bitpos: where the channel data starts.
struct DLocalTree { uint32_t present, tree_off, tree_size, uses_wp; uint64_t data_bitpos; DCode code; };
// kind 2 Squeeze (one channel of one squeeze step): average plane p[0] (w x h) and residual plane p[1] (rw x rh) interleave into the new plane out[0]
// ((w + rw) x h when rct_type != 0 = horizontal, w x (h + rh) otherwise); num_c / pal_w carry rw / rh.
struct DModOp { uint32_t kind, rct_type, num_c, pal_w, nb_deltas, predictor, w, h; uint64_t p[3]; uint64_t out[4]; };

// Per-device constant tables (built once): scaled DCT cosines c[k*N+i] = ck*cos((2i+1)k*pi/2N) for N = 1..256 and
// the LF->LLF resample scales. cos_off[log2 N] is the float offset of the N x N table.
struct DTables { float cosines[1 + 4 + 16 + 64 + 256 + 1024 + 4096 + 16384 + 65536]; float resample[6][32]; };
__host__ __device__ inline uint32_t CosOff(int log2n) { const uint32_t o[9] = {0, 1, 5, 21, 85, 341, 1365, 5461, 21845}; return o[log2n]; }

struct DFrame {
  // Band decode (one rank's share of a frame that is sharded by group rows): output rows [out_y0, out_y1) = group rows
  // [out_g0, out_g1); entropy decode + reconstruction run on group rows [comp_g0, comp_g1) (one extra row each side: the filters' halo).
  uint32_t band_on, out_g0, out_g1, comp_g0, comp_g1, out_y0, out_y1, band_pad;
  uint32_t xsize, ysize, xb, yb, xpad, ypad, xt, yt, xgroups, ygroups, num_groups, xlfgroups, ylfgroups, num_lf_groups, group_dim, num_passes, encoding, flags;
  uint32_t pass_shift[kMaxPasses]; int32_t pass_min_shift[kMaxPasses], pass_max_shift[kMaxPasses];
  float lf_fac[3], cfl_x_lf, cfl_b_lf, inv_gs, xm, bm, base_x, base_b, inv_color_factor, quant_bias[4], quant_scale;
  uint32_t nb_block_ctx, num_lf_ctxs, n_lf_thr[3], n_qf_thr; int32_t lf_thr[3][15]; uint32_t qf_thr[15]; uint32_t bctx_map_off;
  uint32_t num_hf_presets; DCode mod_code; uint32_t has_tree, tree_off, tree_size, uses_wp; DCode ac_code[kMaxPasses]; uint32_t order_off[kMaxPasses][13 * 3];
  uint32_t dq_off[17];       // float[3*size] per quant table, byte offsets into blob
  uint32_t lf_smem, ac_smem, lf_cta_offset, ac_cta_offset, ac_fast; // dynamic shared memory budgets (bytes) for the table staging of k_lf_group / k_ac_group
  uint32_t sec_off;          // uint64 sec_bitpos[nsec] then uint64 sec_bitend[nsec], byte offset into blob
  uint32_t num_mod_channels, first_group_channel; DModChannel mod_ch[8];   // the channels as CODED (after the file's forward transforms): what the entropy kernels fill
  DModChannel out_ch[8];   // the image's own channels (colour, then extra channels) after the inverse transforms: what the output kernels read
  uint32_t mod_ch_off, ops_off;   // more than 8 coded channels / more than 4 ops (squeeze): the tables live in the blob at these offsets (0: the arrays here)
  uint32_t mod_bitdepth, mod_wide; uint32_t num_ops, local_off /* DLocalTree[num_groups + 1] in the blob, 0: every stream uses the global tree */; DModOp ops[4]; DWPHeader global_wp;   // weighted-predictor parameters of the global stream's header
  DLoopFilter lpf; DColor color; DOutput out;
  // device buffers
  const uint8_t* comp; const uint8_t* blob; const uint8_t* static_blob;
  int32_t* lfq; float* lf; float* lf_tmp; uint8_t* acs; uint8_t* hf_mul_m1; uint8_t* sharp; uint8_t* lf_idx; int8_t* ytox; int8_t* ytob; int32_t* hfmeta_scratch;
  const float* lf_src; const struct DTables* tables;
  int16_t* coeffs; float* xyb; float* xyb_tmp; float* inv_sigma; int32_t* mod_planes; int32_t* wp_scratch; uint8_t* out_px; uint32_t* err; uint64_t* end_bitpos; uint64_t* ac_endpos; uint8_t* nz_scratch; uint32_t* host_flags; int32_t* group_pal;   /* kGroupPalInts ints per group: the palette channels of palettes listed in group sections' own headers */ uint32_t* lz_window;   /* LZ77 windows: kLzWindow values per stream (slots: LF group g | group g | last: global), null when no code uses LZ77 */ uint32_t* group_other;   // group_other[g]: number of varblocks in group g that are not plain DCT8
};
// A bundle of images decoded by one launch of an entropy kernel (batches): passed by value like DFrame (4 x 3.4 KB of the 32 KB
// parameter space). CTA `first[i]` .. `first[i+1]-1` (after `cta_offset` empty CTAs) belong to image i. Raises the number of images
// in flight past the 128-resident-grid limit of the device.
static const int kMaxBundle = 4;
struct DFrameSet { uint32_t n, cta_offset; uint32_t first[kMaxBundle + 1]; uint32_t pad; DFrame f[kMaxBundle]; };
static const uint32_t kGroupPalInts = 4096;   // colours x channels a group section's own palettes may hold
static const uint32_t kHfMetaScratchInts = 2 * 1024 + 2 * 65536 + 65536;

// Offsets with the top bit set address the per-device static blob (default dequant tables, natural coefficient orders)
// instead of the per-image blob: defaults are uploaded once per device, not once per image.
static const uint32_t kStaticBlobBit = 0x80000000u;
#ifdef __CUDACC__
__device__ __forceinline__ const uint8_t* BlobAt(const DFrame& f, uint32_t off) { return (off & kStaticBlobBit) ? f.static_blob + (off & ~kStaticBlobBit) : f.blob + off; }
// The strategy tables as packed nibbles (27 entries in two 64-bit immediates): register arithmetic only. The array forms below are
// rebuilt on the stack at every call site when the index is a run-time value (ncu/SASS: 98 STL + 84 LDL in the AC block set-up).
__device__ __forceinline__ uint32_t NibbleAt(uint64_t lo, uint64_t hi, int s) { return uint32_t(((s < 16 ? lo : hi) >> ((s & 15) * 4)) & 15u); }
__device__ __forceinline__ uint32_t CoveredXLog2Dev(int s) { return NibbleAt(0x212010210000ull, 0x54543432300ull, s); }
__device__ __forceinline__ uint32_t CoveredYLog2Dev(int s) { return NibbleAt(0x120201210000ull, 0x45534423300ull, s); }
__device__ __forceinline__ uint32_t StrategyOrderDev(int s) { return NibbleAt(0x1111665544321110ull, 0xccbaa988711ull, s); }
__device__ __forceinline__ const DModChannel& ModCh(const DFrame& f, uint32_t c) { return f.mod_ch_off ? reinterpret_cast<const DModChannel*>(f.blob + f.mod_ch_off)[c] : f.mod_ch[c]; }
__device__ __forceinline__ const DModOp& ModOp(const DFrame& f, uint32_t i) { return f.ops_off ? reinterpret_cast<const DModOp*>(f.blob + f.ops_off)[i] : f.ops[i]; }
__device__ __forceinline__ bool GroupInBand(const DFrame& f, int g) { if (!f.band_on) return true; const uint32_t gy = uint32_t(g) / f.xgroups; return gy >= f.comp_g0 && gy < f.comp_g1; }
__device__ __forceinline__ const uint64_t* SecBitPos(const DFrame& f) { return reinterpret_cast<const uint64_t*>(f.blob + f.sec_off); }
#endif

// AC strategy geometry tables (SURVEY.md A.8 "AC strategies")
__host__ __device__ inline int CoveredX(int s) { const uint8_t t[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32}; return t[s]; }
__host__ __device__ inline int CoveredY(int s) { const uint8_t t[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16}; return t[s]; }
__host__ __device__ inline int StrategyOrder(int s) { const uint8_t t[27] = {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12}; return t[s]; }
__host__ __device__ inline int QuantTableOf(int s) { const uint8_t t[27] = {0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10, 11, 12, 12, 13, 14, 14, 15, 16, 16}; return t[s]; }
// "cell-chunked" coefficient addressing inside a 256x256 group: storage position p of the varblock whose first cell
// is (by,bx) lives in the (p/64)-th covered cell (raster order inside the block).
__host__ __device__ inline uint32_t CoefAddr(int by, int bx, int bw, uint32_t p) { uint32_t j = p >> 6; return (uint32_t(by + int(j) / bw) * 32u + uint32_t(bx + int(j) % bw)) * 64u + (p & 63u); }

}  // namespace jxlgpu
