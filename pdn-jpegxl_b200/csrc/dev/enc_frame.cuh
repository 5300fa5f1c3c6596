// pdn-jpegxl_b200 engine — device descriptors of the encode pipeline (SaveImage path).
#pragma once
#include "frame.cuh"

namespace jxlgpu {

static const uint32_t kEncAlphabet = 128;                    // token-symbol stride of the device histograms / frequency tables
static const uint32_t kMaxAcTokensPerGroup = 3 * 1024 * 64;  // 3 channels x 1024 blocks x (1 + 63) tokens

struct DEncFrame {
  uint32_t xsize, ysize, stride, xpad, ypad, xb, yb, xgroups, ygroups, num_groups, gray, alpha, alpha_plane, hf_mul;
  float inv_gs, xm, bm, kx, kb, lf_fac[3], cfl_x_lf, cfl_b_lf, quant_bias[4];
  // XYB planes cover `ext_rows` rows: `ext_top` halo rows, the ypad rows of the frame (or of this band of it), halo rows below. A whole frame has no
  // halo (ext_top = 0, ext_rows = ypad). Source rows are clamped to [src_row_min, src_row_max] (relative to the first row of the band): the rows
  // the caller handed over, so that a band's halo holds the neighbouring band's pixels and the bottom padding repeats the frame's last row.
  uint32_t ext_top, ext_rows; int32_t src_row_min, src_row_max;
  float* xyb; int32_t* planes; float* lf; int32_t* lfq; int16_t* coeffs; uint8_t* nz; const float* dequant8; const uint16_t* order8; const DTables* tables;
  const float* src_lut; float src_matrix[9]; uint32_t has_src_profile, pad0;   // ICC-described source: per-channel tone LUT [3][256] + matrix to linear sRGB (null / 0: sRGB input)
  // Effort >= 5 (null otherwise): per-block quantiser multiplier (adaptive quantisation) and per-64x64-tile chroma-from-luma factors, both
  // derived on the device from block statistics (k_enc_block_stats, k_enc_tile_params) and coded in the HF metadata.
  uint8_t* hf_mul_map; int8_t* ytox_map; int8_t* ytob_map; float* block_stats; uint32_t xt, yt; float aq_ref; uint32_t aq_on, cfl_on, pad1;
  uint2* tokens; uint64_t ac_token_off; uint32_t* ac_token_count; uint8_t* stream_bytes; uint64_t* stream_bits;
};
struct DEncModStream { uint32_t x0, y0, w, h, kind, pad; uint64_t token_off; };
struct DEncStream { uint64_t token_off; uint64_t byte_off; uint32_t count, pad; };
struct DEncCode { const uint8_t* ctx_map; const uint16_t* freq; const uint16_t* start; const uint16_t* rev; };

void UploadSrgbLut(const float* lut);
void EncLaunchScan(const uint8_t* bgra, uint32_t w, uint32_t h, uint32_t stride, uint32_t* flags, cudaStream_t st);
void EncLaunchToXyb(const DEncFrame* d, const DEncFrame& h, const uint8_t* bgra, cudaStream_t st);
void EncLaunchToPlanes(const DEncFrame* d, const DEncFrame& h, const uint8_t* bgra, cudaStream_t st);
void EncLaunchSharpen(float* cur, const float* orig, const float* blur, size_t n, cudaStream_t st);
void EncLaunchDct8(const DEncFrame* d, const DEncFrame& h, cudaStream_t st);
void EncLaunchBlockParams(const DEncFrame* d, const DEncFrame& h, cudaStream_t st);   // block statistics -> hf_mul_map, ytox_map, ytob_map (before EncLaunchDct8)
void EncLaunchModTokens(const DEncFrame* d, const DEncModStream* streams, uint32_t nstreams, uint32_t max_tokens, const int32_t* planes, uint32_t pw, uint32_t ph, uint32_t nch, const uint16_t* leaf_lut, cudaStream_t st);
void EncLaunchAcTokens(const DEncFrame* d, const DEncFrame& h, cudaStream_t st);
void EncLaunchHistogram(const uint2* tokens, const DEncStream* streams, uint32_t nstreams, uint32_t max_count, uint32_t* hist, cudaStream_t st);
void EncLaunchPrefix(const DEncFrame* d, const DEncStream* streams, uint32_t nstreams, const DEncCode* code, uint32_t bits_off, cudaStream_t st);   // prefix codes: code->freq = lengths, code->start = bit-reversed code words
void EncLaunchAns(const DEncFrame* d, const DEncStream* streams, uint32_t nstreams, const DEncCode* code, uint32_t bits_off, cudaStream_t st);   // stream si reports its bit count in stream_bits[bits_off + si]
void EncLaunchCompact(const uint8_t* src, const DEncStream* streams, uint32_t nstreams, const uint64_t* bits, const uint64_t* dst_off, uint8_t* dst, cudaStream_t st);
void LaunchGaborishPlanes(const DFrame* d, const DFrame& h, const float* src, float* dst, cudaStream_t st);

}  // namespace jxlgpu
