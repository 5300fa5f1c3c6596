// pdn-jpegxl_b200 engine — device-side primitives shared by every entropy-decoding kernel:
// LSB-first bit reader with a one-word register prefetch, ANS (12-bit, alias table) and
// prefix-code symbol readers, hybrid-uint expansion. One *thread* owns one section stream
// (JPEG XL sections are independent, byte-aligned bit streams — SURVEY.md A.5/A.6); the
// parallelism comes from the number of sections in flight, not from inside a stream.
// Replaces libjxl's ANSSymbolReader reached from N/Decoder/JxlDecoder.cpp:252.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace jxlgpu {

// ---- flat table descriptors (host fills, device reads). Offsets are byte offsets into the job's table blob.
struct DHybrid { uint8_t split_exp, msb, lsb, pad; };
struct DCode {
  uint32_t num_ctx, num_clusters, log_alpha, use_prefix;
  uint32_t ctx_map_off;   // uint8[num_ctx]
  uint32_t cfg_off;       // DHybrid[num_clusters]
  uint32_t alias_off;     // uint64[num_clusters << log_alpha]            (ANS)
  uint32_t prefix_off;    // uint32[2*num_clusters] {lut byte offset, max_len} (prefix)
};
// alias entry: cutoff[0,13) right[13,21) off1[21,34) freq0[34,47) freq1[47,60)
__host__ __device__ inline uint64_t PackAlias(uint32_t cutoff, uint32_t right, uint32_t off1, uint32_t freq0, uint32_t freq1) {
  return uint64_t(cutoff) | (uint64_t(right) << 13) | (uint64_t(off1) << 21) | (uint64_t(freq0) << 34) | (uint64_t(freq1) << 47);
}

enum DevError : uint32_t {
  kErrNone = 0, kErrOverrun = 1, kErrAnsFinal = 2, kErrBadStrategy = 3, kErrBlockBounds = 4, kErrTooManyNz = 5, kErrNzMismatch = 6, kErrUnsupportedStream = 7,
  kErrCoefRange = 8, kErrHfMeta = 9, kErrLocalTree = 10, kErrGroupTransform = 11, kErrHybrid = 12, kErrPrefix = 13, kErrCflRange = 14, kErrSharpness = 15, kErrPreset = 16, kErrRefProps = 17,
};

#ifdef __CUDACC__
__device__ __forceinline__ void SetError(uint32_t* err, uint32_t code) { if (code) atomicCAS(err, 0u, code); }

struct BitRd {
  const uint32_t* words; uint64_t buf; int n; uint32_t nxt; uint32_t widx;   // widx: index of the next word to prefetch
  __device__ __forceinline__ void Init(const uint8_t* base16, uint64_t bitpos) {   // base16: 4-byte aligned buffer start
    words = reinterpret_cast<const uint32_t*>(base16); widx = uint32_t(bitpos >> 5); int drop = int(bitpos & 31);
    uint32_t w0 = __ldg(words + widx); uint32_t w1 = __ldg(words + widx + 1); nxt = __ldg(words + widx + 2); widx += 3;
    buf = (uint64_t(w0) | (uint64_t(w1) << 32)) >> drop; n = 64 - drop;
  }
  __device__ __forceinline__ void Refill() { if (n <= 32) { buf |= uint64_t(nxt) << n; n += 32; nxt = __ldg(words + widx); widx++; } }
  __device__ __forceinline__ uint32_t Peek(int nb) { Refill(); return uint32_t(buf) & ((nb >= 32) ? 0xffffffffu : ((1u << nb) - 1u)); }
  __device__ __forceinline__ void Skip(int nb) { buf >>= nb; n -= nb; }
  __device__ __forceinline__ uint32_t Read(int nb) { if (nb == 0) return 0; uint32_t v = Peek(nb); Skip(nb); return v; }
  __device__ __forceinline__ uint64_t BitPos() const { return uint64_t(widx - 1) * 32 - uint64_t(n); }   // absolute bit position consumed so far
  __device__ __forceinline__ uint32_t ReadU32(int n0, uint32_t o0, int n1, uint32_t o1, int n2, uint32_t o2, int n3, uint32_t o3) {
    uint32_t sel = Read(2); int nb = sel == 0 ? n0 : sel == 1 ? n1 : sel == 2 ? n2 : n3; uint32_t off = sel == 0 ? o0 : sel == 1 ? o1 : sel == 2 ? o2 : o3; return Read(nb) + off;
  }
};

struct CodeView {   // resolved pointers (generic address space: shared or global)
  const uint8_t* ctx_map; const DHybrid* cfg; const uint64_t* alias; const uint32_t* prefix_desc; const uint8_t* blob; uint32_t log_alpha, use_prefix;
  __device__ __forceinline__ void Bind(const uint8_t* blob_, const DCode& c) {
    blob = blob_; ctx_map = blob_ + c.ctx_map_off; cfg = reinterpret_cast<const DHybrid*>(blob_ + c.cfg_off); alias = reinterpret_cast<const uint64_t*>(blob_ + c.alias_off);
    prefix_desc = reinterpret_cast<const uint32_t*>(blob_ + c.prefix_off); log_alpha = c.log_alpha; use_prefix = c.use_prefix;
  }
};

struct SymReader {
  BitRd br; uint32_t state; uint32_t err;
  __device__ __forceinline__ void Init(const CodeView& cv) { err = 0; state = cv.use_prefix ? 0 : br.Read(32); }
  __device__ __forceinline__ uint32_t ReadToken(const CodeView& cv, uint32_t cluster) {
    if (cv.use_prefix) {
      uint32_t lut_off = cv.prefix_desc[2 * cluster], max_len = cv.prefix_desc[2 * cluster + 1]; const uint32_t* lut = reinterpret_cast<const uint32_t*>(cv.blob + lut_off);
      if (max_len == 0) return lut[0] >> 4;
      uint32_t e = lut[br.Peek(int(max_len))]; uint32_t len = e & 15; if (len == 0) { err = kErrPrefix; len = 1; } br.Skip(int(len)); return e >> 4;
    }
    uint32_t log_entry = 12 - cv.log_alpha; uint32_t idx = state & 0xfff, i = idx >> log_entry, pos = idx & ((1u << log_entry) - 1);
    uint64_t e = cv.alias[(cluster << cv.log_alpha) + i];
    bool g = pos >= uint32_t(e & 0x1fff); uint32_t sym = g ? uint32_t(e >> 13) & 0xff : i; uint32_t off = (g ? uint32_t(e >> 21) & 0x1fff : 0u) + pos; uint32_t freq = g ? uint32_t(e >> 47) & 0x1fff : uint32_t(e >> 34) & 0x1fff;
    state = freq * (state >> 12) + off;
    if (state < 65536u) state = (state << 16) | br.Read(16);
    return sym;
  }
  __device__ __forceinline__ uint32_t Hybrid(const DHybrid h, uint32_t t) {
    uint32_t split = 1u << h.split_exp; if (t < split) return t;
    uint32_t ml = uint32_t(h.msb) + h.lsb; uint32_t nb = h.split_exp - ml + ((t - split) >> ml);
    if (nb >= 32) { err = kErrHybrid; return 0; }
    uint32_t low = t & ((1u << h.lsb) - 1); t >>= h.lsb; uint32_t hi = (t & ((1u << h.msb) - 1)) | (1u << h.msb);
    return (((hi << nb) | br.Read(int(nb))) << h.lsb) | low;
  }
  __device__ __forceinline__ uint32_t Read(const CodeView& cv, uint32_t ctx) { uint32_t cl = cv.ctx_map[ctx]; uint32_t t = ReadToken(cv, cl); return Hybrid(cv.cfg[cl], t); }
  __device__ __forceinline__ bool FinalOk(const CodeView& cv) const { return cv.use_prefix || state == 0x130000u; }
};

__device__ __forceinline__ int32_t UnpackSignedDev(uint32_t u) { return int32_t(u >> 1) ^ -int32_t(u & 1); }
__device__ __forceinline__ int CeilLog2Dev(uint32_t x) { return x <= 1 ? 0 : 32 - __clz(x - 1); }
#endif

}  // namespace jxlgpu
