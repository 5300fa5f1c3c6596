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
  uint32_t info_off;      // uint32[num_clusters]: split_exp[0,8) | msb[8,12) | lsb[12,16) | const symbol[16,32) (0xffff: none). A constant
                          // symbol is the only symbol of a zero-entropy cluster: decoding it reads no bits and leaves the ANS state unchanged.
  // LZ77 (C.2.5): tokens >= lz_min_symbol start a copy of lz_min_length + hybrid(lz_len_info, token - lz_min_symbol) earlier values; the
  // distance is coded with cluster lz_dist_cluster (the cluster of the extra context the map carries when LZ77 is on).
  uint32_t lz77, lz_min_symbol, lz_min_length, lz_len_info, lz_dist_cluster;
};
// alias entry as two 32-bit words: x = cutoff[0,8) | right[8,16) | off1[16,32); y = freq0[0,16) | freq1[16,32)
struct DAlias { uint32_t x, y; };
__host__ __device__ inline DAlias PackAlias(uint32_t cutoff, uint32_t right, uint32_t off1, uint32_t freq0, uint32_t freq1) {
  DAlias a; a.x = cutoff | (right << 8) | (off1 << 16); a.y = freq0 | (freq1 << 16); return a;
}

enum DevError : uint32_t {
  kErrNone = 0, kErrOverrun = 1, kErrAnsFinal = 2, kErrBadStrategy = 3, kErrBlockBounds = 4, kErrTooManyNz = 5, kErrNzMismatch = 6, kErrUnsupportedStream = 7,
  kErrCoefRange = 8, kErrHfMeta = 9, kErrLocalTree = 10, kErrGroupTransform = 11, kErrHybrid = 12, kErrPrefix = 13, kErrCflRange = 14, kErrSharpness = 15, kErrPreset = 16, kErrRefProps = 17, kErrPaletteDelta = 18,
};

#ifdef __CUDACC__
__device__ __forceinline__ void SetError(uint32_t* err, uint32_t code) { if (code) atomicCAS(err, 0u, code); }

// 32-bit-lane bit reader: `lo`/`hi` hold 64 buffered bits, `used` (< 32) bits of `lo` are consumed, `nxt` is prefetched.
struct BitRd {
  const uint32_t* words; uint32_t lo, hi, nxt, used, widx;   // widx: index of the next word to prefetch
  __device__ __forceinline__ void Init(const uint8_t* base4, uint64_t bitpos) {   // base4: 4-byte aligned buffer start
    words = reinterpret_cast<const uint32_t*>(base4); widx = uint32_t(bitpos >> 5); used = uint32_t(bitpos & 31);
    lo = __ldg(words + widx); hi = __ldg(words + widx + 1); nxt = __ldg(words + widx + 2); widx += 3;
  }
  __device__ __forceinline__ uint32_t Peek32() const { return __funnelshift_r(lo, hi, used); }
  __device__ __forceinline__ uint32_t Peek(int nb) const { uint32_t v = Peek32(); return nb >= 32 ? v : (v & ((1u << nb) - 1u)); }
  __device__ __forceinline__ void Skip(int nb) { used += uint32_t(nb); if (used >= 32) { used -= 32; lo = hi; hi = nxt; nxt = __ldg(words + widx); widx++; } }   // nb <= 32
  __device__ __forceinline__ uint32_t Read(int nb) { uint32_t v = Peek(nb); Skip(nb); return v; }
  __device__ __forceinline__ uint64_t BitPos() const { return uint64_t(widx - 3) * 32 + used; }   // absolute bit position consumed so far
  __device__ __forceinline__ uint32_t ReadU32(int n0, uint32_t o0, int n1, uint32_t o1, int n2, uint32_t o2, int n3, uint32_t o3) {
    uint32_t sel = Read(2); int nb = sel == 0 ? n0 : sel == 1 ? n1 : sel == 2 ? n2 : n3; uint32_t off = sel == 0 ? o0 : sel == 1 ? o1 : sel == 2 ? o2 : o3; return Read(nb) + off;
  }
};

// Cooperative copy of a read-only table into the CTA's dynamic shared memory; when it does not fit it stays in global
// memory (generic pointers work for both). Every thread computes the same pointer; caller must __syncthreads() after.
__device__ __forceinline__ const void* StageBytes(uint8_t* dsm, uint32_t cap, uint32_t& used, const void* src, uint32_t bytes, int tid, int nt) {
  uint32_t need = (bytes + 15u) & ~15u; if (bytes == 0 || used + need > cap) return src;
  uint32_t* d = reinterpret_cast<uint32_t*>(dsm + used); const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  for (uint32_t i = tid; i < need / 4; i += nt) d[i] = s[i];
  used += need; return d;
}

struct CodeView {   // resolved pointers (generic address space: shared or global)
  const uint8_t* ctx_map; const DHybrid* cfg; const DAlias* alias; const uint32_t* prefix_desc; const uint32_t* info; const uint8_t* blob; uint32_t log_alpha, use_prefix, num_ctx, num_clusters, lz77, lz_min_symbol, lz_min_length, lz_len_info, lz_dist_cluster;
  __device__ __forceinline__ void Bind(const uint8_t* blob_, const DCode& c) {
    lz77 = c.lz77; lz_min_symbol = c.lz_min_symbol; lz_min_length = c.lz_min_length; lz_len_info = c.lz_len_info; lz_dist_cluster = c.lz_dist_cluster;
    blob = blob_; ctx_map = blob_ + c.ctx_map_off; cfg = reinterpret_cast<const DHybrid*>(blob_ + c.cfg_off); alias = reinterpret_cast<const DAlias*>(blob_ + c.alias_off);
    prefix_desc = reinterpret_cast<const uint32_t*>(blob_ + c.prefix_off); info = reinterpret_cast<const uint32_t*>(blob_ + c.info_off); log_alpha = c.log_alpha; use_prefix = c.use_prefix; num_ctx = c.num_ctx; num_clusters = c.num_clusters;
  }
  // true when every hot table sits in shared memory (lets the compiler emit LDS instead of generic loads)
  __device__ __forceinline__ bool AllShared() const { return __isShared(ctx_map) && __isShared(info) && (use_prefix || __isShared(alias)); }
  __device__ __forceinline__ void AssumeShared() const { __builtin_assume(__isShared(ctx_map)); __builtin_assume(__isShared(info)); __builtin_assume(__isShared(alias)); }
  // stage the hot tables (cluster map, hybrid configs, constant flags, alias tables) into shared memory
  __device__ __forceinline__ void Stage(uint8_t* dsm, uint32_t cap, uint32_t& used, int tid, int nt) {
    info = static_cast<const uint32_t*>(StageBytes(dsm, cap, used, info, num_clusters * 4, tid, nt));
    ctx_map = static_cast<const uint8_t*>(StageBytes(dsm, cap, used, ctx_map, num_ctx, tid, nt));
    if (!use_prefix) alias = static_cast<const DAlias*>(StageBytes(dsm, cap, used, alias, (num_clusters << log_alpha) * 8, tid, nt));
  }
};

static const uint32_t kLzWindow = 1u << 20;   // values an LZ77 copy can reach back (per stream)
struct SymReader {
  BitRd br; uint32_t state; uint32_t err;
  uint32_t* win = nullptr; uint32_t num_to_copy = 0, copy_pos = 0, num_decoded = 0;   // LZ77 state; win: kLzWindow values in global memory
  __device__ __forceinline__ void Init(const CodeView& cv) { err = 0; state = cv.use_prefix ? 0 : br.Read(32); num_to_copy = 0; copy_pos = 0; num_decoded = 0; }
  __device__ __forceinline__ uint32_t ReadToken(const CodeView& cv, uint32_t cluster) {
    if (cv.use_prefix) {
      uint32_t lut_off = cv.prefix_desc[2 * cluster], max_len = cv.prefix_desc[2 * cluster + 1]; const uint32_t* lut = reinterpret_cast<const uint32_t*>(cv.blob + lut_off);
      if (max_len == 0) return lut[0] >> 4;
      uint32_t e = lut[br.Peek(int(max_len))]; uint32_t len = e & 15; if (len == 0) { err = kErrPrefix; len = 1; } br.Skip(int(len)); return e >> 4;
    }
    const uint32_t log_entry = 12 - cv.log_alpha, idx = state & 0xfff, i = idx >> log_entry, pos = idx & ((1u << log_entry) - 1);
    const DAlias e = cv.alias[(cluster << cv.log_alpha) + i];
    const bool g = pos >= (e.x & 0xffu); const uint32_t sym = g ? ((e.x >> 8) & 0xffu) : i, off = g ? (e.x >> 16) : 0u, freq = g ? (e.y >> 16) : (e.y & 0xffffu);
    state = freq * (state >> 12) + off + pos;
    if (state < 65536u) { state = (state << 16) | br.Peek(16); br.Skip(16); }
    return sym;
  }
  __device__ __forceinline__ uint32_t ReadTokenAns(const CodeView& cv, uint32_t cluster) {
    const uint32_t log_entry = 12 - cv.log_alpha, idx = state & 0xfff, i = idx >> log_entry, pos = idx & ((1u << log_entry) - 1);
    const DAlias e = cv.alias[(cluster << cv.log_alpha) + i];
    const bool g = pos >= (e.x & 0xffu); const uint32_t sym = g ? ((e.x >> 8) & 0xffu) : i, off = g ? (e.x >> 16) : 0u, freq = g ? (e.y >> 16) : (e.y & 0xffffu);
    state = freq * (state >> 12) + off + pos;
    if (state < 65536u) { state = (state << 16) | br.Peek(16); br.Skip(16); }
    return sym;
  }
  // hybrid-uint tail for tokens >= split (info: packed per-cluster word)
  // (must inline: a non-inlined member call takes the address of the reader and forces its whole state — ANS state, bit buffer — into local memory)
  __device__ __forceinline__ uint32_t HybridSlow(uint32_t info, uint32_t t) { DHybrid h; h.split_exp = uint8_t(info & 0xff); h.msb = uint8_t((info >> 8) & 15); h.lsb = uint8_t((info >> 12) & 15); return Hybrid(h, t); }
  __device__ __forceinline__ uint32_t ReadClusterAns(const CodeView& cv, uint32_t cl) {
    const uint32_t info = cv.info[cl]; uint32_t t = info >> 16; if (t == 0xffffu) t = ReadTokenAns(cv, cl);
    if (t < (1u << (info & 0xff))) return t;
    return HybridSlow(info, t);
  }
  __device__ __forceinline__ uint32_t ReadAns(const CodeView& cv, uint32_t ctx) { return ReadClusterAns(cv, cv.ctx_map[ctx]); }
  __device__ __forceinline__ uint32_t Hybrid(const DHybrid h, uint32_t t) {
    uint32_t split = 1u << h.split_exp; if (t < split) return t;
    uint32_t ml = uint32_t(h.msb) + h.lsb; uint32_t nb = h.split_exp - ml + ((t - split) >> ml);
    if (nb >= 32) { err = kErrHybrid; return 0; }
    uint32_t low = t & ((1u << h.lsb) - 1); t >>= h.lsb; uint32_t hi = (t & ((1u << h.msb) - 1)) | (1u << h.msb);
    return (((hi << nb) | br.Read(int(nb))) << h.lsb) | low;
  }
  __device__ __forceinline__ uint32_t ReadCluster(const CodeView& cv, uint32_t cl) {
    const uint32_t info = cv.info[cl]; uint32_t t = info >> 16; if (t == 0xffffu) t = ReadToken(cv, cl);
    if (t < (1u << (info & 0xff))) return t;
    DHybrid h; h.split_exp = uint8_t(info & 0xff); h.msb = uint8_t((info >> 8) & 15); h.lsb = uint8_t((info >> 12) & 15); return Hybrid(h, t);
  }
  __device__ __forceinline__ uint32_t Read(const CodeView& cv, uint32_t ctx) { return ReadCluster(cv, cv.ctx_map[ctx]); }
  // The reader of LZ77-enabled codes. dist_mult: widest channel of a Modular sub-bitstream (distance symbols below 120 then index a table
  // of (dx, dy) offsets), 0 for the other streams.
  __device__ __noinline__ uint32_t ReadLz(const CodeView& cv, uint32_t ctx, uint32_t dist_mult) {
    const uint32_t mask = kLzWindow - 1;
    if (!win) { err = err ? err : kErrUnsupportedStream; return 0; }
    if (num_to_copy == 0) {
      const uint32_t cl = cv.ctx_map[ctx], info = cv.info[cl]; uint32_t tok = info >> 16; if (tok == 0xffffu) tok = ReadToken(cv, cl);
      if (tok < cv.lz_min_symbol) {   // literal
        uint32_t r = tok; if (tok >= (1u << (info & 0xff))) r = HybridSlow(info, tok);
        win[(num_decoded++) & mask] = r; return r;
      }
      const uint32_t lt = tok - cv.lz_min_symbol; uint32_t len = lt; if (lt >= (1u << (cv.lz_len_info & 0xff))) len = HybridSlow(cv.lz_len_info, lt);
      num_to_copy = len + cv.lz_min_length;
      const uint32_t dcl = cv.lz_dist_cluster, dinfo = cv.info[dcl]; uint32_t dtok = dinfo >> 16; if (dtok == 0xffffu) dtok = ReadToken(cv, dcl);
      uint32_t distance = dtok; if (dtok >= (1u << (dinfo & 0xff))) distance = HybridSlow(dinfo, dtok);
      if (dist_mult == 0) distance++;
      else if (distance < 120) {   // special distances: (dx, dy) pairs in a fixed order, offset = dx + width * dy, at least 1
        const int8_t kdx[120] = {0,1,1,-1,0,2,1,-1,2,-2,2,-2,0,3,1,-1,3,-3,2,-2,3,-3,0,4,1,-1,4,-4,3,-3,2,-2,4,-4,0,3,-3,4,-4,5,1,-1,5,-5,2,-2,5,-5,4,-4,3,-3,5,-5,0,6,1,-1,6,-6,2,-2,6,-6,4,-4,5,-5,3,-3,6,-6,0,7,1,-1,5,-5,7,-7,4,-4,6,-6,2,-2,7,-7,3,-3,7,-7,5,-5,6,-6,8,4,-4,7,-7,8,8,6,-6,8,5,-5,7,-7,8,6,-6,7,-7,8,7,-7,8,8};
        const int8_t kdy[120] = {1,0,1,1,2,0,2,2,1,1,2,2,3,0,3,3,1,1,3,3,2,2,4,0,4,4,1,1,3,3,4,4,2,2,5,4,4,3,3,0,5,5,1,1,5,5,2,2,4,4,5,5,3,3,6,0,6,6,1,1,6,6,2,2,5,5,4,4,6,6,3,3,7,0,7,7,5,5,1,1,6,6,4,4,7,7,2,2,7,7,3,3,6,6,5,5,0,7,7,4,4,1,2,6,6,3,7,7,5,5,4,7,7,6,6,5,7,7,6,7};
        const int off = int(kdx[distance]) + int(dist_mult) * int(kdy[distance]); distance = off < 1 ? 1u : uint32_t(off);
      } else distance -= 119;
      distance = min(distance, min(num_decoded, kLzWindow));
      copy_pos = num_decoded - distance;
      if (distance == 0) { const uint32_t n = min(num_to_copy, kLzWindow); for (uint32_t i = 0; i < n; i++) win[i] = 0; }
      if (num_to_copy < cv.lz_min_length) { err = err ? err : kErrHybrid; num_to_copy = 0; return 0; }
    }
    const uint32_t r = win[(copy_pos++) & mask]; num_to_copy--; win[(num_decoded++) & mask] = r; return r;
  }
  __device__ __forceinline__ bool FinalOk(const CodeView& cv) const { return cv.use_prefix || state == 0x130000u; }
};

__device__ __forceinline__ int32_t UnpackSignedDev(uint32_t u) { return int32_t(u >> 1) ^ -int32_t(u & 1); }
__device__ __forceinline__ int CeilLog2Dev(uint32_t x) { return x <= 1 ? 0 : 32 - __clz(x - 1); }
#endif

}  // namespace jxlgpu
