// Shared device/host helpers for libduckdb_mb_gpu (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "duckdb_mb_gpu.h"

namespace dmb {

constexpr int kThreads = 256;          // CTA size of every streaming kernel
constexpr int kVec = DMB_VECTOR_SIZE;  // 2048 rows per DuckDB vector / output tile
constexpr int kNumSMs = 148;           // B200

void set_error(const char *fmt, ...);
int32_t check_cuda(cudaError_t e, const char *what);

// ---- 16-byte payload types -------------------------------------------------------------
struct alignas(16) i128 {
  uint64_t lo;
  int64_t hi;
};
struct alignas(16) u128 {
  uint64_t lo;
  uint64_t hi;
};
struct alignas(16) interval_t {  // duckdb_interval
  int32_t months;
  int32_t days;
  int64_t micros;
};
struct alignas(16) month_day_nano_t {  // Arrow month_day_nano_interval
  int32_t months;
  int32_t days;
  int64_t nanos;
};

// R consecutive values moved as one naturally aligned vector (1..16 bytes)
template <typename T, int R>
struct alignas(sizeof(T) * R) Pack {
  T v[R];
};

// ---- streaming (evict-first) vector loads / stores: every byte is touched exactly once ----
template <int N>
struct RawVec;
template <>
struct RawVec<1> { using type = unsigned char; };
template <>
struct RawVec<2> { using type = unsigned short; };
template <>
struct RawVec<4> { using type = unsigned int; };
template <>
struct RawVec<8> { using type = uint2; };
template <>
struct RawVec<16> { using type = uint4; };

template <typename P>
__device__ __forceinline__ P ld_stream(const P *p) {
  using R = typename RawVec<sizeof(P)>::type;
  R r = __ldcs(reinterpret_cast<const R *>(p));
  P out;
  memcpy(&out, &r, sizeof(P));
  return out;
}
template <typename P>
__device__ __forceinline__ void st_stream(P *p, const P &v) {
  using R = typename RawVec<sizeof(P)>::type;
  R r;
  memcpy(&r, &v, sizeof(P));
  __stcs(reinterpret_cast<R *>(p), r);
}

// ---- chunk geometry ------------------------------------------------------------------------
struct BatchView {
  const uint32_t *counts;  // [nchunks]
  const int64_t *row_off;  // [nchunks+1]
  int64_t nchunks;
  int64_t nrows;
};

// largest c in [0, nchunks) with row_off[c] <= row; for row < nrows that chunk is non-empty
__device__ __forceinline__ int64_t find_chunk(const BatchView &b, int64_t row) {
  int64_t lo = 0, hi = b.nchunks;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(b.row_off + mid) <= row) lo = mid; else hi = mid;
  }
  return lo;
}

// `take` (1..64) validity bits of chunk rows [local, local+take); bit i = row local+i valid
__device__ __forceinline__ uint64_t chunk_valid_bits(const uint64_t *vslab, const dmb_vec_desc *vecs,
                                                     int64_t c, int local, int take) {
  uint64_t keep = take >= 64 ? ~0ull : ((1ull << take) - 1ull);
  int64_t vo = vecs[c].val_off;
  if (vo < 0) return keep;  // NULL validity pointer = all valid (reference: duckdb_native.c:531-533)
  const uint64_t *m = vslab + vo;
  int w = local >> 6, s = local & 63;
  uint64_t bits = __ldg(m + w) >> s;
  if (s && s + take > 64) bits |= __ldg(m + w + 1) << (64 - s);
  return bits & keep;
}

// 64 validity bits of output rows [r0, r0+64) gathered across chunk boundaries (rows >= nrows: 0)
__device__ __forceinline__ uint64_t gather_valid64(const BatchView &b, const uint64_t *vslab,
                                                   const dmb_vec_desc *vecs, int64_t r0) {
  if (r0 >= b.nrows) return 0ull;
  int64_t c = find_chunk(b, r0);
  uint64_t word = 0;
  int filled = 0;
  while (filled < 64 && r0 + filled < b.nrows) {
    int64_t local = r0 + filled - __ldg(b.row_off + c);
    int64_t avail = (int64_t)__ldg(b.counts + c) - local;
    if (avail <= 0) { ++c; continue; }
    int take = (int)(avail < (int64_t)(64 - filled) ? avail : (int64_t)(64 - filled));
    word |= chunk_valid_bits(vslab, vecs, c, (int)local, take) << filled;
    filled += take;
  }
  return word;
}

// tile t (output rows [t*2048, t*2048+2048)) is exactly chunk t, rows [0, count)
__device__ __forceinline__ bool tile_is_regular(const BatchView &b, int64_t t) {
  if (t >= b.nchunks) return false;
  int64_t ro = __ldg(b.row_off + t);
  if (ro != t * (int64_t)kVec) return false;
  uint32_t cnt = __ldg(b.counts + t);
  return cnt == (uint32_t)kVec || ro + cnt == b.nrows;
}

// 8 bytes that are each 0/1 -> 8 bits (byte k -> bit k)
__device__ __forceinline__ uint32_t pack8(uint64_t bytes01) {
  return (uint32_t)((bytes01 * 0x0102040810204080ull) >> 56);
}
// 8 bits -> 8 bytes that are each 0/1
__device__ __forceinline__ uint64_t spread8(uint32_t bits) {
  uint64_t x = (uint64_t)(bits & 0xff) * 0x0101010101010101ull;  // byte m = bits
  x &= 0x8040201008040201ull;                                      // byte m keeps bit m
  x += 0x7f7f7f7f7f7f7f7full;                                      // bit 7 of byte m = (bit m != 0)
  return (x >> 7) & 0x0101010101010101ull;
}

// ---- bitmaps as aligned 32-bit words, misaligned 16-byte vectors (reverse kernels, LIST child copy) -----------------
// An LSB bitmap read as aligned 32-bit words: `w` is the bitmap address rounded down to 4 bytes and
// `base` the bit position of row 0 counted from there (Arrow array offset + the rounded-off bytes).
struct BitSrc {
  const uint32_t *w;
  int64_t base;
};
__device__ __forceinline__ BitSrc bit_src(const uint8_t *bm, int64_t bit_offset) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(bm);
  return BitSrc{reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3), bit_offset + (int64_t)(a & 3u) * 8};
}

// `take` (1..32) bits from bit position p; only the words that hold a requested bit are touched
__device__ __forceinline__ uint32_t load_bits32(const uint32_t *w, int64_t p, int take) {
  const int64_t wi = p >> 5;
  const int s = (int)(p & 31);
  const uint32_t lo = __ldg(w + wi);
  const uint32_t hi = (s + take > 32) ? __ldg(w + wi + 1) : 0u;
  const uint32_t r = __funnelshift_r(lo, hi, (uint32_t)s);
  return take >= 32 ? r : (r & ((1u << take) - 1u));
}

// words [ws, ws+4] of the 8 words (a, b), shifted right by sh bits: the 16 bytes that start m = 4*ws + sh/8
// bytes into a
__device__ __forceinline__ uint4 shift_words(const uint4 &a, const uint4 &b, int ws, uint32_t sh) {
  uint32_t w0, w1, w2, w3, w4;
  switch (ws) {
    case 0: w0 = a.x; w1 = a.y; w2 = a.z; w3 = a.w; w4 = b.x; break;
    case 1: w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; break;
    case 2: w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; break;
    default: w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; break;
  }
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}

// keep-mask of 32-bit word i of a 16-byte vector of 16/W elements, bit r of `bits` = element r valid
template <int W>
__device__ __forceinline__ uint32_t word_keep(uint32_t bits, int i) {
  if (W >= 4) return ((bits >> (i * 4 / W)) & 1u) ? 0xffffffffu : 0u;
  if (W == 2) {
    const uint32_t b = bits >> (2 * i);
    return ((b & 1u) ? 0x0000ffffu : 0u) | ((b & 2u) ? 0xffff0000u : 0u);
  }
  const uint32_t b = bits >> (4 * i);
  return ((b & 1u) | ((b & 2u) << 7) | ((b & 4u) << 14) | ((b & 8u) << 21)) * 0xffu;
}

// the string_t of dictionary entry idx (<= 12 bytes inlined, else prefix + pointer relative to dict_host_base)
__device__ __forceinline__ uint4 enum_entry(const dmb_enum_job &job, uint32_t idx) {
  const uint32_t o0 = __ldg(job.dict_offsets + idx), o1 = __ldg(job.dict_offsets + idx + 1);
  const uint32_t len = o1 - o0;
  const uint8_t *q = job.dict_data + o0;
  uint4 e = make_uint4(len, 0, 0, 0);
  if (len <= 12u) {
    uint32_t w[3] = {0u, 0u, 0u};
    for (uint32_t k = 0; k < len; ++k) w[k >> 2] |= (uint32_t)__ldg(q + k) << (8u * (k & 3u));
    e.y = w[0]; e.z = w[1]; e.w = w[2];
  } else {
    e.y = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16) | ((uint32_t)__ldg(q + 3) << 24);
    const uint64_t p = job.dict_host_base + o0;
    e.z = (uint32_t)p;
    e.w = (uint32_t)(p >> 32);
  }
  return e;
}

}  // namespace dmb
