// K5 string_unpack: duckdb_string_t[16 B] -> Arrow utf8 (offsets + data) in ONE pass.
//
// Tile = 1024 consecutive rows of one chunk (two tiles per 2048-row vector).  A CTA
//   1. takes a ticket (tiles are processed in ticket order => look-back always makes progress),
//   2. loads the tile's string_t into shared memory with coalesced 128-bit loads (read once),
//   3. block-scans the (validity-masked) lengths into tile-local offsets,
//   4. publishes its aggregate and resolves its exclusive base by decoupled look-back over the
//      predecessors' 64-bit status words (flag | value in one word, so no fences are needed),
//   5. writes offsets (coalesced), and
//   6. gathers the bytes OUTPUT-centrically: a thread owns one 16-byte aligned vector of the
//      output stream, finds the row that covers its first byte by binary search over the
//      tile-local offsets, assembles the 16 bytes from the rows it overlaps (inline bytes come
//      from the shared-memory copy of string_t, pointer strings from the device heap with the
//      host pointer rebased; aligned 32-bit loads + funnel shifts) and stores them with one
//      128-bit streaming store.  Every load of a thread is independent of the others, consecutive
//      threads read consecutive heap bytes and write consecutive vectors, and there is no
//      shared-memory staging of the data bytes.
//
// Replaces the reference's per-cell string_t read src/duckdb_native.c:597-603 and the two-pass
// malloc/strlen/memcpy getters :2474-2510 and :2699-2755.  DMB_STR_REF_BLOB reproduces the
// getter's NUL-terminated stream (strlen semantics) instead of Arrow offsets.

#include "dmb_common.cuh"

namespace dmb {

constexpr int kStrTileRows = 1024;
constexpr int kStrPerThread = kStrTileRows / kThreads;  // 4 consecutive rows per thread

constexpr uint64_t kFlagAggregate = 1ull << 62;
constexpr uint64_t kFlagPrefix = 2ull << 62;
constexpr uint64_t kValueMask = (1ull << 62) - 1ull;

// scratch layout (uint64 words): [0] ticket  [1] error flags  [2..] tile status
enum { kErrTileTooBig = 1, kErrOffsetOverflow = 2, kErrHeapRange = 4 };

struct StrSmem {
  uint4 str[kStrTileRows];           // string_t copies
  uint32_t off[kStrTileRows + 1];    // tile-local exclusive offsets; off[nrows] = tile total
  uint64_t warp_sum[kThreads / 32];
  uint64_t base;
  int64_t tile;
};

__device__ __forceinline__ uint64_t ld_status(const unsigned long long *p) {
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// 4 bytes starting at byte offset o (0..11) of the 12 inline bytes (y,z,w) of a string_t
__device__ __forceinline__ uint32_t inline_bytes4(const uint4 &e, uint32_t o) {
  const uint32_t wi = o >> 2, sh = (o & 3u) * 8u;
  const uint32_t x = wi == 0 ? e.y : (wi == 1 ? e.z : e.w);
  const uint32_t y = wi == 0 ? e.z : (wi == 1 ? e.w : 0u);
  return __funnelshift_r(x, y, sh);
}

// 4 bytes starting at an arbitrary global address; only the low `nb` bytes are needed.  Reads
// whole aligned words (the heap copy carries >= 16 bytes of padding past its end).
__device__ __forceinline__ uint32_t global_bytes4(const uint8_t *p, uint32_t nb) {
  const uint32_t *a = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
  const uint32_t x = __ldg(a);
  const uint32_t y = (sh + 8u * nb > 32u) ? __ldg(a + 1) : 0u;
  return __funnelshift_r(x, y, sh);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
string_batch_kernel(dmb_string_job job, BatchView b, unsigned long long *scratch, int64_t ntiles) {
  __shared__ StrSmem sm;
  unsigned long long *status = scratch + 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) sm.tile = (int64_t)atomicAdd(scratch, 1ull);
  __syncthreads();
  const int64_t tile = sm.tile;
  if (tile >= ntiles) return;
  const int64_t c = tile >> 1;
  const int r_begin = (int)(tile & 1) * kStrTileRows;
  const int count = (int)__ldg(b.counts + c);
  int nrows_tile = count - r_begin;
  nrows_tile = nrows_tile < 0 ? 0 : (nrows_tile > kStrTileRows ? kStrTileRows : nrows_tile);
  const dmb_vec_desc vd = job.vecs[c];
  const uint4 *in = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(job.in) + vd.data_off) + r_begin;
  const uint64_t *mask = vd.val_off < 0 ? nullptr : job.in_validity + vd.val_off;

  // 2. string_t tile -> shared memory
  for (int i = tid; i < nrows_tile; i += kThreads) sm.str[i] = ld_stream(in + i);
  __syncthreads();

  // 3. lengths (4 consecutive rows per thread) and block scan
  uint32_t len[kStrPerThread];
  uint64_t tsum = 0;
  bool bad_heap = false;
#pragma unroll
  for (int k = 0; k < kStrPerThread; ++k) {
    const int i = tid * kStrPerThread + k;
    len[k] = 0;
    if (i < nrows_tile) {
      const int row = r_begin + i;
      const bool valid = mask ? ((__ldg(mask + (row >> 6)) >> (row & 63)) & 1ull) : true;
      if (valid) {
        const uint4 e = sm.str[i];
        uint32_t l = e.x;
        const uint8_t *src = reinterpret_cast<const uint8_t *>(&sm.str[i]) + 4;
        if (l > 12u) {
          const uint64_t p = ((uint64_t)e.w << 32) | (uint64_t)e.z;
          const uint64_t rel = p - job.heap_host_base;
          if (p < job.heap_host_base || rel + l > job.heap_len) { bad_heap = true; l = 0; }
          src = job.heap_dev + rel;
        }
        if (MODE == DMB_STR_REF_BLOB) {  // strlen() of the malloc'ed copy: stop at an embedded NUL
          uint32_t n = 0;
          while (n < l && src[n] != 0) ++n;
          l = n;
        }
        len[k] = l;
      }
      if (MODE == DMB_STR_REF_BLOB) len[k] += 1;  // terminator; a NULL row is a lone '\0'
    }
    tsum += len[k];
  }
  uint64_t incl = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t n = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += n;
  }
  if (lane == 31) sm.warp_sum[warp] = incl;
  __syncthreads();
  uint64_t warp_excl = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    uint64_t s = sm.warp_sum[w];
    if (w < warp) warp_excl += s;
    tile_total += s;
  }
  const uint64_t excl = warp_excl + incl - tsum;
  const bool too_big = tile_total > 0xfffffff0ull;
  if (too_big && tid == 0) atomicOr(scratch + 1, (unsigned long long)kErrTileTooBig);
  if (bad_heap) atomicOr(scratch + 1, (unsigned long long)kErrHeapRange);
  {
    uint32_t o = (uint32_t)excl;
#pragma unroll
    for (int k = 0; k < kStrPerThread; ++k) {
      sm.off[tid * kStrPerThread + k] = o;
      o += len[k];
    }
    if (tid == kThreads - 1) sm.off[kStrTileRows] = o;
  }

  // 4. decoupled look-back (warp 0)
  if (warp == 0) {
    uint64_t agg = tile_total & kValueMask;
    if (lane == 0) atomicExch(status + tile, (tile == 0 ? kFlagPrefix : kFlagAggregate) | agg);
    uint64_t prefix = 0;
    if (tile > 0) {
      int64_t look = tile - 1;
      while (true) {
        const int64_t idx = look - lane;
        uint64_t st = kFlagPrefix;  // before tile 0: prefix 0
        if (idx >= 0) {
          do { st = ld_status(status + idx); } while ((st >> 62) == 0);
        }
        const uint32_t is_p = __ballot_sync(0xffffffffu, (st >> 62) == 2);
        const int first_p = is_p ? (__ffs(is_p) - 1) : 32;
        uint64_t v = lane <= first_p ? (st & kValueMask) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        prefix += v;
        if (is_p) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(status + tile, kFlagPrefix | ((prefix + agg) & kValueMask));
    }
    if (lane == 0) sm.base = prefix;
  }
  __syncthreads();
  const uint64_t base = sm.base;

  // 5. offsets
  {
    const int64_t out_row0 = __ldg(b.row_off + c) + r_begin;
    const bool overflow = MODE != DMB_STR_ARROW_LARGE && base + tile_total > 0x7fffffffull;
    if (overflow && tid == 0) atomicOr(scratch + 1, (unsigned long long)kErrOffsetOverflow);
    if (MODE == DMB_STR_ARROW_LARGE) {
      int64_t *oo = reinterpret_cast<int64_t *>(job.out_offsets);
      for (int i = tid; i < nrows_tile; i += kThreads) __stcs(reinterpret_cast<long long *>(oo + out_row0 + i), (long long)(base + sm.off[i]));
      if (tile == ntiles - 1 && tid == 0) oo[b.nrows] = (int64_t)(base + tile_total);
    } else {
      int32_t *oo = reinterpret_cast<int32_t *>(job.out_offsets);
      for (int i = tid; i < nrows_tile; i += kThreads) __stcs(oo + out_row0 + i, (int32_t)(base + sm.off[i]));
      if (tile == ntiles - 1 && tid == 0) oo[b.nrows] = (int32_t)(base + tile_total);
    }
    if (tile == ntiles - 1 && tid == 0 && job.total_bytes) *job.total_bytes = base + tile_total;
  }
  if (tile_total == 0 || too_big) return;

  // 6. output-centric gather.  Vector v covers global bytes [16v - mis, 16v - mis + 16) relative
  //    to the tile's first output byte, so every vector is 16-byte aligned in out_data.
  const uint32_t total = (uint32_t)tile_total;
  const uint32_t mis = (uint32_t)(base & 15ull);
  uint8_t *gbase = job.out_data + (base - mis);
  const uint32_t nvec = (mis + total + 15u) >> 4;
  // rows >= nrows_tile have len 0, so off[] is non-decreasing over the whole [0, 1024] range
  for (uint32_t v = tid; v < nvec; v += kThreads) {
    const uint32_t vbeg = v << 4;                          // position + mis of the vector's first byte
    const uint32_t lo = v ? vbeg - mis : 0u;               // tile-local byte range [lo, hi) owned by this vector
    const uint32_t hi = (vbeg + 16u - mis) < total ? (vbeg + 16u - mis) : total;
    int l = 0, h = kStrTileRows;                           // invariant: off[l] <= lo < off[h]
#pragma unroll 1
    while (h - l > 1) {
      const int m = (l + h) >> 1;
      if (sm.off[m] <= lo) l = m; else h = m;
    }
    int r = l;
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    uint32_t pos = lo;
    uint32_t r_off = sm.off[r], r_end = sm.off[r + 1];
#pragma unroll 1
    while (pos < hi) {
      if (r_end <= pos) {  // empty (or exhausted) row: next
        ++r;
        r_off = r_end;
        r_end = sm.off[r + 1];
        continue;
      }
      const uint32_t pay_end = MODE == DMB_STR_REF_BLOB ? r_end - 1u : r_end;  // the terminator byte stays 0
      const uint32_t seg_end = pay_end < hi ? pay_end : hi;
      if (seg_end > pos) {
        const uint4 e = sm.str[r];
        const uint32_t soff = pos - r_off;                 // offset inside the string
        const uint32_t d = pos + mis - vbeg;               // destination byte inside the vector
        const uint32_t k = seg_end - pos;                  // bytes to place (1..16)
        const bool is_inline = e.x <= 12u;
        const uint8_t *gsrc = job.heap_dev + ((((uint64_t)e.w << 32) | (uint64_t)e.z) - job.heap_host_base) + soff;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t a = d > 4u * j ? d : 4u * j;
          const uint32_t bb = (d + k) < (4u * j + 4u) ? (d + k) : (4u * j + 4u);
          if (bb > a) {
            const uint32_t nb = bb - a;
            const uint32_t val = is_inline ? inline_bytes4(e, soff + (a - d)) : global_bytes4(gsrc + (a - d), nb);
            const uint32_t m = nb == 4u ? 0xffffffffu : ((1u << (8u * nb)) - 1u);
            const uint32_t piece = (val & m) << (8u * (a - 4u * j));
            if (j == 0) w0 |= piece; else if (j == 1) w1 |= piece; else if (j == 2) w2 |= piece; else w3 |= piece;
          }
        }
      }
      pos = (MODE == DMB_STR_REF_BLOB && seg_end == pay_end && pay_end < hi) ? r_end : seg_end;
    }
    const bool full = (v > 0 || mis == 0) && (vbeg + 16u - mis <= total);
    if (full) {
      st_stream(reinterpret_cast<uint4 *>(gbase + vbeg), make_uint4(w0, w1, w2, w3));
    } else {  // the neighbouring tiles own the other bytes of this vector
      const uint32_t q0 = lo + mis - vbeg, q1 = hi + mis - vbeg;
      for (uint32_t q = q0; q < q1; ++q) {
        const uint32_t word = q < 4 ? w0 : (q < 8 ? w1 : (q < 12 ? w2 : w3));
        gbase[vbeg + q] = (uint8_t)(word >> (8u * (q & 3u)));
      }
    }
  }
}

// bench/test helper: DuckDB-shaped string_t from lengths + heap offsets
__global__ void __launch_bounds__(kThreads)
make_string_t_kernel(const uint32_t *__restrict__ lengths, const uint64_t *__restrict__ heap_off,
                     const uint8_t *__restrict__ heap_dev, uint64_t heap_host_base,
                     dmb_string_t *__restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    uint32_t l = lengths[i];
    uint64_t ho = heap_off[i];
    uint4 e = make_uint4(l, 0, 0, 0);
    uint8_t *eb = reinterpret_cast<uint8_t *>(&e);
    if (l <= 12u) {
      for (uint32_t k = 0; k < l; ++k) eb[4 + k] = heap_dev[ho + k];
    } else {
      for (uint32_t k = 0; k < 4; ++k) eb[4 + k] = heap_dev[ho + k];
      uint64_t p = heap_host_base + ho;
      e.z = (uint32_t)p;
      e.w = (uint32_t)(p >> 32);
    }
    reinterpret_cast<uint4 *>(out)[i] = e;
  }
}

}  // namespace dmb

using namespace dmb;

extern "C" size_t dmb_dev_string_scratch_bytes(int64_t nchunks) {
  return (size_t)(2 + 2 * (nchunks > 0 ? nchunks : 0)) * sizeof(unsigned long long);
}

extern "C" int32_t dmb_dev_string_batch(const dmb_string_job *job, const uint32_t *counts,
                                        const int64_t *row_off, int64_t nchunks, int64_t nrows,
                                        void *scratch, void *stream) {
  if (!job) { set_error("dmb_dev_string_batch: job is null"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  if (nchunks <= 0 || nrows <= 0) return 0;
  const int64_t ntiles = 2 * nchunks;
  if (check_cuda(cudaMemsetAsync(scratch, 0, dmb_dev_string_scratch_bytes(nchunks), st), "string scratch memset")) return -1;
  BatchView b{counts, row_off, nchunks, nrows};
  auto launch = [&](auto kernel) -> int32_t {
    kernel<<<(unsigned)ntiles, kThreads, 0, st>>>(*job, b, (unsigned long long *)scratch, ntiles);
    return check_cuda(cudaGetLastError(), "string_batch_kernel launch");
  };
  switch (job->mode) {
    case DMB_STR_ARROW_UTF8: return launch(string_batch_kernel<DMB_STR_ARROW_UTF8>);
    case DMB_STR_ARROW_LARGE: return launch(string_batch_kernel<DMB_STR_ARROW_LARGE>);
    case DMB_STR_REF_BLOB: return launch(string_batch_kernel<DMB_STR_REF_BLOB>);
    default: set_error("dmb_dev_string_batch: bad mode %d", job->mode); return -1;
  }
}

// error flags of the last string launch that used `scratch` (host reads after a sync)
extern "C" int32_t dmb_dev_string_error(const void *scratch, void *stream) {
  unsigned long long flags = 0;
  if (check_cuda(cudaMemcpyAsync(&flags, (const unsigned long long *)scratch + 1, sizeof(flags), cudaMemcpyDeviceToHost, (cudaStream_t)stream), "string error copy")) return -1;
  if (check_cuda(cudaStreamSynchronize((cudaStream_t)stream), "string error sync")) return -1;
  if (flags & kErrHeapRange) set_error("string_t pointer outside the registered heap");
  else if (flags & kErrOffsetOverflow) set_error("utf8 data exceeds int32 offsets; use large offsets or smaller batches");
  else if (flags & kErrTileTooBig) set_error("a 1024-row tile holds more than 4 GiB of string bytes");
  return (int32_t)flags;
}

extern "C" int32_t dmb_dev_make_string_t(const uint32_t *lengths, const uint64_t *heap_off,
                                         const uint8_t *heap_dev, uint64_t heap_host_base,
                                         dmb_string_t *out, int64_t n, void *stream) {
  if (n <= 0) return 0;
  int64_t blocks = (n + kThreads - 1) / kThreads;
  int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  make_string_t_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(lengths, heap_off, heap_dev, heap_host_base, out, n);
  return check_cuda(cudaGetLastError(), "make_string_t_kernel launch");
}
