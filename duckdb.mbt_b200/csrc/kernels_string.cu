// K5 string_unpack: duckdb_string_t[16 B] -> Arrow utf8 (offsets + data) in ONE pass.
//
// Tile = 512 consecutive rows of one chunk (four tiles per 2048-row vector).  A CTA
//   1. takes a ticket (tiles are processed in ticket order => look-back always makes progress),
//   2. loads the tile's string_t into shared memory with coalesced 128-bit loads (read once),
//   3. block-scans the (validity-masked) lengths into tile-local offsets,
//   4. publishes its aggregate and resolves its exclusive base by decoupled look-back over the
//      predecessors' 64-bit status words (flag | value in one word, so no fences are needed),
//   5. writes offsets (coalesced), and
//   6. gathers the bytes row by row into a shared-memory stage laid out with the destination's
//      16-byte phase (lane i of a warp takes row i, so a warp reads one contiguous stretch of the
//      heap; inline bytes come from the shared-memory copy of string_t, pointer strings from the
//      device heap with the host pointer rebased; aligned 32-bit loads, five in flight, funnel
//      shifted to the stage's word alignment), then writes the stage with coalesced 128-bit
//      streaming stores.
//
// Replaces the reference's per-cell string_t read src/duckdb_native.c:597-603 and the two-pass
// malloc/strlen/memcpy getters :2474-2510 and :2699-2755.  DMB_STR_REF_BLOB reproduces the
// getter's NUL-terminated stream (strlen semantics) instead of Arrow offsets.

#include "dmb_common.cuh"

namespace dmb {

constexpr int kStrTileRows = 512;
constexpr int kStrTilesPerChunk = kVec / kStrTileRows;
constexpr int kStrPerThread = kStrTileRows / kThreads;  // 2 consecutive rows per thread in the scan
constexpr int kStageBytes = 16384;

constexpr uint64_t kFlagAggregate = 1ull << 62;
constexpr uint64_t kFlagPrefix = 2ull << 62;
constexpr uint64_t kValueMask = (1ull << 62) - 1ull;

// scratch layout (uint64 words): [0] ticket  [1] error flags  [2..] tile status
enum { kErrTileTooBig = 1, kErrOffsetOverflow = 2, kErrHeapRange = 4 };

struct StrSmem {
  uint4 str[kStrTileRows];           // string_t copies
  uint32_t off[kStrTileRows + 4];    // tile-local exclusive offsets; off[kStrTileRows] = tile total
  alignas(16) uint8_t stage[kStageBytes + 16];
  uint64_t warp_sum[kThreads / 32];
  uint64_t base;
  int64_t tile;
};

__device__ __forceinline__ uint64_t ld_status(const unsigned long long *p) {
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Copy len bytes from src (generic address: device heap or the shared-memory string_t copy, any
// alignment) to dst in shared memory (any alignment).  Whole destination words are written with
// one store each from two aligned source words (funnel shift), five source loads in flight; only
// the ragged ends go byte by byte.  Source words are read whole: the heap copy carries >= 16
// bytes of padding and the string_t copy is followed by other shared-memory fields.
__device__ __forceinline__ void copy_bytes(uint8_t *dst, const uint8_t *src, uint32_t len) {
  uint32_t i = 0;
  uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
  if (head > len) head = len;
  for (; i < head; ++i) dst[i] = src[i];
  const uint32_t nwords = (len - i) >> 2;
  if (nwords) {
    const uint8_t *s = src + i;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 3u) * 8u;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t)3);
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst + i);
    const uint32_t last = sh ? nwords : nwords - 1u;  // highest source word index that holds needed bytes
#pragma unroll 1
    for (uint32_t w = 0; w < nwords; w += 4) {
      const uint32_t x0 = sw[w];
      const uint32_t x1 = w + 1 <= last ? sw[w + 1] : 0u;
      const uint32_t x2 = w + 2 <= last ? sw[w + 2] : 0u;
      const uint32_t x3 = w + 3 <= last ? sw[w + 3] : 0u;
      const uint32_t x4 = w + 4 <= last ? sw[w + 4] : 0u;
      dw[w] = __funnelshift_r(x0, x1, sh);
      if (w + 1 < nwords) dw[w + 1] = __funnelshift_r(x1, x2, sh);
      if (w + 2 < nwords) dw[w + 2] = __funnelshift_r(x2, x3, sh);
      if (w + 3 < nwords) dw[w + 3] = __funnelshift_r(x3, x4, sh);
    }
    i += nwords * 4u;
  }
  for (; i < len; ++i) dst[i] = src[i];
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 8)
string_batch_kernel(dmb_string_job job, BatchView b, unsigned long long *scratch, int64_t ntiles) {
  __shared__ StrSmem sm;
  unsigned long long *status = scratch + 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) sm.tile = (int64_t)atomicAdd(scratch, 1ull);
  __syncthreads();
  const int64_t tile = sm.tile;
  if (tile >= ntiles) return;
  const int64_t c = tile / kStrTilesPerChunk;
  const int r_begin = (int)(tile % kStrTilesPerChunk) * kStrTileRows;
  const int count = (int)__ldg(b.counts + c);
  int nrows_tile = count - r_begin;
  nrows_tile = nrows_tile < 0 ? 0 : (nrows_tile > kStrTileRows ? kStrTileRows : nrows_tile);
  const dmb_vec_desc vd = job.vecs[c];
  const uint4 *in = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(job.in) + vd.data_off) + r_begin;
  const uint64_t *mask = vd.val_off < 0 ? nullptr : job.in_validity + vd.val_off;

  // 2. string_t tile -> shared memory
  for (int i = tid; i < nrows_tile; i += kThreads) sm.str[i] = ld_stream(in + i);
  __syncthreads();

  // 3. lengths (consecutive rows per thread) and block scan.  A row that contributes no bytes
  //    (NULL, empty, bad pointer) gets length 0 and is never touched again.
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

  // 6. gather bytes through the shared-memory stage.  Stage position p <-> global byte
  //    out_data[base - mis + p], so p % 16 == 0 is a 16-byte aligned global address.
  const uint32_t mis = (uint32_t)(base & 15ull);
  uint8_t *gbase = job.out_data + (base - mis);
  const uint32_t end = mis + (uint32_t)tile_total;
  for (uint32_t w0 = 0; w0 < end; w0 += kStageBytes) {
    const uint32_t w1 = w0 + kStageBytes;
#pragma unroll 1
    for (int i = tid; i < nrows_tile; i += kThreads) {  // lane -> consecutive rows
      const uint32_t s0 = mis + sm.off[i], s1 = mis + sm.off[i + 1];
      if (s1 == s0 || s1 <= w0 || s0 >= w1) continue;
      const uint4 e = sm.str[i];
      const uint8_t *src = e.x <= 12u ? reinterpret_cast<const uint8_t *>(&sm.str[i]) + 4
                                      : job.heap_dev + ((((uint64_t)e.w << 32) | (uint64_t)e.z) - job.heap_host_base);
      const uint32_t pay_end = MODE == DMB_STR_REF_BLOB ? s1 - 1u : s1;  // payload bytes [s0, pay_end)
      const uint32_t lo = s0 > w0 ? s0 : w0;
      const uint32_t hi = pay_end < w1 ? pay_end : w1;
      if (hi > lo) copy_bytes(sm.stage + (lo - w0), src + (lo - s0), hi - lo);
      if (MODE == DMB_STR_REF_BLOB && pay_end >= w0 && pay_end < w1) sm.stage[pay_end - w0] = 0;
    }
    __syncthreads();
    const uint32_t lo = w0 > mis ? w0 : mis;
    const uint32_t hi = w1 < end ? w1 : end;
    for (uint32_t v = tid; v < kStageBytes / 16; v += kThreads) {
      const uint32_t p = w0 + 16u * v;
      if (p >= hi) break;
      if (p + 16u <= lo) continue;
      if (p >= lo && p + 16u <= hi) {
        st_stream(reinterpret_cast<uint4 *>(gbase + p), *reinterpret_cast<const uint4 *>(sm.stage + 16u * v));
      } else {  // the neighbouring tiles own the other bytes of this vector
        const uint32_t q0 = p > lo ? p : lo, q1 = (p + 16u) < hi ? (p + 16u) : hi;
        for (uint32_t q = q0; q < q1; ++q) gbase[q] = sm.stage[q - w0];
      }
    }
    if (w1 < end) __syncthreads();
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
  return (size_t)(2 + kStrTilesPerChunk * (nchunks > 0 ? nchunks : 0)) * sizeof(unsigned long long);
}

extern "C" int32_t dmb_dev_string_batch(const dmb_string_job *job, const uint32_t *counts,
                                        const int64_t *row_off, int64_t nchunks, int64_t nrows,
                                        void *scratch, void *stream) {
  if (!job) { set_error("dmb_dev_string_batch: job is null"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  if (nchunks <= 0 || nrows <= 0) return 0;
  const int64_t ntiles = (int64_t)kStrTilesPerChunk * nchunks;
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
