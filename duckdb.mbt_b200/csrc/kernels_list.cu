// K9 list_to_arrow: DuckDB LIST vectors -> Arrow list<child> ("nested offsets kernels").
//
// DuckDB side, per chunk: a vector of duckdb_list_entry {uint64 offset, uint64 length} (what
// duckdb_vector_get_data returns for a LIST vector) + its validity mask, and ONE child vector per chunk
// (duckdb_list_vector_get_child / _get_size) that the entries index.  Entries need not be in row order or
// contiguous (slices, selections), and the entry of a NULL row is unspecified.
// Arrow side: offsets[n+1] (int32, or int64 for large_list) = running sum of the valid rows' lengths, the
// child values gathered in row order, the child validity bitmap gathered with them (payload under a NULL
// child element zeroed, like every other export of this library).
//
// The reference rejects LIST on its chunk path (src/duckdb_native.c:271-303) and its README lists
// List/Struct/Map as "not yet supported" for Arrow: there is no reference output to match; the contract is
// the Arrow format (validated with pyarrow) and the oracle's restatement of the loops above.
//
// One launch (list_emit_kernel, ONEPASS: the chunk base by a decoupled look-back over per-chunk status words); the three-launch
// form below (sum, scan, emit with precomputed bases) stays behind DMB_LIST_THREE_PASS for A/B measurements:
//   list_sum_kernel    one CTA per chunk (grid-stride): sum of the valid rows' lengths -> chunk_sum[k]
//   list_scan_kernel   one CTA: exclusive scan of chunk_sum -> chunk_base[k], total
//   list_emit_kernel   one CTA per chunk: entries striped into shared memory, block scan of the lengths -> offsets; then the
//                      chunk's child elements.  A chunk whose entries are one run in row order (what a scan produces;
//                      detected from entry.offset - start being the same for every non-empty row) is copied as aligned
//                      16-byte output vectors (source misaligned by whole elements: two aligned loads + funnel shift),
//                      NULL elements zeroed from the mask bits, the child bitmap as a shifted word copy.  Any other chunk
//                      is gathered output-centric, one element per lane: row of an output element by binary search over
//                      the chunk's 2048 row starts in shared memory, child validity by warp ballot over 32-aligned groups
//                      of OUTPUT elements.  Whole bitmap words are stored; the ragged first / last word of a chunk is
//                      merged with atomicOr into the pre-zeroed bitmap.

#include <stdlib.h>

#include "dmb_common.cuh"

namespace dmb {

constexpr int kListRpt = kVec / kThreads;  // rows per thread: 8

struct ListEntry {
  uint64_t offset, length;
};

__device__ __forceinline__ bool list_row_valid(const uint64_t *mask, int i) {
  return mask ? ((__ldg(mask + (i >> 6)) >> (i & 63)) & 1ull) : true;
}

// block-wide exclusive scan of one value per thread (kThreads threads); returns the exclusive prefix, *total = sum
__device__ __forceinline__ uint64_t block_exscan(uint64_t v, uint64_t *total, uint64_t *s_warp /* [kThreads / 32 + 1] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < kThreads / 32 ? s_warp[lane] : 0ull;
    uint64_t winc = w;
#pragma unroll
    for (int d = 1; d < kThreads / 32; d <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += t;
    }
    if (lane < kThreads / 32) s_warp[lane] = winc - w;
    if (lane == kThreads / 32 - 1) s_warp[kThreads / 32] = winc;
  }
  __syncthreads();
  const uint64_t out = s_warp[warp] + inc - v;
  *total = s_warp[kThreads / 32];
  __syncthreads();
  return out;
}

__global__ void __launch_bounds__(kThreads)
list_sum_kernel(dmb_list_job job, const uint32_t *__restrict__ counts, int64_t nchunks, unsigned long long *chunk_sum) {
  __shared__ uint64_t s_warp[kThreads / 32 + 1];
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int count = (int)__ldg(counts + c);
    const dmb_vec_desc vd = job.vecs[c];
    const ListEntry *ent = reinterpret_cast<const ListEntry *>(reinterpret_cast<const uint8_t *>(job.in_entries) + vd.data_off);
    const uint64_t *mask = vd.val_off < 0 ? nullptr : job.in_validity + vd.val_off;
    uint64_t sum = 0;
    const uint64_t csize = job.child_sizes ? __ldg(job.child_sizes + c) : ~0ull;
    for (int i = threadIdx.x; i < count; i += kThreads) {
      const ulonglong2 e = __ldg(reinterpret_cast<const ulonglong2 *>(ent) + i);  // consecutive lanes, consecutive 16-byte entries
      if (list_row_valid(mask, i) && e.x <= csize && e.y <= csize - e.x) sum += e.y;  // (an entry outside the child vector counts as empty, see list_emit_kernel)
    }
    uint64_t total;
    block_exscan(sum, &total, s_warp);
    if (threadIdx.x == 0) chunk_sum[c] = total;
  }
}

// chunk_sum -> chunk_base (exclusive); one CTA of 1024 threads, thread t owns the consecutive chunks [t*m, t*m + m)
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads)
list_scan_kernel(const unsigned long long *chunk_sum, unsigned long long *chunk_base, int64_t nchunks, unsigned long long *total_out,
                 unsigned long long *flags, int large) {
  __shared__ uint64_t s_warp[kScanThreads / 32 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m = (nchunks + kScanThreads - 1) / kScanThreads;
  const int64_t c0 = (int64_t)threadIdx.x * m, c1 = c0 + m < nchunks ? c0 + m : nchunks;
  uint64_t mine = 0;
  for (int64_t c = c0; c < c1; ++c) mine += chunk_sum[c];
  uint64_t inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint64_t w = s_warp[lane];
    uint64_t winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += t;
    }
    s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  uint64_t run = s_warp[warp] + inc - mine;
  for (int64_t c = c0; c < c1; ++c) {
    chunk_base[c] = run;
    run += chunk_sum[c];
  }
  if (threadIdx.x == 0) {
    const uint64_t total = s_warp[32];
    if (total_out) *total_out = total;
    if (!large && total > 0x7fffffffull && flags) atomicOr(flags, 1ull);  // int32 offsets overflow: use large_list
  }
}

template <int W>
__device__ __forceinline__ typename RawVec<W>::type load_elem(const uint8_t *src, bool valid) {
  using T = typename RawVec<W>::type;
  T v;
  if (valid) v = *reinterpret_cast<const T *>(src); else memset(&v, 0, sizeof(T));
  return v;
}

// ---- one-pass variant: the chunk bases come from a decoupled look-back over per-chunk status words instead of the
// sum + scan launches (the entries are then read once).  status[c]: bits 63..62 = 1 (aggregate) / 2 (inclusive prefix),
// low 62 bits the value.  The grid is persistent and no larger than what is resident at once, CTA b takes chunks b,
// b + grid, ...: every predecessor a look-back waits for belongs to a resident CTA that waits only on earlier chunks.
constexpr unsigned long long kStatusMask = (1ull << 62) - 1ull;
#ifndef DMB_LIST_GROUPS
#define DMB_LIST_GROUPS 0  // the two-level look-back of the string kernels: built and measured, 0.266 vs 0.255 ms per 20 M rows here (the
                           // chunk-by-chunk walk is not what bounds this kernel, the extra atomic per chunk shows), so it stays off
#endif
__device__ __forceinline__ unsigned long long list_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// exclusive prefix of chunk c, executed by one warp (lane 0 looks at c-1, lane 1 at c-2, ...)
__device__ __forceinline__ uint64_t list_lookback(const unsigned long long *status, int64_t c, int lane, unsigned long long *flags) {
  uint64_t prefix = 0;
  int64_t pos = c - 1;
  const unsigned long long t0 = list_now_ns();
  while (pos >= 0) {
    const int64_t idx = pos - lane;
    const unsigned long long w = idx >= 0 ? *reinterpret_cast<const volatile unsigned long long *>(status + idx) : (2ull << 62);  // before chunk 0: prefix 0
    const uint32_t flag = (uint32_t)(w >> 62);
    const uint32_t is_prefix = __ballot_sync(0xffffffffu, flag == 2u), is_empty = __ballot_sync(0xffffffffu, flag == 0u);
    const int first_empty = is_empty ? __ffs((int)is_empty) - 1 : 32;
    const int first_prefix = is_prefix ? __ffs((int)is_prefix) - 1 : 32;
    const bool done = first_prefix < first_empty;
    const int upto = done ? first_prefix + 1 : first_empty;  // lanes [0, upto) hold published words
    uint64_t v = lane < upto ? (w & kStatusMask) : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    prefix += v;
    if (done) break;
    pos -= upto;
    if (upto == 0) {
      if (list_now_ns() - t0 > 2000000000ull) {  // never hang the GPU: report and leave
        if (lane == 0) atomicOr(flags, 4ull);
        break;
      }
      __nanosleep(64);
    }
  }
  return prefix;
}

template <int W, bool LARGE, bool ONEPASS>
__global__ void __launch_bounds__(kThreads)
list_emit_kernel(dmb_list_job job, BatchView b, unsigned long long *chunk_sum /* ONEPASS: the status words */,
                 const unsigned long long *__restrict__ chunk_base, unsigned long long *flags) {
  __shared__ unsigned long long s_cbase;
  __shared__ uint64_t s_warp[kThreads / 32 + 1];
  __shared__ uint32_t s_start[kVec + 1];  // row -> first output element of the row, relative to the chunk
  __shared__ uint64_t s_src[kVec];        // row -> entry.offset
  __shared__ unsigned long long s_wmin[kThreads / 32], s_wmax[kThreads / 32];
  __shared__ unsigned s_nulls;
  const int lane = threadIdx.x & 31;
  for (int64_t c = blockIdx.x; c < b.nchunks; c += gridDim.x) {
    const int count = (int)__ldg(b.counts + c);
    const int64_t row0 = __ldg(b.row_off + c);
    const dmb_vec_desc vd = job.vecs[c];
    const ListEntry *ent = reinterpret_cast<const ListEntry *>(reinterpret_cast<const uint8_t *>(job.in_entries) + vd.data_off);
    const uint64_t *mask = vd.val_off < 0 ? nullptr : job.in_validity + vd.val_off;
    uint64_t cbase = 0, csum = 0;
    if (!ONEPASS) {
      cbase = chunk_base[c];
      csum = chunk_sum[c];
    }
    if (threadIdx.x == 0) s_nulls = 0u;
    if (!ONEPASS && csum > 0xffffffffull) {  // one chunk with more than 4 G child elements (uniform branch)
      if (threadIdx.x == 0) atomicOr(flags, 2ull);
      continue;
    }
    // ---- entries into shared memory, striped (consecutive lanes, consecutive 16-byte entries); NULL rows: length 0
    {
      // all of a thread's loads are issued before any is used: the mask words and the entries do not depend on each other
      // (the entry of a NULL row is read and dropped: it is storage of the vector, only its content is unspecified)
      ulonglong2 e[kListRpt];
      uint64_t mw[kListRpt];
      bool big = false, outside = false;
      const uint64_t csize = job.child_sizes ? __ldg(job.child_sizes + c) : ~0ull;  // elements in this chunk's child vector
#pragma unroll
      for (int k = 0; k < kListRpt; ++k) {
        const int i = threadIdx.x + k * kThreads;
        mw[k] = (mask && i < count) ? __ldg(mask + (i >> 6)) : ~0ull;
        e[k] = i < count ? __ldg(reinterpret_cast<const ulonglong2 *>(ent) + i) : make_ulonglong2(0ull, 0ull);
      }
#pragma unroll
      for (int k = 0; k < kListRpt; ++k) {
        const int i = threadIdx.x + k * kThreads;
        bool valid = i < count && ((mw[k] >> (i & 63)) & 1ull);
        if (valid && (e[k].x > csize || e[k].y > csize - e[k].x)) {  // a malformed / stale entry: never read outside the child vector
          outside = true;
          valid = false;  // contributes no elements; the error flag makes the host discard the output
        }
        s_src[i] = valid ? e[k].x : 0ull;
        s_start[i] = valid ? (uint32_t)e[k].y : 0u;  // csum <= 4 G: every length fits
        big |= valid && (e[k].y >> 32) != 0ull;
      }
      if (ONEPASS && big) atomicOr(flags, 2ull);  // a list of more than 4 G elements: reported, the output is not usable
      if (outside) atomicOr(flags, 8ull);
    }
    __syncthreads();
    // ---- this thread's kListRpt consecutive rows: block scan of the lengths, starts written back in place
    const int i0 = threadIdx.x * kListRpt;
    uint32_t len[kListRpt];
    uint64_t mine = 0;
#pragma unroll
    for (int k = 0; k < kListRpt; ++k) {
      len[k] = s_start[i0 + k];
      mine += len[k];
    }
    uint64_t total;
    uint64_t ex = block_exscan(mine, &total, s_warp);
    if (ONEPASS) {
      csum = total;
      if (threadIdx.x < 32) {  // warp 0: publish the aggregate, resolve the base, publish the inclusive prefix
#if DMB_LIST_GROUPS
        // two-level look-back (dmb_common.cuh): per-chunk words for the <= 63 nearest chunks, per-group (32 chunks) sums and
        // prefixes before them -- one L2 round trip where the 32-wide chunk-by-chunk walk needed one per 32 chunks
        unsigned long long *gsum = chunk_sum + 2 * b.nchunks, *gpre = gsum + ((b.nchunks + 31) >> 5);
        if (lane == 0) {
          if (c > 0) atomicExch(chunk_sum + c, (1ull << 62) | (csum & kStatusMask));
          atomicAdd(gsum + (c >> 5), kGroupOne | (csum & kGroupSumMask));
        }
        const uint64_t base = c > 0 ? lookback_groups(chunk_sum, gsum, gpre, c, lane, flags, 4ull, 2000000000ull) : 0ull;
        if (lane == 0 && (c & 31) == 31) atomicExch(gpre + (c >> 5), (2ull << 62) | ((base + csum) & kStatusMask));
#else
        if (lane == 0 && c > 0) atomicExch(chunk_sum + c, (1ull << 62) | (csum & kStatusMask));
        const uint64_t base = c > 0 ? list_lookback(chunk_sum, c, lane, flags) : 0ull;
#endif
        if (lane == 0) {
          atomicExch(chunk_sum + c, (2ull << 62) | ((base + csum) & kStatusMask));
          s_cbase = base;
          if (c == b.nchunks - 1) {
            if (job.total) *job.total = base + csum;
            if (!LARGE && base + csum > 0x7fffffffull) atomicOr(flags, 1ull);  // int32 offsets overflow: use large_list
          }
        }
      }
    }
    // contiguity: the chunk's entries are one run in row order iff entry.offset - start is the same for every non-empty row
    unsigned long long dmin = ~0ull, dmax = 0ull;
#pragma unroll
    for (int k = 0; k < kListRpt; ++k) {
      s_start[i0 + k] = (uint32_t)ex;
      if (len[k]) {
        const unsigned long long d = s_src[i0 + k] - ex;
        dmin = d < dmin ? d : dmin;
        dmax = d > dmax ? d : dmax;
      }
      ex += len[k];
    }
    // warp reduction, one slot per warp (64-bit shared atomics are CAS loops: 256 of them on one word cost 30 us per chunk)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long omin = __shfl_xor_sync(0xffffffffu, dmin, d), omax = __shfl_xor_sync(0xffffffffu, dmax, d);
      dmin = omin < dmin ? omin : dmin;
      dmax = omax > dmax ? omax : dmax;
    }
    if (lane == 0) { s_wmin[threadIdx.x >> 5] = dmin; s_wmax[threadIdx.x >> 5] = dmax; }
    if (threadIdx.x == 0) s_start[count] = (uint32_t)csum;
    __syncthreads();
    if (ONEPASS) cbase = s_cbase;
    if (threadIdx.x == 0 && row0 + count == b.nrows) {  // the last chunk writes offsets[nrows]
      const uint64_t o = cbase + csum;
      if (LARGE) reinterpret_cast<long long *>(job.out_offsets)[b.nrows] = (long long)o;
      else reinterpret_cast<int32_t *>(job.out_offsets)[b.nrows] = (int32_t)o;
    }
    for (int i = threadIdx.x; i < count; i += kThreads) {  // offsets: coalesced
      const uint64_t o = cbase + s_start[i];
      if (LARGE) reinterpret_cast<long long *>(job.out_offsets)[row0 + i] = (long long)o;
      else reinterpret_cast<int32_t *>(job.out_offsets)[row0 + i] = (int32_t)o;
    }
    unsigned long long cmin = ~0ull, cmax = 0ull;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      cmin = s_wmin[w] < cmin ? s_wmin[w] : cmin;
      cmax = s_wmax[w] > cmax ? s_wmax[w] : cmax;
    }
    const bool contiguous = cmin == cmax;
    const uint64_t run_d = cmin;  // source element of output element e of a contiguous chunk: run_d + e
    // ---- gather the chunk's child elements, output-centric, in 32-aligned groups of OUTPUT elements
    const uint64_t child0 = __ldg(job.child_base + c);  // element index of the chunk's child vector in the staged slab
    // child validity: one padded mask per chunk (bit = element index inside the chunk's child vector), or, for a
    // second-level gather, ONE bitmap over the whole slab (bit = element index in the slab: cbit0 = the chunk's first element)
    const bool dense_bits = (job.large & DMB_LIST_DENSE_CHILD_BITS) != 0;
    const int64_t cvo = dense_bits ? 0 : (job.child_val_off ? __ldg(job.child_val_off + c) : -1);
    const uint64_t *cmask = (job.child_validity && cvo >= 0) ? job.child_validity + cvo : nullptr;
    const uint64_t cbit0 = dense_bits ? child0 : 0ull;
    const uint64_t first = cbase & ~31ull, end = cbase + csum, stop = (end + 31ull) & ~31ull;
    uint32_t *bm32 = reinterpret_cast<uint32_t *>(job.out_child_validity);
    unsigned nulls = 0;
    if (contiguous) {
      // ---- the chunk's output is one run of its child vector: out[cbase + e] = child[run_d + e].  Aligned 16-byte output
      // vectors (source misaligned by whole elements: two aligned loads + funnel shift), NULL elements zeroed from the mask
      // bits read as aligned 32-bit words; the bitmap is a shifted word copy, one output word per thread.
      constexpr int R = 16 / W;
      const uint8_t *srcb = reinterpret_cast<const uint8_t *>(job.child_data) + (child0 + run_d) * W;
      uint8_t *dstb = reinterpret_cast<uint8_t *>(job.out_child) + cbase * W;
      const uintptr_t d0 = reinterpret_cast<uintptr_t>(dstb);
      const uintptr_t dA = (d0 + 15u) & ~(uintptr_t)15, dE = (d0 + csum * W) & ~(uintptr_t)15;
      const uint32_t *cm32 = reinterpret_cast<const uint32_t *>(cmask);
      uint32_t head = (uint32_t)csum, tail0 = (uint32_t)csum;  // elements [0, head) and [tail0, csum) are copied one by one
      if (dE > dA) {
        head = (uint32_t)((dA - d0) / W);
        tail0 = (uint32_t)((dE - d0) / W);
        const uint32_t nvec = (uint32_t)((dE - dA) >> 4);
        const uint8_t *s0 = srcb + (dA - d0);
        const int m = (int)(reinterpret_cast<uintptr_t>(s0) & 15u);
        const uint4 *sal = reinterpret_cast<const uint4 *>(s0 - m);
        const int ws = m >> 2;
        const uint32_t sh = (uint32_t)(m & 3) * 8u;
        const uint64_t sbit0 = cbit0 + run_d + head;  // mask bit of the first element of vector 0
        uint4 *dal = reinterpret_cast<uint4 *>(dA);
#pragma unroll 2
        for (uint32_t v = threadIdx.x; v < nvec; v += kThreads) {
          const uint4 a = ld_stream(sal + v);
          uint4 o = a;
          if (m) o = shift_words(a, __ldg(sal + v + 1), ws, sh);
          if (cm32) {
            const uint32_t bits = load_bits32(cm32, (int64_t)(sbit0 + (uint64_t)v * R), R);
            o.x &= word_keep<W>(bits, 0);
            o.y &= word_keep<W>(bits, 1);
            o.z &= word_keep<W>(bits, 2);
            o.w &= word_keep<W>(bits, 3);
          }
          st_stream(dal + v, o);
        }
      }
      // the elements in front of / behind the aligned vectors (fewer than 16 / W each; a tiny chunk: all of them)
      {
        using T = typename RawVec<W>::type;
        const uint32_t ntail = (uint32_t)csum - tail0;
        uint32_t e = 0xffffffffu;
        if (threadIdx.x < head) e = threadIdx.x;
        else if (threadIdx.x >= 32 && threadIdx.x - 32 < ntail && tail0 >= head) e = tail0 + (threadIdx.x - 32);
        if (head == (uint32_t)csum) {  // no aligned vector at all: csum < 2 * R elements... or more when dE <= dA: walk them
          for (uint32_t q = threadIdx.x; q < (uint32_t)csum; q += kThreads) {
            const uint64_t src = run_d + q;
            const bool valid = cm32 ? (load_bits32(cm32, (int64_t)(cbit0 + src), 1) != 0u) : true;
            T v;
            memset(&v, 0, sizeof(T));
            if (valid) v = *reinterpret_cast<const T *>(srcb + (uint64_t)q * W);
            *reinterpret_cast<T *>(dstb + (uint64_t)q * W) = v;
          }
        } else if (e != 0xffffffffu) {
          const uint64_t src = run_d + e;
          const bool valid = cm32 ? (load_bits32(cm32, (int64_t)(cbit0 + src), 1) != 0u) : true;
          T v;
          memset(&v, 0, sizeof(T));
          if (valid) v = *reinterpret_cast<const T *>(srcb + (uint64_t)e * W);
          *reinterpret_cast<T *>(dstb + (uint64_t)e * W) = v;
        }
      }
      // bitmap: output word Ew holds the mask bits of elements [Ew, Ew + 32) of the run
      unsigned nulls = 0;
      for (uint64_t Ew = first + 32ull * threadIdx.x; Ew < stop; Ew += 32ull * kThreads) {
        const uint64_t lo = Ew > cbase ? Ew : cbase, hi = Ew + 32 < end ? Ew + 32 : end;
        if (hi <= lo) continue;
        const int nb = (int)(hi - lo);
        const uint32_t bits = cm32 ? load_bits32(cm32, (int64_t)(cbit0 + run_d + (lo - cbase)), nb) : (nb == 32 ? 0xffffffffu : ((1u << nb) - 1u));
        nulls += (unsigned)(nb - __popc(bits));
        if (bm32) {
          if (nb == 32) bm32[Ew >> 5] = bits;
          else if (bits) atomicOr(bm32 + (Ew >> 5), bits << (unsigned)(lo - Ew));  // ragged first / last word: shared with the neighbouring chunks
        }
      }
      if (job.child_null_count) {
        nulls = __reduce_add_sync(0xffffffffu, nulls);
        if (lane == 0 && nulls) atomicAdd(&s_nulls, nulls);
      }
      __syncthreads();
      if (threadIdx.x == 0 && job.child_null_count && s_nulls) atomicAdd(job.child_null_count, (unsigned long long)s_nulls);
      continue;
    }
    constexpr int U = 4;  // output elements per thread in flight
    using T = typename RawVec<W>::type;
    for (uint64_t E0 = first + threadIdx.x; E0 < stop; E0 += (uint64_t)U * kThreads) {  // a warp's 32 elements share one bitmap word
      T v[U];
      bool valid[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint64_t E = E0 + (uint64_t)u * kThreads;
        valid[u] = false;
        memset(&v[u], 0, sizeof(T));
        if (E >= cbase && E < end) {
          const uint32_t e = (uint32_t)(E - cbase);
          uint64_t src;
          if (contiguous) {
            src = run_d + e;
          } else {
            int lo = 0, hi = count;  // the last row whose start is <= e: rows after it start later, empty rows before it are skipped
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if (s_start[mid] <= e) lo = mid; else hi = mid;
            }
            src = s_src[lo] + (e - s_start[lo]);
          }
          // the element is read whether or not it is NULL (it is storage of the child vector): mask word and element in flight together
          const uint64_t cw = cmask ? __ldg(cmask + ((cbit0 + src) >> 6)) : ~0ull;
          v[u] = *reinterpret_cast<const T *>(reinterpret_cast<const uint8_t *>(job.child_data) + (child0 + src) * W);
          valid[u] = (cw >> ((cbit0 + src) & 63)) & 1ull;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint64_t E = E0 + (uint64_t)u * kThreads;
        if (E >= cbase && E < end && !valid[u]) {
          memset(&v[u], 0, sizeof(T));
          ++nulls;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint64_t E = E0 + (uint64_t)u * kThreads;
        if (E >= cbase && E < end) *reinterpret_cast<T *>(reinterpret_cast<uint8_t *>(job.out_child) + E * W) = v[u];
        const uint32_t word = __ballot_sync(0xffffffffu, valid[u]);
        if (lane == 0 && bm32 && E < stop) {
          if (E >= cbase && E + 32 <= end) bm32[E >> 5] = word;   // lane 0's E is 32-aligned
          else if (word) atomicOr(bm32 + (E >> 5), word);         // ragged first / last word: shared with the neighbouring chunks
        }
      }
    }
    if (job.child_null_count) {  // one global atomic per chunk
      nulls = __reduce_add_sync(0xffffffffu, nulls);
      if (lane == 0 && nulls) atomicAdd(&s_nulls, nulls);
    }
    __syncthreads();
    if (threadIdx.x == 0 && job.child_null_count && s_nulls) atomicAdd(job.child_null_count, (unsigned long long)s_nulls);
  }
}

}  // namespace dmb

using namespace dmb;

extern "C" size_t dmb_dev_list_scratch_bytes(int64_t nchunks) {
  const int64_t n = nchunks > 0 ? nchunks : 0;
  return (size_t)(2 * n + 2 + 2 * ((n + 31) / 32)) * sizeof(unsigned long long);
}

// scratch: [0] error flags (1: int32 offsets overflow, 2: a chunk with > 4 G child elements, 4: a look-back gave up
// waiting, 8: a list entry reaches outside its chunk's child vector), [1] unused,
// then chunk_sum[nchunks], chunk_base[nchunks], and a sum word + a prefix word per group of 32 chunks (the one-pass kernel's
// two-level look-back).  out_child_validity must hold ceil(total / 64) + 1 words.
extern "C" int32_t dmb_dev_list_batch(const dmb_list_job *job, const uint32_t *counts, const int64_t *row_off, int64_t nchunks,
                                      int64_t nrows, int64_t child_capacity, void *scratch, void *stream) {
  if (!job) { set_error("dmb_dev_list_batch: job is null"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  if (nchunks <= 0 || nrows <= 0) return 0;
  const int w = job->child_width;
  if (w != 1 && w != 2 && w != 4 && w != 8 && w != 16) { set_error("dmb_dev_list_batch: child width %d (fixed-width children of 1/2/4/8/16 bytes)", w); return -1; }
  unsigned long long *flags = (unsigned long long *)scratch;
  unsigned long long *chunk_sum = flags + 2, *chunk_base = chunk_sum + nchunks;
  if (check_cuda(cudaMemsetAsync(scratch, 0, 16, st), "list scratch memset")) return -1;
  if (job->out_child_validity && child_capacity > 0 &&
      check_cuda(cudaMemsetAsync(job->out_child_validity, 0, (size_t)((child_capacity + 63) / 64 + 1) * 8, st), "list child bitmap memset")) return -1;
  if (job->child_null_count && check_cuda(cudaMemsetAsync(job->child_null_count, 0, 8, st), "list null count memset")) return -1;
  const int64_t max_grid = (int64_t)kNumSMs * 8;
  const int grid = (int)(nchunks < max_grid ? nchunks : max_grid);
  BatchView b{counts, row_off, nchunks, nrows};
  static const bool three_pass = getenv("DMB_LIST_THREE_PASS") != nullptr;  // A/B knob: the sum + scan + emit launches
  if (three_pass) {
    list_sum_kernel<<<grid, kThreads, 0, st>>>(*job, counts, nchunks, chunk_sum);
    list_scan_kernel<<<1, kScanThreads, 0, st>>>(chunk_sum, chunk_base, nchunks, job->total, flags, job->large & 1);
  } else if (check_cuda(cudaMemsetAsync(chunk_sum, 0, (size_t)nchunks * 8, st), "list status memset") ||
             check_cuda(cudaMemsetAsync(chunk_sum + 2 * nchunks, 0, (size_t)(2 * ((nchunks + 31) / 32)) * 8, st), "list group status memset")) {
    return -1;
  }
  auto launch = [&](auto kernel3, auto kernel1) -> int32_t {
    if (three_pass) {
      kernel3<<<grid, kThreads, 0, st>>>(*job, b, chunk_sum, chunk_base, flags);
      return 0;
    }
    // persistent grid, no larger than what is resident at once (the look-back relies on it)
    int per_sm = 0, dev = 0, sms = kNumSMs;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel1, kThreads, 0), "list_emit_kernel occupancy")) return -1;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (per_sm < 1) { set_error("list_emit_kernel does not fit an SM"); return -1; }
    const int64_t resident = (int64_t)per_sm * sms;
    const int g1 = (int)(nchunks < resident ? nchunks : resident);
    kernel1<<<g1, kThreads, 0, st>>>(*job, b, chunk_sum, chunk_base, flags);
    return 0;
  };
#define DMB_LIST_LAUNCH(W)                                                                                          \
  do {                                                                                                              \
    if (job->large & 1) { if (launch(list_emit_kernel<W, true, false>, list_emit_kernel<W, true, true>)) return -1; }   \
    else { if (launch(list_emit_kernel<W, false, false>, list_emit_kernel<W, false, true>)) return -1; }            \
  } while (0)
  switch (w) {
    case 1: DMB_LIST_LAUNCH(1); break;
    case 2: DMB_LIST_LAUNCH(2); break;
    case 4: DMB_LIST_LAUNCH(4); break;
    case 8: DMB_LIST_LAUNCH(8); break;
    default: DMB_LIST_LAUNCH(16); break;
  }
#undef DMB_LIST_LAUNCH
  return check_cuda(cudaGetLastError(), "list kernels launch");
}
