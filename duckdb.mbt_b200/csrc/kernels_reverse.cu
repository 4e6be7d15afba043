// K6 arrow_to_chunk: Arrow buffers -> DuckDB DataChunk vectors (the bulk door behind the
// appender).  Chunk k of the output is rows [2048k, 2048k+2048): payload slab at k*2048*W bytes,
// validity slab at k*32 uint64 words, so both slabs are dense in row order.
//   * fixed width: copy / narrow with NULL payloads zeroed, 128-bit accesses when the Arrow slice
//     start is 16-byte aligned (an Arrow `offset` only guarantees element alignment)
//   * Arrow validity bitmap at an arbitrary bit offset -> uint64 masks (funnel shift)
//   * Arrow bool bits -> DuckDB bool bytes
//   * decimal128 -> DECIMAL int64/int32/int16 (low bytes)
//   * utf8 -> duckdb_string_t: <= 12 bytes inlined (zero padded), else 4-byte prefix + pointer to
//     the bytes in place in the *host* Arrow data buffer; validity via warp ballot
//
// Replaces the reference's row-at-a-time appends src/duckdb_native.c:1116-1235 (one FFI call and
// one duckdb_append_* per cell); the vectors are what duckdb_append_data_chunk (:2109-2132) takes.

#include "dmb_common.cuh"

namespace dmb {

// `take` (<= 57) bits starting at bit position p of an LSB bitmap of nbytes bytes
__device__ __forceinline__ uint64_t load_bits(const uint8_t *bm, int64_t nbytes, int64_t p, int take) {
  int64_t b0 = p >> 3;
  uint64_t acc = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int64_t b = b0 + k;
    uint64_t byte = b < nbytes ? (uint64_t)__ldg(bm + b) : 0ull;
    acc |= byte << (8 * k);
  }
  acc >>= (p & 7);
  return take >= 64 ? acc : (acc & ((1ull << take) - 1ull));
}

// 64 validity bits for output word w (rows 64w..64w+63), rows >= nrows are 0
__device__ __forceinline__ uint64_t rev_valid_word(const uint8_t *bm, int64_t bit_offset, int64_t nrows, int64_t w) {
  int64_t r0 = w << 6;
  if (r0 >= nrows) return 0ull;
  int64_t left = nrows - r0;
  int live = left < 64 ? (int)left : 64;
  if (!bm) return live == 64 ? ~0ull : ((1ull << live) - 1ull);
  int64_t nbytes = (bit_offset + nrows + 7) >> 3;
  int64_t p = bit_offset + r0;
  uint64_t lo = load_bits(bm, nbytes, p, 32);
  uint64_t hi = load_bits(bm, nbytes, p + 32, 32);
  uint64_t word = lo | (hi << 32);
  return live == 64 ? word : (word & ((1ull << live) - 1ull));
}

template <typename S, typename D>
__device__ __forceinline__ D narrow(const S &v) {
  D d;
  memcpy(&d, &v, sizeof(D));  // little endian: low bytes
  return d;
}

template <typename S, typename D>
__device__ __forceinline__ void rev_convert(const dmb_rev_fixed_job &job, int64_t nrows) {
  constexpr int W = sizeof(S) > sizeof(D) ? sizeof(S) : sizeof(D);
  constexpr int R = 16 / W;
  using PS = Pack<S, R>;
  using PD = Pack<D, R>;
  const S *in = reinterpret_cast<const S *>(job.in_values);
  D *out = reinterpret_cast<D *>(job.out_data);
  const uint8_t *bm = job.in_validity;
  const int64_t nbytes = (job.in_bit_offset + nrows + 7) >> 3;
  const int64_t nvec = nrows / R;
  const bool in_vec_ok = (reinterpret_cast<uintptr_t>(in) % sizeof(PS)) == 0;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += stride) {
    PS x;
    if (in_vec_ok) {
      x = ld_stream(reinterpret_cast<const PS *>(in) + v);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) x.v[r] = in[v * R + r];
    }
    uint32_t bits = bm ? (uint32_t)load_bits(bm, nbytes, job.in_bit_offset + v * R, R) : 0xffffffffu;
    PD y;
#pragma unroll
    for (int r = 0; r < R; ++r) y.v[r] = ((bits >> r) & 1u) ? narrow<S, D>(x.v[r]) : narrow<S, D>(S{});
    st_stream(reinterpret_cast<PD *>(out) + v, y);
  }
  for (int64_t row = nvec * R + (int64_t)blockIdx.x * kThreads + threadIdx.x; row < nrows; row += stride) {
    bool valid = bm ? (load_bits(bm, nbytes, job.in_bit_offset + row, 1) != 0) : true;
    out[row] = valid ? narrow<S, D>(in[row]) : narrow<S, D>(S{});
  }
}

// Arrow bool bits -> bool bytes: thread handles 8 rows -> one 8-byte store
__device__ __forceinline__ void rev_bits_to_bool(const dmb_rev_fixed_job &job, int64_t nrows) {
  const uint8_t *vals = reinterpret_cast<const uint8_t *>(job.in_values);
  const uint8_t *bm = job.in_validity;
  uint8_t *out = reinterpret_cast<uint8_t *>(job.out_data);
  const int64_t nbytes = (job.in_bit_offset + nrows + 7) >> 3;
  const int64_t ngroups = (nrows + 7) >> 3;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < ngroups; g += stride) {
    int64_t r0 = g << 3;
    int live = nrows - r0 < 8 ? (int)(nrows - r0) : 8;
    uint32_t bits = (uint32_t)load_bits(vals, nbytes, job.in_bit_offset + r0, live);
    if (bm) bits &= (uint32_t)load_bits(bm, nbytes, job.in_bit_offset + r0, live);
    uint64_t bytes = spread8(bits);
    if (live == 8) {
      __stcs(reinterpret_cast<unsigned long long *>(out + r0), bytes);
    } else {
      for (int k = 0; k < live; ++k) out[r0 + k] = (uint8_t)(bytes >> (8 * k));
    }
  }
}

__global__ void __launch_bounds__(kThreads)
rev_fixed_kernel(const dmb_rev_fixed_job *__restrict__ jobs, int64_t nrows) {
  __shared__ dmb_rev_fixed_job s_job;
  if (threadIdx.x < sizeof(dmb_rev_fixed_job) / 8)
    reinterpret_cast<uint64_t *>(&s_job)[threadIdx.x] = reinterpret_cast<const uint64_t *>(jobs + blockIdx.y)[threadIdx.x];
  __syncthreads();
  const dmb_rev_fixed_job &job = s_job;
  if (job.out_data) {
    switch (job.op) {
      case DMB_REV_COPY1: rev_convert<uint8_t, uint8_t>(job, nrows); break;
      case DMB_REV_COPY2: rev_convert<uint16_t, uint16_t>(job, nrows); break;
      case DMB_REV_COPY4: rev_convert<uint32_t, uint32_t>(job, nrows); break;
      case DMB_REV_COPY8: rev_convert<uint64_t, uint64_t>(job, nrows); break;
      case DMB_REV_COPY16: rev_convert<u128, u128>(job, nrows); break;
      case DMB_REV_BITS_TO_BOOL: rev_bits_to_bool(job, nrows); break;
      case DMB_REV_I128_TO_I64: rev_convert<u128, uint64_t>(job, nrows); break;
      case DMB_REV_I128_TO_I32: rev_convert<u128, uint32_t>(job, nrows); break;
      case DMB_REV_I128_TO_I16: rev_convert<u128, uint16_t>(job, nrows); break;
      default: break;
    }
  }
  // validity masks for every vector slot (capacity rows: bits past nrows are 0)
  if (job.out_validity) {
    const int64_t nwords = ((nrows + kVec - 1) / kVec) * DMB_VALIDITY_WORDS;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int nulls = 0;
    for (int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x; w < nwords; w += stride) {
      uint64_t word = rev_valid_word(job.in_validity, job.in_bit_offset, nrows, w);
      __stcs(reinterpret_cast<unsigned long long *>(job.out_validity + w), word);
      int64_t left = nrows - (w << 6);
      int live = left <= 0 ? 0 : (left < 64 ? (int)left : 64);
      nulls += live - __popcll(word);
    }
    if (job.null_count) {
      nulls = __reduce_add_sync(0xffffffffu, nulls);
      if ((threadIdx.x & 31) == 0 && nulls) atomicAdd(job.null_count, (unsigned long long)nulls);
    }
  }
}

template <bool LARGE>
__global__ void __launch_bounds__(kThreads)
rev_string_kernel(dmb_rev_string_job job, int64_t nrows) {
  const int lane = threadIdx.x & 31;
  const int64_t capacity = ((nrows + kVec - 1) / kVec) * (int64_t)kVec;
  const int64_t nbytes = (job.in_bit_offset + nrows + 7) >> 3;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  uint4 *out = reinterpret_cast<uint4 *>(job.out);
  uint32_t *out_val32 = reinterpret_cast<uint32_t *>(job.out_validity);
  int nulls = 0;
  // capacity is a multiple of 2048, so every warp iteration is full: ballots are warp-wide
  for (int64_t row = (int64_t)blockIdx.x * kThreads + threadIdx.x; row < capacity; row += stride) {
    const bool live = row < nrows;
    bool valid = live;
    if (live && job.in_validity) valid = load_bits(job.in_validity, nbytes, job.in_bit_offset + row, 1) != 0;
    uint4 e = make_uint4(0, 0, 0, 0);
    if (valid) {
      int64_t o0, o1;
      if (LARGE) {
        const int64_t *off = reinterpret_cast<const int64_t *>(job.in_offsets);
        o0 = __ldg(off + row); o1 = __ldg(off + row + 1);
      } else {
        const int32_t *off = reinterpret_cast<const int32_t *>(job.in_offsets);
        o0 = __ldg(off + row); o1 = __ldg(off + row + 1);
      }
      const uint32_t len = (uint32_t)(o1 - o0);
      // first 12 bytes of the string with 4 aligned 32-bit loads + funnel shifts
      const uint8_t *q = job.in_data + o0;
      const uint32_t *base = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)3);
      const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3u) * 8u;
      const uint32_t need = len < 12u ? len : 12u;
      const uint32_t nw = need ? ((need + (sh >> 3) + 3u) >> 2) : 0u;  // aligned words that hold needed bytes
      uint32_t w0 = nw > 0 ? __ldg(base) : 0u, w1 = nw > 1 ? __ldg(base + 1) : 0u;
      uint32_t w2 = nw > 2 ? __ldg(base + 2) : 0u, w3 = nw > 3 ? __ldg(base + 3) : 0u;
      uint32_t r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh), r2 = __funnelshift_r(w2, w3, sh);
      e.x = len;
      if (len <= 12u) {
        auto keep = [](uint32_t word, int nb) { return nb >= 4 ? word : (nb <= 0 ? 0u : (word & ((1u << (8 * nb)) - 1u))); };
        e.y = keep(r0, (int)len);
        e.z = keep(r1, (int)len - 4);
        e.w = keep(r2, (int)len - 8);
      } else {
        e.y = r0;  // prefix
        uint64_t p = job.data_host_base + (uint64_t)o0;
        e.z = (uint32_t)p;
        e.w = (uint32_t)(p >> 32);
      }
    }
    st_stream(out + row, e);
    const uint32_t word = __ballot_sync(0xffffffffu, valid);
    if (lane == 0 && out_val32) out_val32[row >> 5] = word;
    nulls += (live && !valid) ? 1 : 0;
  }
  if (job.null_count) {
    nulls = __reduce_add_sync(0xffffffffu, nulls);
    if (lane == 0 && nulls) atomicAdd(job.null_count, (unsigned long long)nulls);
  }
}

}  // namespace dmb

using namespace dmb;

extern "C" int32_t dmb_dev_rev_fixed_batch(const dmb_rev_fixed_job *jobs_dev, const dmb_rev_fixed_job *jobs_host,
                                           int32_t njobs, int64_t nrows, void *stream) {
  if (njobs <= 0 || nrows <= 0) return 0;
  if (!jobs_dev || !jobs_host) { set_error("dmb_dev_rev_fixed_batch: jobs is null"); return -1; }
  for (int32_t j = 0; j < njobs; ++j)
    if (jobs_host[j].op < 0 || jobs_host[j].op >= DMB_REV_COUNT) { set_error("dmb_dev_rev_fixed_batch: bad op %d", jobs_host[j].op); return -1; }
  int64_t blocks = (nrows / 2 + kThreads - 1) / kThreads;  // >= 2 rows per thread at 8 B
  int64_t max_grid = (int64_t)kNumSMs * 8;
  int gx = (int)(blocks < 1 ? 1 : (blocks < max_grid ? blocks : max_grid));
  rev_fixed_kernel<<<dim3(gx, njobs), kThreads, 0, (cudaStream_t)stream>>>(jobs_dev, nrows);
  return check_cuda(cudaGetLastError(), "rev_fixed_kernel launch");
}

extern "C" int32_t dmb_dev_rev_string_batch(const dmb_rev_string_job *job, int64_t nrows, void *stream) {
  if (!job) { set_error("dmb_dev_rev_string_batch: job is null"); return -1; }
  if (nrows <= 0) return 0;
  int64_t capacity = ((nrows + kVec - 1) / kVec) * (int64_t)kVec;
  int64_t blocks = capacity / kThreads;
  int64_t max_grid = (int64_t)kNumSMs * 8;
  int gx = (int)(blocks < max_grid ? blocks : max_grid);
  if (job->large_offsets) rev_string_kernel<true><<<gx, kThreads, 0, (cudaStream_t)stream>>>(*job, nrows);
  else rev_string_kernel<false><<<gx, kThreads, 0, (cudaStream_t)stream>>>(*job, nrows);
  return check_cuda(cudaGetLastError(), "rev_string_kernel launch");
}
