// K8 enum_to_string_t: ENUM vectors (uint8/16/32 indices into the type's dictionary) -> DuckDB-shaped
// string_t that refer to the dictionary's labels, so the VARCHAR form of an ENUM column is one lookup
// kernel followed by the ordinary string_t -> utf8 kernels (K5) with the dictionary as the string heap.
//
// The reference renders an ENUM cell through libduckdb (duckdb_value_varchar, src/duckdb_native.c:215-238,
// 2474-2510) and keeps it as Value::String (src/duckdb_parsing.mbt:119-122); its stream whitelist rejects the type
// (src/duckdb_native.c:271-303), SURVEY.md §8f item 3.  The dictionary is what duckdb_enum_dictionary_size /
// duckdb_enum_dictionary_value return, packed as offsets + bytes (dmb_enum_dict).

#include "dmb_common.cuh"

namespace dmb {

constexpr uint32_t kEnumTable = 1024;  // dictionaries up to this size are turned into a string_t table in shared memory once per CTA

template <typename I>
__global__ void __launch_bounds__(kThreads)
enum_to_string_t_kernel(dmb_enum_job job, const uint32_t *__restrict__ counts, int64_t nchunks) {
  __shared__ uint4 s_tab[kEnumTable];
  const bool table = job.dict_size <= kEnumTable;  // (every uint8 ENUM, most uint16 ones)
  if (table) {
    for (uint32_t t = threadIdx.x; t < job.dict_size; t += kThreads) s_tab[t] = enum_entry(job, t);
    __syncthreads();
  }
  unsigned long long bad = 0;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int count = (int)__ldg(counts + c);
    const dmb_vec_desc vd = job.vecs[c];
    const I *in = reinterpret_cast<const I *>(reinterpret_cast<const uint8_t *>(job.in_data) + vd.data_off);
    const uint64_t *mask = vd.val_off < 0 ? nullptr : job.in_validity + vd.val_off;
    uint4 *out = reinterpret_cast<uint4 *>(job.out) + c * (int64_t)kVec;
    constexpr int kRows = kVec / kThreads;  // 8 rows per thread, all loads issued before the first use
    uint32_t idx[kRows];
    uint64_t mw[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int i = threadIdx.x + k * kThreads;
      idx[k] = i < count ? (uint32_t)in[i] : 0u;   // the index of a NULL row is read and dropped (storage of the vector)
      mw[k] = (mask && i < count) ? __ldg(mask + (i >> 6)) : ~0ull;
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int i = threadIdx.x + k * kThreads;
      if (i >= count) continue;
      uint4 e = make_uint4(0, 0, 0, 0);
      if ((mw[k] >> (i & 63)) & 1ull) {
        if (idx[k] < job.dict_size) e = table ? s_tab[idx[k]] : enum_entry(job, idx[k]);
        else ++bad;  // an index past the dictionary: reported, rendered as the empty string
      }
      st_stream(out + i, e);
    }
  }
  if (job.bad_index && bad) atomicAdd(job.bad_index, bad);
}

}  // namespace dmb

using namespace dmb;

extern "C" int32_t dmb_dev_enum_to_string_t(const dmb_enum_job *job, const uint32_t *counts, int64_t nchunks, void *stream) {
  if (!job) { set_error("dmb_dev_enum_to_string_t: job is null"); return -1; }
  if (nchunks <= 0) return 0;
  if (job->dict_size && (!job->dict_offsets || !job->dict_data)) { set_error("dmb_dev_enum_to_string_t: dictionary is null"); return -1; }
  const int64_t max_grid = (int64_t)kNumSMs * 8;
  const int grid = (int)(nchunks < max_grid ? nchunks : max_grid);
  cudaStream_t st = (cudaStream_t)stream;
  switch (job->phys) {
    case DMB_PHYS_U8: enum_to_string_t_kernel<uint8_t><<<grid, kThreads, 0, st>>>(*job, counts, nchunks); break;
    case DMB_PHYS_U16: enum_to_string_t_kernel<uint16_t><<<grid, kThreads, 0, st>>>(*job, counts, nchunks); break;
    case DMB_PHYS_U32: enum_to_string_t_kernel<uint32_t><<<grid, kThreads, 0, st>>>(*job, counts, nchunks); break;
    default: set_error("dmb_dev_enum_to_string_t: ENUM indices are uint8/uint16/uint32, not physical type %d", job->phys); return -1;
  }
  return check_cuda(cudaGetLastError(), "enum_to_string_t_kernel launch");
}
