// Reverse path: Arrow record batches (or rows buffered column-wise) -> DuckDB DataChunk vectors,
// gated by the appender protocol of the reference (src/duckdb_appender_state_machine.mbt:54-239).
//
// The reference appends one cell per FFI call (src/duckdb_native.mbt:974-1058 ->
// src/duckdb_native.c:1100-1235).  Here a whole record batch is converted by kernels_reverse.cu in
// sub-batches of a few million rows: copy-in of sub-batch b+1, the kernels of b and copy-out of
// b-1 overlap on the context's three streams, and every finished 2048-row chunk is handed to the
// sink in the layout duckdb_append_data_chunk takes (src/duckdb_native.c:2109-2132).
//
// The row-at-a-time calls of the reference (begin_row / append_* / end_row) are kept as a
// column-wise host buffer that is converted by the same kernels on flush, so the state machine
// is the reference's own.

#include <chrono>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "host_common.hpp"

namespace dmb {
namespace {

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

enum { kNotCreated = 0, kReady = 1, kRowInProgress = 2, kFlushed = 3, kClosed = 4, kError = 5 };

struct RowBuf {  // one column of buffered rows (row API)
  int width = 0;               // bytes per value; 0: string
  int dec_prec = 0, dec_scale = 0;  // DECIMAL: the table column's precision / scale (0: not declared yet); values buffered as int128
  std::vector<uint8_t> values, valid;
  std::vector<int32_t> offsets;  // strings: n+1 entries
  std::vector<uint8_t> data;
};

// one column of a conversion request (host pointers, whole batch)
struct ColInput {
  int rev_op = 0, w_in = 0, w_out = 0;
  bool is_string = false, large = false, is_bits = false;
  const uint8_t *values = nullptr;       // fixed: first value of the slice; bits: bitmap base
  const uint8_t *validity = nullptr;     // Arrow bitmap base or NULL
  int64_t bit_offset = 0;                // array offset in bits (validity and bool values)
  const uint8_t *valid_bytes = nullptr;  // row API: one byte per row instead of a bitmap
  const uint8_t *offsets = nullptr;      // strings: first offset of the slice
  const uint8_t *data = nullptr;         // strings: Arrow data buffer base
};

}  // namespace
}  // namespace dmb

using namespace dmb;

struct duckdb_mb_gpu_appender {
  std::shared_ptr<CtxCore> core;
  int32_t ncols = 0;
  std::vector<int32_t> type_ids;
  dmb_chunk_sink sink = nullptr;
  void *user = nullptr;
  dmb_appender_flush_hook on_flush = nullptr;      // glue: duckdb_appender_flush after the chunks have been handed over
  dmb_appender_destroy_hook on_destroy = nullptr;  // glue: duckdb_appender_destroy + free of `user`
  int state = kNotCreated;
  int cur_col = 0;
  int64_t row_count = 0, flushed_row_count = 0, buffered_rows = 0;
  char error[256] = {0};
  std::vector<RowBuf> rows;
  int64_t sub_rows = 4 << 20;
  double t[4] = {0, 0, 0, 0};
  uint64_t bytes_h2d = 0, bytes_d2h = 0;
};

namespace dmb {
namespace {

typedef duckdb_mb_gpu_appender App;

int32_t fail(App *a, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(a->error, sizeof(a->error), fmt, ap);
  va_end(ap);
  set_error("%s", a->error);
  return 0;
}

// any command the model does not allow in the current state moves it to Error
// (src/duckdb_appender_state_machine.mbt:228-238); Closed and Error are sticky (:218-227)
int32_t illegal(App *a, const char *cmd) {
  if (a->state == kClosed) return fail(a, "%s: appender is closed", cmd);
  if (a->state == kError) return fail(a, "%s: appender is in the error state", cmd);
  static const char *names[] = {"NotCreated", "Ready", "RowInProgress", "Flushed", "Closed", "Error"};
  fail(a, "%s is not allowed in state %s", cmd, names[a->state]);
  a->state = kError;
  return 0;
}

int row_width(int32_t type_id) {
  switch (type_id) {
    case DMB_TYPE_BOOLEAN: case DMB_TYPE_TINYINT: case DMB_TYPE_UTINYINT: return 1;
    case DMB_TYPE_SMALLINT: case DMB_TYPE_USMALLINT: return 2;
    case DMB_TYPE_INTEGER: case DMB_TYPE_UINTEGER: case DMB_TYPE_FLOAT: case DMB_TYPE_DATE: return 4;
    case DMB_TYPE_BIGINT: case DMB_TYPE_UBIGINT: case DMB_TYPE_DOUBLE: case DMB_TYPE_TIMESTAMP:
    case DMB_TYPE_TIMESTAMP_S: case DMB_TYPE_TIMESTAMP_MS: case DMB_TYPE_TIMESTAMP_NS: case DMB_TYPE_TIMESTAMP_TZ:
    case DMB_TYPE_TIME: case DMB_TYPE_TIME_NS: return 8;
    case DMB_TYPE_VARCHAR: case DMB_TYPE_BLOB: return 0;
    case DMB_TYPE_INTERVAL: case DMB_TYPE_DECIMAL: return 16;  // DECIMAL: int128 in the buffer, narrowed on flush
    default: return -1;
  }
}

int copy_op(int w) {
  switch (w) {
    case 1: return DMB_REV_COPY1;
    case 2: return DMB_REV_COPY2;
    case 4: return DMB_REV_COPY4;
    case 8: return DMB_REV_COPY8;
    default: return DMB_REV_COPY16;
  }
}

// Arrow format string of a child -> conversion for a DuckDB column of `type_id`
bool map_format(App *a, int c, const char *fmt, int32_t type_id, ColInput *in) {
  auto fixed = [&](int w) { in->rev_op = copy_op(w); in->w_in = w; in->w_out = w; return true; };
  auto is = [&](const char *f) { return strcmp(fmt, f) == 0; };
  auto starts = [&](const char *f) { return strncmp(fmt, f, strlen(f)) == 0; };
  switch (type_id) {
    case DMB_TYPE_BOOLEAN:
      if (is("b")) { in->rev_op = DMB_REV_BITS_TO_BOOL; in->is_bits = true; in->w_in = 0; in->w_out = 1; return true; }
      break;
    case DMB_TYPE_TINYINT: if (is("c")) return fixed(1); break;
    case DMB_TYPE_UTINYINT: if (is("C")) return fixed(1); break;
    case DMB_TYPE_SMALLINT: if (is("s")) return fixed(2); break;
    case DMB_TYPE_USMALLINT: if (is("S")) return fixed(2); break;
    case DMB_TYPE_INTEGER: if (is("i")) return fixed(4); break;
    case DMB_TYPE_UINTEGER: if (is("I")) return fixed(4); break;
    case DMB_TYPE_FLOAT: if (is("f")) return fixed(4); break;
    case DMB_TYPE_DATE: if (is("tdD")) return fixed(4); break;
    case DMB_TYPE_BIGINT: if (is("l")) return fixed(8); break;
    case DMB_TYPE_UBIGINT: if (is("L")) return fixed(8); break;
    case DMB_TYPE_DOUBLE: if (is("g")) return fixed(8); break;
    case DMB_TYPE_TIME: if (is("ttu")) return fixed(8); break;
    case DMB_TYPE_TIME_NS: if (is("ttn")) return fixed(8); break;
    case DMB_TYPE_TIME_TZ: if (is("L")) return fixed(8); break;
    case DMB_TYPE_TIMESTAMP: case DMB_TYPE_TIMESTAMP_TZ: if (starts("tsu:")) return fixed(8); break;
    case DMB_TYPE_TIMESTAMP_S: if (starts("tss:")) return fixed(8); break;
    case DMB_TYPE_TIMESTAMP_MS: if (starts("tsm:")) return fixed(8); break;
    case DMB_TYPE_TIMESTAMP_NS: if (starts("tsn:")) return fixed(8); break;
    case DMB_TYPE_HUGEINT: if (starts("d:")) return fixed(16); break;
    case DMB_TYPE_UHUGEINT: case DMB_TYPE_UUID: if (is("w:16")) return fixed(16); break;
    case DMB_TYPE_DECIMAL:
      if (starts("d:")) {  // decimal128 -> the physical width DuckDB uses for this precision
        int p = atoi(fmt + 2);
        in->w_in = 16;
        if (p <= 4) { in->rev_op = DMB_REV_I128_TO_I16; in->w_out = 2; }
        else if (p <= 9) { in->rev_op = DMB_REV_I128_TO_I32; in->w_out = 4; }
        else if (p <= 18) { in->rev_op = DMB_REV_I128_TO_I64; in->w_out = 8; }
        else { in->rev_op = DMB_REV_COPY16; in->w_out = 16; }
        return true;
      }
      break;
    case DMB_TYPE_VARCHAR:
      if (is("u") || is("U")) { in->is_string = true; in->large = is("U"); in->w_out = 16; return true; }
      break;
    case DMB_TYPE_BLOB:
      if (is("z") || is("Z")) { in->is_string = true; in->large = is("Z"); in->w_out = 16; return true; }
      break;
    default: break;
  }
  fail(a, "column %d: Arrow format '%s' cannot be appended to a column of DuckDB type %d", c, fmt, type_id);
  return false;
}

struct Slot {  // device + pinned buffers of one in-flight sub-batch
  std::vector<uint8_t *> d_out, h_out;
  std::vector<uint64_t *> d_val, h_val;
  cudaEvent_t ev_out = nullptr;
  int64_t r0 = 0, r1 = 0;
  bool busy = false;
};

int32_t sink_slot(App *a, Slot &s, const std::vector<ColInput> &cols, int64_t nrows) {
  if (!s.busy) return 1;
  if (check_cuda(cudaEventSynchronize(s.ev_out), "appender copy-out wait")) return fail(a, "%s", duckdb_mb_gpu_last_error());
  s.busy = false;
  if (!a->sink) return 1;
  const int nc = (int)cols.size();
  std::vector<const void *> vd((size_t)nc);
  std::vector<const uint64_t *> vv((size_t)nc);
  for (int64_t r = s.r0, k = 0; r < s.r1; r += DMB_VECTOR_SIZE, ++k) {
    const uint32_t count = (uint32_t)(s.r1 - r < DMB_VECTOR_SIZE ? s.r1 - r : DMB_VECTOR_SIZE);
    for (int c = 0; c < nc; ++c) {
      vd[(size_t)c] = s.h_out[(size_t)c] + (size_t)k * DMB_VECTOR_SIZE * (size_t)cols[(size_t)c].w_out;
      vv[(size_t)c] = s.h_val[(size_t)c] + (size_t)k * DMB_VALIDITY_WORDS;
    }
    if (!a->sink(a->user, nc, count, vd.data(), vv.data())) return fail(a, "chunk sink aborted at row %lld", (long long)r);
  }
  (void)nrows;
  return 1;
}

// convert `nrows` rows described by `cols` and hand the chunks to the sink
int32_t convert_rows(App *a, const std::vector<ColInput> &cols, int64_t nrows) {
  if (nrows <= 0) return 1;
  CtxCore &c = *a->core;
  std::lock_guard<std::mutex> g(c.mu);
  if (!c.bind()) return fail(a, "%s", duckdb_mb_gpu_last_error());
  const double t0 = now_ms();
  const int nc = (int)cols.size();
  int64_t sub = a->sub_rows < nrows ? a->sub_rows : ((nrows + DMB_VECTOR_SIZE - 1) / DMB_VECTOR_SIZE) * DMB_VECTOR_SIZE;
  const int64_t sub_chunks = sub / DMB_VECTOR_SIZE;
  Slot slots[2];
  Scope sc(c);
  cudaEvent_t in0 = sc.event(true), in1 = sc.event(true), out0 = sc.event(true), out1 = sc.event(true);
  if (!in0 || !in1 || !out0 || !out1) return fail(a, "%s", duckdb_mb_gpu_last_error());
  const int nslots = nrows > sub ? 2 : 1;
  for (int s = 0; s < nslots; ++s) {
    slots[s].ev_out = sc.event(false);
    for (int j = 0; j < nc; ++j) {
      const size_t ob = (size_t)sub * (size_t)cols[(size_t)j].w_out, vb = (size_t)sub_chunks * DMB_VALIDITY_WORDS * 8;
      uint8_t *d = (uint8_t *)sc.dalloc(ob), *h = (uint8_t *)sc.palloc(ob);
      uint64_t *dv = (uint64_t *)sc.dalloc(vb), *hv = (uint64_t *)sc.palloc(vb);
      if (!d || !h || !dv || !hv || !slots[s].ev_out) return fail(a, "%s", duckdb_mb_gpu_last_error());
      slots[s].d_out.push_back(d);
      slots[s].h_out.push_back(h);
      slots[s].d_val.push_back(dv);
      slots[s].h_val.push_back(hv);
    }
  }
  // per-slot device inputs
  struct InBuf { uint8_t *values = nullptr, *validity = nullptr, *offsets = nullptr, *data = nullptr; size_t data_cap = 0; };
  std::vector<InBuf> inbuf((size_t)nslots * (size_t)nc);
  for (int s = 0; s < nslots; ++s)
    for (int j = 0; j < nc; ++j) {
      const ColInput &ci = cols[(size_t)j];
      InBuf &ib = inbuf[(size_t)s * (size_t)nc + (size_t)j];
      if (ci.is_string) ib.offsets = (uint8_t *)sc.dalloc((size_t)(sub + 1) * (ci.large ? 8 : 4) + 64);
      else ib.values = (uint8_t *)sc.dalloc(ci.is_bits ? (size_t)(sub / 8 + 16) : (size_t)sub * (size_t)ci.w_in + 64);
      if (ci.validity) ib.validity = (uint8_t *)sc.dalloc((size_t)(sub / 8 + 16) + 64);
      else if (ci.valid_bytes) ib.validity = (uint8_t *)sc.dalloc((size_t)sub + (size_t)(sub / 8) + 128);  // bytes, then masks
      if ((ci.is_string && !ib.offsets) || (!ci.is_string && !ib.values) || ((ci.validity || ci.valid_bytes) && !ib.validity))
        return fail(a, "%s", duckdb_mb_gpu_last_error());
    }
  dmb_rev_fixed_job *h_jobs = (dmb_rev_fixed_job *)sc.palloc(sizeof(dmb_rev_fixed_job) * (size_t)nc * 2);
  dmb_rev_fixed_job *d_jobs = (dmb_rev_fixed_job *)sc.dalloc(sizeof(dmb_rev_fixed_job) * (size_t)nc * 2);
  if (!h_jobs || !d_jobs) return fail(a, "%s", duckdb_mb_gpu_last_error());
  auto offset_at = [](const ColInput &ci, int64_t row) -> int64_t {
    if (ci.large) { int64_t v; memcpy(&v, ci.offsets + 8 * row, 8); return v; }
    int32_t v; memcpy(&v, ci.offsets + 4 * row, 4); return v;
  };
  cudaEventRecord(in0, c.s_in);
  cudaEventRecord(out0, c.s_out);
  int64_t b = 0;
  for (int64_t r0 = 0; r0 < nrows; r0 += sub, ++b) {
    const int64_t r1 = r0 + sub < nrows ? r0 + sub : nrows;
    const int64_t n = r1 - r0;
    const int si = (int)(b % nslots);
    Slot &slot = slots[si];
    if (!sink_slot(a, slot, cols, nrows)) return 0;  // the slot's previous sub-batch leaves first
    // the job array of this slot is reused as well: its previous launch has completed (sink_slot waited)
    dmb_rev_fixed_job *hj = h_jobs + (size_t)si * (size_t)nc, *dj = d_jobs + (size_t)si * (size_t)nc;
    int nfixed = 0;
    cudaEvent_t ev_in = sc.event(false), ev_k = sc.event(false), k0 = sc.event(true), k1 = sc.event(true);
    if (!ev_in || !ev_k || !k0 || !k1) return fail(a, "%s", duckdb_mb_gpu_last_error());
    struct StrLaunch { dmb_rev_string_job job; };
    std::vector<StrLaunch> strs;
    std::vector<std::pair<uint8_t *, int>> masks_from_bytes;  // (buffer, column)
    for (int j = 0; j < nc; ++j) {
      const ColInput &ci = cols[(size_t)j];
      InBuf &ib = inbuf[(size_t)si * (size_t)nc + (size_t)j];
      const uint8_t *d_validity = nullptr;
      int64_t bit_off = 0;
      if (ci.validity) {
        const int64_t p0 = ci.bit_offset + r0, p1 = ci.bit_offset + r1;
        const int64_t byte0 = p0 >> 3, byte1 = (p1 + 7) >> 3;
        if (stage_contiguous(c, c.s_in, ib.validity, ci.validity + byte0, (size_t)(byte1 - byte0), -1, &a->bytes_h2d)) return fail(a, "%s", duckdb_mb_gpu_last_error());
        d_validity = ib.validity;
        bit_off = p0 & 7;
      } else if (ci.valid_bytes) {
        if (stage_contiguous(c, c.s_in, ib.validity, ci.valid_bytes + r0, (size_t)n, -1, &a->bytes_h2d)) return fail(a, "%s", duckdb_mb_gpu_last_error());
        uint8_t *masks = ib.validity + (((size_t)sub + 63) & ~(size_t)63);
        masks_from_bytes.emplace_back(ib.validity, j);
        d_validity = masks;
        bit_off = 0;
      }
      if (ci.is_string) {
        const int ow = ci.large ? 8 : 4;
        if (stage_contiguous(c, c.s_in, ib.offsets, ci.offsets + (size_t)r0 * (size_t)ow, (size_t)(n + 1) * (size_t)ow, -1, &a->bytes_h2d)) return fail(a, "%s", duckdb_mb_gpu_last_error());
        const int64_t o_first = offset_at(ci, r0), o_last = offset_at(ci, r1);
        if (o_last < o_first || o_first < 0) return fail(a, "column %d: utf8 offsets are not monotonic", j);
        const size_t dbytes = (size_t)(o_last - o_first);
        if (dbytes + 64 > ib.data_cap) {  // grows to the largest sub-batch seen; old block returns to the pool with the scope
          ib.data_cap = dbytes + dbytes / 8 + 64;
          ib.data = (uint8_t *)sc.dalloc(ib.data_cap);
          if (!ib.data) return fail(a, "%s", duckdb_mb_gpu_last_error());
        }
        if (dbytes && stage_contiguous(c, c.s_in, ib.data, ci.data + o_first, dbytes, -1, &a->bytes_h2d)) return fail(a, "%s", duckdb_mb_gpu_last_error());
        StrLaunch sl;
        memset(&sl.job, 0, sizeof(sl.job));
        sl.job.in_offsets = ib.offsets;
        sl.job.in_data = ib.data - o_first;  // indexed by absolute Arrow offsets
        sl.job.in_validity = d_validity;
        sl.job.in_bit_offset = bit_off;
        sl.job.data_host_base = (uint64_t)(uintptr_t)ci.data;
        sl.job.out = (dmb_string_t *)slot.d_out[(size_t)j];
        sl.job.out_validity = slot.d_val[(size_t)j];
        sl.job.large_offsets = ci.large ? 1 : 0;
        strs.push_back(sl);
      } else {
        dmb_rev_fixed_job &job = hj[nfixed++];
        memset(&job, 0, sizeof(job));
        if (ci.is_bits) {
          const int64_t p0 = ci.bit_offset + r0, p1 = ci.bit_offset + r1;
          const int64_t byte0 = p0 >> 3, byte1 = (p1 + 7) >> 3;
          if (stage_contiguous(c, c.s_in, ib.values, ci.values + byte0, (size_t)(byte1 - byte0), -1, &a->bytes_h2d)) return fail(a, "%s", duckdb_mb_gpu_last_error());
          bit_off = p0 & 7;  // same array offset as the validity bitmap
        } else {
          if (stage_contiguous(c, c.s_in, ib.values, ci.values + (size_t)r0 * (size_t)ci.w_in, (size_t)n * (size_t)ci.w_in, -1, &a->bytes_h2d)) return fail(a, "%s", duckdb_mb_gpu_last_error());
        }
        job.in_values = ib.values;
        job.in_validity = d_validity;
        job.in_bit_offset = bit_off;
        job.out_data = slot.d_out[(size_t)j];
        job.out_validity = slot.d_val[(size_t)j];
        job.op = ci.rev_op;
      }
    }
    cudaEventRecord(ev_in, c.s_in);
    if (check_cuda(cudaStreamWaitEvent(c.s_compute, ev_in, 0), "wait copy-in")) return fail(a, "%s", duckdb_mb_gpu_last_error());
    cudaEventRecord(k0, c.s_compute);
    for (auto &mb : masks_from_bytes) {
      uint8_t *masks = mb.first + (((size_t)sub + 63) & ~(size_t)63);
      if (dmb_dev_valid_bytes_to_masks(mb.first, (uint64_t *)masks, nullptr, n, c.s_compute)) return fail(a, "%s", duckdb_mb_gpu_last_error());
    }
    if (nfixed) {
      if (check_cuda(cudaMemcpyAsync(dj, hj, sizeof(dmb_rev_fixed_job) * (size_t)nfixed, cudaMemcpyHostToDevice, c.s_compute), "job H2D")) return fail(a, "%s", duckdb_mb_gpu_last_error());
      if (dmb_dev_rev_fixed_batch(dj, hj, nfixed, n, c.s_compute)) return fail(a, "%s", duckdb_mb_gpu_last_error());
    }
    for (auto &sl : strs)
      if (dmb_dev_rev_string_batch(&sl.job, n, c.s_compute)) return fail(a, "%s", duckdb_mb_gpu_last_error());
    cudaEventRecord(k1, c.s_compute);
    sc.kernel_spans.emplace_back(k0, k1);
    cudaEventRecord(ev_k, c.s_compute);
    if (check_cuda(cudaStreamWaitEvent(c.s_out, ev_k, 0), "wait kernels")) return fail(a, "%s", duckdb_mb_gpu_last_error());
    const int64_t nch = (n + DMB_VECTOR_SIZE - 1) / DMB_VECTOR_SIZE;
    for (int j = 0; j < nc; ++j) {
      const size_t ob = (size_t)n * (size_t)cols[(size_t)j].w_out, vb = (size_t)nch * DMB_VALIDITY_WORDS * 8;
      if (check_cuda(cudaMemcpyAsync(slot.h_out[(size_t)j], slot.d_out[(size_t)j], ob, cudaMemcpyDeviceToHost, c.s_out), "vectors D2H") ||
          check_cuda(cudaMemcpyAsync(slot.h_val[(size_t)j], slot.d_val[(size_t)j], vb, cudaMemcpyDeviceToHost, c.s_out), "validity D2H"))
        return fail(a, "%s", duckdb_mb_gpu_last_error());
      a->bytes_d2h += ob + vb;
    }
    cudaEventRecord(slot.ev_out, c.s_out);
    slot.r0 = r0;
    slot.r1 = r1;
    slot.busy = true;
  }
  cudaEventRecord(in1, c.s_in);
  for (int k = 0; k < nslots; ++k) {  // oldest first
    Slot &slot = slots[(int)((b + k) % nslots)];
    if (!sink_slot(a, slot, cols, nrows)) return 0;
  }
  cudaEventRecord(out1, c.s_out);
  if (check_cuda(cudaStreamSynchronize(c.s_out), "appender sync") || check_cuda(cudaStreamSynchronize(c.s_in), "appender sync"))
    return fail(a, "%s", duckdb_mb_gpu_last_error());
  float f = 0;
  a->t[0] = cudaEventElapsedTime(&f, in0, in1) == cudaSuccess ? f : 0;
  a->t[1] = sc.kernel_ms();
  a->t[2] = cudaEventElapsedTime(&f, out0, out1) == cudaSuccess ? f : 0;
  a->t[3] = now_ms() - t0;
  return 1;
}

// rows buffered by the row API -> chunks
int32_t flush_rows(App *a) {
  if (a->buffered_rows == 0) return 1;
  std::vector<ColInput> cols((size_t)a->ncols);
  for (int j = 0; j < a->ncols; ++j) {
    RowBuf &rb = a->rows[(size_t)j];
    ColInput &ci = cols[(size_t)j];
    ci.valid_bytes = rb.valid.data();
    if (rb.width == 0) {
      ci.is_string = true;
      ci.w_out = 16;
      ci.offsets = (const uint8_t *)rb.offsets.data();
      if (rb.data.empty()) rb.data.push_back(0);
      ci.data = rb.data.data();
    } else {
      ci.rev_op = copy_op(rb.width);
      ci.w_in = ci.w_out = rb.width;
      ci.values = rb.values.data();
      if (a->type_ids[(size_t)j] == DMB_TYPE_DECIMAL) {  // int128 -> the physical width DuckDB uses for the precision
        if (rb.dec_prec <= 0) {
          fail(a, "flush: DECIMAL column %d has no declared precision (duckdb_mb_gpu_appender_set_decimal)", j);
          return 0;
        }
        if (rb.dec_prec <= 4) { ci.rev_op = DMB_REV_I128_TO_I16; ci.w_out = 2; }
        else if (rb.dec_prec <= 9) { ci.rev_op = DMB_REV_I128_TO_I32; ci.w_out = 4; }
        else if (rb.dec_prec <= 18) { ci.rev_op = DMB_REV_I128_TO_I64; ci.w_out = 8; }
      }
    }
  }
  const int32_t ok = convert_rows(a, cols, a->buffered_rows);
  for (RowBuf &rb : a->rows) {
    rb.values.clear();
    rb.valid.clear();
    rb.data.clear();
    rb.offsets.assign(1, 0);
  }
  a->buffered_rows = 0;
  return ok;
}

RowBuf *cell(App *a, const char *cmd) {
  if (a->state != kRowInProgress) { illegal(a, cmd); return nullptr; }
  if (a->cur_col >= a->ncols) {  // over-filled row (src/duckdb_appender_state_machine.mbt:96-111)
    fail(a, "%s: row already has %d values", cmd, a->ncols);
    a->state = kError;
    return nullptr;
  }
  return &a->rows[(size_t)a->cur_col];
}

template <typename T>
void push_value(RowBuf *rb, T v) {
  const size_t at = rb->values.size();
  rb->values.resize(at + sizeof(T));
  memcpy(rb->values.data() + at, &v, sizeof(T));
}

int32_t type_mismatch(App *a, const char *cmd) {
  fail(a, "%s: column %d has DuckDB type %d", cmd, a->cur_col, a->type_ids[(size_t)a->cur_col]);
  a->state = kError;
  return 0;
}

// numeric cell into a numeric column, C conversion like DuckDB's implicit appender cast
int32_t push_number(App *a, const char *cmd, int64_t iv, double dv, bool is_double) {
  RowBuf *rb = cell(a, cmd);
  if (!rb) return 0;
  const int32_t t = a->type_ids[(size_t)a->cur_col];
  const int64_t as_int = is_double ? (int64_t)nearbyint(dv) : iv;
  const double as_dbl = is_double ? dv : (double)iv;
  switch (t) {
    case DMB_TYPE_TINYINT: case DMB_TYPE_UTINYINT: push_value<int8_t>(rb, (int8_t)as_int); break;
    case DMB_TYPE_SMALLINT: case DMB_TYPE_USMALLINT: push_value<int16_t>(rb, (int16_t)as_int); break;
    case DMB_TYPE_INTEGER: case DMB_TYPE_UINTEGER: case DMB_TYPE_DATE: push_value<int32_t>(rb, (int32_t)as_int); break;
    case DMB_TYPE_BIGINT: case DMB_TYPE_UBIGINT: case DMB_TYPE_TIMESTAMP: case DMB_TYPE_TIMESTAMP_S: case DMB_TYPE_TIMESTAMP_MS:
    case DMB_TYPE_TIMESTAMP_NS: case DMB_TYPE_TIMESTAMP_TZ: case DMB_TYPE_TIME: case DMB_TYPE_TIME_NS:
      push_value<int64_t>(rb, as_int); break;
    case DMB_TYPE_FLOAT: push_value<float>(rb, (float)as_dbl); break;
    case DMB_TYPE_DOUBLE: push_value<double>(rb, as_dbl); break;
    default: return type_mismatch(a, cmd);
  }
  rb->valid.push_back(1);
  a->cur_col++;
  return 1;
}

}  // namespace
}  // namespace dmb

extern "C" duckdb_mb_gpu_appender *duckdb_mb_gpu_appender_create(duckdb_mb_gpu_ctx *ctx, int32_t ncols, const int32_t *type_ids,
                                                                 dmb_chunk_sink sink, void *user) {
  if (!ctx || !ctx->core) { set_error("duckdb_mb_gpu_appender_create: null context"); return nullptr; }
  if (ncols <= 0 || !type_ids) { set_error("duckdb_mb_gpu_appender_create: no columns"); return nullptr; }
  App *a = new App();
  a->core = ctx->core;
  a->ncols = ncols;
  a->type_ids.assign(type_ids, type_ids + ncols);
  a->sink = sink;
  a->user = user;
  a->rows.resize((size_t)ncols);
  for (int j = 0; j < ncols; ++j) {
    a->rows[(size_t)j].width = row_width(type_ids[j]);  // -1: column only reachable through Arrow batches
    a->rows[(size_t)j].offsets.assign(1, 0);
  }
  const char *env = getenv("DMB_REV_BATCH_ROWS");
  if (env && atoll(env) >= DMB_VECTOR_SIZE) a->sub_rows = (atoll(env) / DMB_VECTOR_SIZE) * DMB_VECTOR_SIZE;
  a->state = kReady;  // Create: NotCreated -> Ready
  return a;
}

extern "C" void duckdb_mb_gpu_appender_destroy(duckdb_mb_gpu_appender *a) {
  if (!a) return;
  if (a->on_destroy) a->on_destroy(a->user);
  delete a;
}

extern "C" void duckdb_mb_gpu_appender_set_hooks(duckdb_mb_gpu_appender *a, dmb_appender_flush_hook on_flush, dmb_appender_destroy_hook on_destroy) {
  if (!a) return;
  a->on_flush = on_flush;
  a->on_destroy = on_destroy;
}

extern "C" moonbit_bytes_t duckdb_mb_gpu_appender_error(duckdb_mb_gpu_appender *a) {
  const char *msg = a ? a->error : "";
  const size_t n = strlen(msg);
  moonbit_bytes_t b = moonbit_make_bytes_raw((int32_t)n);
  if (b) memcpy(b, msg, n);
  return b;
}

extern "C" int32_t duckdb_mb_gpu_appender_state(duckdb_mb_gpu_appender *a) { return a ? a->state : kNotCreated; }
extern "C" int64_t duckdb_mb_gpu_appender_row_count(duckdb_mb_gpu_appender *a) { return a ? a->row_count : 0; }
extern "C" int64_t duckdb_mb_gpu_appender_flushed_row_count(duckdb_mb_gpu_appender *a) { return a ? a->flushed_row_count : 0; }

extern "C" int32_t duckdb_mb_gpu_append_arrow_batch(duckdb_mb_gpu_appender *a, const struct ArrowArray *batch,
                                                    const struct ArrowSchema *schema) {
  NvtxRange nvtx("dmb::append_arrow_batch");
  if (!a) { set_error("null appender"); return 0; }
  if (a->state != kReady && a->state != kFlushed) return illegal(a, "append_arrow_batch");  // legal where BeginRow is
  if (!batch || !schema || !schema->format || strcmp(schema->format, "+s") != 0) {
    fail(a, "append_arrow_batch: expected a struct array (record batch)");
    a->state = kError;
    return 0;
  }
  if (batch->n_children != a->ncols || schema->n_children != a->ncols) {
    fail(a, "append_arrow_batch: batch has %lld columns, table has %d", (long long)batch->n_children, a->ncols);
    a->state = kError;
    return 0;
  }
  if (batch->null_count > 0) {
    fail(a, "append_arrow_batch: null struct rows are not supported");
    a->state = kError;
    return 0;
  }
  const int64_t n = batch->length;
  std::vector<ColInput> cols((size_t)a->ncols);
  for (int j = 0; j < a->ncols; ++j) {
    const ArrowArray *ch = batch->children[j];
    const ArrowSchema *cs = schema->children[j];
    ColInput &ci = cols[(size_t)j];
    if (!ch || !cs || !cs->format || !map_format(a, j, cs->format, a->type_ids[(size_t)j], &ci)) {
      if (!a->error[0]) fail(a, "append_arrow_batch: column %d is malformed", j);
      a->state = kError;
      return 0;
    }
    if (ch->length < batch->offset + n) {
      fail(a, "append_arrow_batch: column %d is shorter than the batch", j);
      a->state = kError;
      return 0;
    }
    const int64_t off = ch->offset + batch->offset;
    ci.bit_offset = off;
    ci.validity = (ch->null_count != 0 && ch->n_buffers > 0) ? (const uint8_t *)ch->buffers[0] : nullptr;
    if (ci.is_string) {
      ci.offsets = (const uint8_t *)ch->buffers[1] + (size_t)off * (ci.large ? 8 : 4);
      ci.data = (const uint8_t *)ch->buffers[2];
      static const uint8_t kNoData[16] = {0};
      if (!ci.data) ci.data = kNoData;
    } else if (ci.is_bits) {
      ci.values = (const uint8_t *)ch->buffers[1];
    } else {
      ci.values = (const uint8_t *)ch->buffers[1] + (size_t)off * (size_t)ci.w_in;
    }
    if (n > 0 && ((ci.is_string && !ci.offsets) || (!ci.is_string && !ci.values))) {
      fail(a, "append_arrow_batch: column %d has no value buffer", j);
      a->state = kError;
      return 0;
    }
  }
  if (!flush_rows(a)) { a->state = kError; return 0; }  // keep row order with earlier row-API appends
  if (!convert_rows(a, cols, n)) { a->state = kError; return 0; }
  a->row_count += n;
  a->state = kReady;
  return 1;
}

// ---- row-at-a-time protocol of the reference (src/duckdb_native.c:1100-1235), column-buffered
extern "C" int32_t duckdb_mb_gpu_begin_row(duckdb_mb_gpu_appender *a) {
  if (!a) return 0;
  if (a->state != kReady && a->state != kFlushed) return illegal(a, "begin_row");
  for (int j = 0; j < a->ncols; ++j)
    if (a->rows[(size_t)j].width < 0) { fail(a, "begin_row: column %d (type %d) is only appendable through Arrow batches", j, a->type_ids[(size_t)j]); a->state = kError; return 0; }
  a->state = kRowInProgress;
  a->cur_col = 0;
  return 1;
}

extern "C" int32_t duckdb_mb_gpu_append_int(duckdb_mb_gpu_appender *a, int32_t v) { return a ? push_number(a, "append_int", v, 0, false) : 0; }
extern "C" int32_t duckdb_mb_gpu_append_bigint(duckdb_mb_gpu_appender *a, int64_t v) { return a ? push_number(a, "append_bigint", v, 0, false) : 0; }
extern "C" int32_t duckdb_mb_gpu_append_double(duckdb_mb_gpu_appender *a, double v) { return a ? push_number(a, "append_double", 0, v, true) : 0; }
extern "C" int32_t duckdb_mb_gpu_append_date(duckdb_mb_gpu_appender *a, int32_t days) { return a ? push_number(a, "append_date", days, 0, false) : 0; }
extern "C" int32_t duckdb_mb_gpu_append_timestamp(duckdb_mb_gpu_appender *a, int64_t micros) { return a ? push_number(a, "append_timestamp", micros, 0, false) : 0; }

// BLOB cell (src/duckdb_native.c:1397-1415): the bytes as they are, no UTF-8 requirement
extern "C" int32_t duckdb_mb_gpu_append_blob(duckdb_mb_gpu_appender *a, const uint8_t *bytes, int32_t len) {
  return duckdb_mb_gpu_append_varchar(a, bytes, len);
}

// LIST / STRUCT / MAP cells of the reference's row protocol: the reference serialises them as text and appends that as
// a VARCHAR cell (src/duckdb_native.c:1735-1790 `["a", "b"]`, :1792-1858 and :1860-1926 `{"k": "v", ...}`; no escaping).
// The text reaches libduckdb through duckdb_append_varchar, a C string: it ends at the first NUL byte of any item.
namespace {
void put_quoted(std::vector<uint8_t> &t, const uint8_t *p, int32_t len) {
  t.push_back('"');
  if (p && len > 0) t.insert(t.end(), p, p + len);
  t.push_back('"');
}
int32_t append_text_cell(duckdb_mb_gpu_appender *a, std::vector<uint8_t> &t) {
  size_t n = 0;
  while (n < t.size() && t[n] != 0) ++n;  // strlen of the reference's buffer
  return duckdb_mb_gpu_append_varchar(a, t.data(), (int32_t)n);
}
}  // namespace

extern "C" int32_t duckdb_mb_gpu_append_list_varchar(duckdb_mb_gpu_appender *a, const uint8_t *const *values, const int32_t *lens, int32_t count) {
  if (!a) return 0;
  if (count < 0 || (count > 0 && (!values || !lens))) { fail(a, "append_list_varchar: bad arguments"); a->state = kError; return 0; }
  std::vector<uint8_t> t;
  t.push_back('[');
  for (int32_t i = 0; i < count; ++i) {
    if (i > 0) { t.push_back(','); t.push_back(' '); }
    put_quoted(t, values[i], lens[i]);
  }
  t.push_back(']');
  return append_text_cell(a, t);
}

static int32_t append_pairs(duckdb_mb_gpu_appender *a, const char *what, const uint8_t *const *keys, const int32_t *key_lens,
                            const uint8_t *const *values, const int32_t *value_lens, int32_t count) {
  if (!a) return 0;
  if (count < 0 || (count > 0 && (!keys || !key_lens || !values || !value_lens))) { fail(a, "%s: bad arguments", what); a->state = kError; return 0; }
  std::vector<uint8_t> t;
  t.push_back('{');
  for (int32_t i = 0; i < count; ++i) {
    if (i > 0) { t.push_back(','); t.push_back(' '); }
    put_quoted(t, keys[i], key_lens[i]);
    t.push_back(':');
    t.push_back(' ');
    put_quoted(t, values[i], value_lens[i]);
  }
  t.push_back('}');
  return append_text_cell(a, t);
}

extern "C" int32_t duckdb_mb_gpu_append_struct_varchar(duckdb_mb_gpu_appender *a, const uint8_t *const *names, const int32_t *name_lens,
                                                      const uint8_t *const *values, const int32_t *value_lens, int32_t count) {
  return append_pairs(a, "append_struct_varchar", names, name_lens, values, value_lens, count);
}

extern "C" int32_t duckdb_mb_gpu_append_map_varchar_varchar(duckdb_mb_gpu_appender *a, const uint8_t *const *keys, const int32_t *key_lens,
                                                           const uint8_t *const *values, const int32_t *value_lens, int32_t count) {
  return append_pairs(a, "append_map_varchar_varchar", keys, key_lens, values, value_lens, count);
}

// INTERVAL cell (src/duckdb_native.c:1511-1533): duckdb_interval {months, days, micros}
extern "C" int32_t duckdb_mb_gpu_append_interval(duckdb_mb_gpu_appender *a, int32_t months, int32_t days, int64_t micros) {
  if (!a) return 0;
  RowBuf *rb = cell(a, "append_interval");
  if (!rb) return 0;
  if (a->type_ids[(size_t)a->cur_col] != DMB_TYPE_INTERVAL) return type_mismatch(a, "append_interval");
  interval_t v{months, days, micros};
  const uint8_t *p = reinterpret_cast<const uint8_t *>(&v);
  rb->values.insert(rb->values.end(), p, p + sizeof(v));
  rb->valid.push_back(1);
  a->cur_col++;
  return 1;
}

// the table column's DECIMAL(width, scale) (duckdb_appender_column_type -> duckdb_decimal_width / _scale)
extern "C" int32_t duckdb_mb_gpu_appender_set_decimal(duckdb_mb_gpu_appender *a, int32_t col, int32_t width, int32_t scale) {
  if (!a) return 0;
  if (col < 0 || col >= a->ncols || a->type_ids[(size_t)col] != DMB_TYPE_DECIMAL || width < 1 || width > 38 || scale < 0 || scale > width) {
    fail(a, "appender_set_decimal: column %d is not a DECIMAL column or DECIMAL(%d,%d) is not a type", col, width, scale);
    return 0;
  }
  RowBuf &rb = a->rows[(size_t)col];
  if (rb.dec_prec && (rb.dec_prec != width || rb.dec_scale != scale)) {
    fail(a, "appender_set_decimal: column %d is already DECIMAL(%d,%d)", col, rb.dec_prec, rb.dec_scale);
    return 0;
  }
  rb.dec_prec = width;
  rb.dec_scale = scale;
  return 1;
}

// DECIMAL cell from hugeint parts (src/duckdb_native.c:1447-1481: duckdb_create_decimal + duckdb_append_value).
// The value is brought to the column's scale like DuckDB's decimal -> decimal cast (exact when the scale grows,
// round half away from zero when it shrinks) and must fit the column's precision.
extern "C" int32_t duckdb_mb_gpu_append_decimal(duckdb_mb_gpu_appender *a, int32_t width, int32_t scale, int64_t lower, int64_t upper) {
  if (!a) return 0;
  RowBuf *rb = cell(a, "append_decimal");
  if (!rb) return 0;
  if (a->type_ids[(size_t)a->cur_col] != DMB_TYPE_DECIMAL) return type_mismatch(a, "append_decimal");
  if (width < 1 || width > 38 || scale < 0 || scale > width) { fail(a, "append_decimal: DECIMAL(%d,%d) is not a type", width, scale); a->state = kError; return 0; }
  if (rb->dec_prec == 0) { rb->dec_prec = width; rb->dec_scale = scale; }  // first value declares the column
  __int128 v = (__int128)(((unsigned __int128)(uint64_t)upper << 64) | (unsigned __int128)(uint64_t)lower);
  auto pow10 = [](int e) { __int128 p = 1; while (e-- > 0) p *= 10; return p; };
  if (scale < rb->dec_scale) {
    const int up = rb->dec_scale - scale;
    const __int128 lim = pow10(38 - up);
    if (v >= lim || v <= -lim) { fail(a, "append_decimal: value out of range for DECIMAL(%d,%d)", rb->dec_prec, rb->dec_scale); a->state = kError; return 0; }
    v *= pow10(up);
  } else if (scale > rb->dec_scale) {
    const __int128 d = pow10(scale - rb->dec_scale), half = d / 2;
    v = v >= 0 ? (v + half) / d : -((-v + half) / d);
  }
  const __int128 bound = pow10(rb->dec_prec);
  if (v >= bound || v <= -bound) { fail(a, "append_decimal: value out of range for DECIMAL(%d,%d)", rb->dec_prec, rb->dec_scale); a->state = kError; return 0; }
  const uint8_t *p = reinterpret_cast<const uint8_t *>(&v);
  rb->values.insert(rb->values.end(), p, p + 16);
  rb->valid.push_back(1);
  a->cur_col++;
  return 1;
}

extern "C" int32_t duckdb_mb_gpu_append_bool(duckdb_mb_gpu_appender *a, int32_t v) {
  if (!a) return 0;
  RowBuf *rb = cell(a, "append_bool");
  if (!rb) return 0;
  if (a->type_ids[(size_t)a->cur_col] != DMB_TYPE_BOOLEAN) return type_mismatch(a, "append_bool");
  rb->values.push_back(v ? 1 : 0);
  rb->valid.push_back(1);
  a->cur_col++;
  return 1;
}

extern "C" int32_t duckdb_mb_gpu_append_varchar(duckdb_mb_gpu_appender *a, const uint8_t *bytes, int32_t len) {
  if (!a) return 0;
  RowBuf *rb = cell(a, "append_varchar");
  if (!rb) return 0;
  if (rb->width != 0) return type_mismatch(a, "append_varchar");
  if (len < 0 || (len > 0 && !bytes)) { fail(a, "append_varchar: bad buffer"); a->state = kError; return 0; }
  if ((int64_t)rb->data.size() + len > 0x7fffffffll) { fail(a, "append_varchar: more than 2 GiB of buffered string data; flush first"); a->state = kError; return 0; }
  rb->data.insert(rb->data.end(), bytes, bytes + len);
  rb->offsets.push_back((int32_t)rb->data.size());
  rb->valid.push_back(1);
  a->cur_col++;
  return 1;
}

extern "C" int32_t duckdb_mb_gpu_append_null(duckdb_mb_gpu_appender *a) {
  if (!a) return 0;
  RowBuf *rb = cell(a, "append_null");
  if (!rb) return 0;
  if (rb->width == 0) rb->offsets.push_back((int32_t)rb->data.size());
  else rb->values.resize(rb->values.size() + (size_t)rb->width, 0);
  rb->valid.push_back(0);
  a->cur_col++;
  return 1;
}

extern "C" int32_t duckdb_mb_gpu_end_row(duckdb_mb_gpu_appender *a) {
  if (!a) return 0;
  if (a->state != kRowInProgress) return illegal(a, "end_row");
  if (a->cur_col != a->ncols) {  // under-filled row (src/duckdb_appender_state_machine.mbt:160-177)
    fail(a, "end_row: row has %d of %d values", a->cur_col, a->ncols);
    a->state = kError;
    return 0;
  }
  a->state = kReady;
  a->row_count++;
  a->buffered_rows++;
  if (a->buffered_rows >= (1 << 20) && !flush_rows(a)) { a->state = kError; return 0; }
  return 1;
}

extern "C" int32_t duckdb_mb_gpu_appender_flush(duckdb_mb_gpu_appender *a) {
  if (!a) { set_error("null appender"); return 0; }
  if (a->state != kReady && a->state != kFlushed) return illegal(a, "flush");
  if (!flush_rows(a)) { a->state = kError; return 0; }
  if (a->on_flush && !a->on_flush(a->user)) { fail(a, "flush: the sink's flush hook failed"); a->state = kError; return 0; }
  a->flushed_row_count = a->row_count;
  a->state = kFlushed;
  return 1;
}

extern "C" int32_t duckdb_mb_gpu_appender_close(duckdb_mb_gpu_appender *a) {
  if (!a) { set_error("null appender"); return 0; }
  if (a->state == kClosed) return 1;  // Closed is terminal: the command is a no-op
  int32_t ok = 1;
  // complete buffered rows still reach the table, like duckdb_appender_destroy's implicit flush;
  // a row in progress is dropped
  if (a->state == kReady || a->state == kFlushed) ok = flush_rows(a);
  a->state = kClosed;
  return ok;
}

extern "C" int32_t duckdb_mb_gpu_appender_timings(duckdb_mb_gpu_appender *a, double *out4) {
  if (!a || !out4) return 0;
  for (int i = 0; i < 4; ++i) out4[i] = a->t[i];
  return 1;
}

extern "C" int32_t duckdb_mb_gpu_appender_link_bytes(duckdb_mb_gpu_appender *a, uint64_t *out2) {
  if (!a || !out2) return 0;
  out2[0] = a->bytes_h2d;
  out2[1] = a->bytes_d2h;
  return 1;
}

// =====================================================================================  L2 drop-in: the appender set
// The reference's own symbols (src/duckdb_native.c:1083-1251, 1313-1533, 1735-1926; MoonBit externs
// src/duckdb_native.mbt:44-110,134-144,165-217,266-288) with the reference's signatures: `duckdb_mb_appender` is this
// library's appender handle, Bytes arrive as moonbit_bytes_t (length from the object header, Moonbit_array_length),
// Array[Bytes] as moonbit_bytes_t*.  All parameters are #borrow: nothing is decref'd.  duckdb_mb_appender_create itself
// needs libduckdb (duckdb_appender_create) and lives in glue/duckdb_gpu_glue.c.
extern "C" void duckdb_mb_appender_destroy(duckdb_mb_appender *a) {  // :1083-1091
  if (!a) return;
  duckdb_mb_gpu_appender_close(a);  // complete rows still reach the table, like duckdb_appender_destroy's implicit flush
  duckdb_mb_gpu_appender_destroy(a);
}
extern "C" moonbit_bytes_t duckdb_mb_appender_error(duckdb_mb_appender *a) { return duckdb_mb_gpu_appender_error(a); }  // :1093-1098
extern "C" int32_t duckdb_mb_is_null_appender(duckdb_mb_appender *a) { return a == nullptr ? 1 : 0; }                    // :1253-1255
extern "C" int32_t duckdb_mb_begin_row(duckdb_mb_appender *a) { return duckdb_mb_gpu_begin_row(a); }                     // :1100
extern "C" int32_t duckdb_mb_append_int(duckdb_mb_appender *a, int32_t v) { return duckdb_mb_gpu_append_int(a, v); }     // :1116
extern "C" int32_t duckdb_mb_append_bigint(duckdb_mb_appender *a, int64_t v) { return duckdb_mb_gpu_append_bigint(a, v); }  // :1132
extern "C" int32_t duckdb_mb_append_double(duckdb_mb_appender *a, double v) { return duckdb_mb_gpu_append_double(a, v); }   // :1148
extern "C" int32_t duckdb_mb_append_varchar(duckdb_mb_appender *a, moonbit_bytes_t value) {                              // :1164
  if (!a) return 0;
  if (!value) { fail(a, "append_varchar: null Bytes"); return 0; }
  return duckdb_mb_gpu_append_varchar(a, value, Moonbit_array_length(value));
}
extern "C" int32_t duckdb_mb_append_bool(duckdb_mb_appender *a, bool v) { return duckdb_mb_gpu_append_bool(a, v ? 1 : 0); }  // :1189
extern "C" int32_t duckdb_mb_append_null(duckdb_mb_appender *a) { return duckdb_mb_gpu_append_null(a); }                 // :1205
extern "C" int32_t duckdb_mb_end_row(duckdb_mb_appender *a) { return duckdb_mb_gpu_end_row(a); }                         // :1221
extern "C" int32_t duckdb_mb_flush(duckdb_mb_appender *a) { return a ? duckdb_mb_gpu_appender_flush(a) : 0; }            // :1237
// exact days / micros into the vectors: the reference's approximate-string path (:1299-1367) is a defect, not a contract
extern "C" int32_t duckdb_mb_append_date(duckdb_mb_appender *a, int32_t days) { return duckdb_mb_gpu_append_date(a, days); }            // :1313
extern "C" int32_t duckdb_mb_append_timestamp(duckdb_mb_appender *a, int64_t micros) { return duckdb_mb_gpu_append_timestamp(a, micros); }  // :1350
extern "C" int32_t duckdb_mb_append_blob(duckdb_mb_appender *a, moonbit_bytes_t data, int32_t length) {                  // :1397
  if (!a) return 0;
  if (!data || length < 0 || length > Moonbit_array_length(data)) { fail(a, "append_blob: bad length"); return 0; }
  return duckdb_mb_gpu_append_blob(a, data, length);
}
extern "C" int32_t duckdb_mb_append_decimal(duckdb_mb_appender *a, uint8_t width, uint8_t scale, int64_t lower, int64_t upper) {  // :1447
  return duckdb_mb_gpu_append_decimal(a, width, scale, lower, upper);
}
extern "C" int32_t duckdb_mb_append_interval(duckdb_mb_appender *a, int32_t months, int32_t days, int64_t micros) {      // :1511
  return duckdb_mb_gpu_append_interval(a, months, days, micros);
}

namespace {
// Array[Bytes] (moonbit_bytes_t *) -> pointer + length tables
bool bytes_array(duckdb_mb_gpu_appender *a, const char *cmd, moonbit_bytes_t *items, int32_t count, std::vector<const uint8_t *> *ptrs,
                 std::vector<int32_t> *lens) {
  if (count < 0 || (count > 0 && !items)) { fail(a, "%s: bad arguments", cmd); return false; }
  ptrs->resize((size_t)count);
  lens->resize((size_t)count);
  for (int32_t i = 0; i < count; ++i) {
    if (!items[i]) { fail(a, "%s: element %d is null", cmd, i); return false; }
    (*ptrs)[(size_t)i] = items[i];
    (*lens)[(size_t)i] = Moonbit_array_length(items[i]);
  }
  return true;
}
}  // namespace

extern "C" int32_t duckdb_mb_append_list_varchar(duckdb_mb_appender *a, moonbit_bytes_t *values, int32_t count) {        // :1735
  if (!a) return 0;
  std::vector<const uint8_t *> p;
  std::vector<int32_t> l;
  if (!bytes_array(a, "append_list_varchar", values, count, &p, &l)) return 0;
  return duckdb_mb_gpu_append_list_varchar(a, p.data(), l.data(), count);
}
extern "C" int32_t duckdb_mb_append_struct_varchar(duckdb_mb_appender *a, moonbit_bytes_t *field_names, moonbit_bytes_t *field_values,
                                                  int32_t field_count) {                                                 // :1792
  if (!a) return 0;
  std::vector<const uint8_t *> pn, pv;
  std::vector<int32_t> ln, lv;
  if (!bytes_array(a, "append_struct_varchar", field_names, field_count, &pn, &ln) ||
      !bytes_array(a, "append_struct_varchar", field_values, field_count, &pv, &lv)) return 0;
  return duckdb_mb_gpu_append_struct_varchar(a, pn.data(), ln.data(), pv.data(), lv.data(), field_count);
}
extern "C" int32_t duckdb_mb_append_map_varchar_varchar(duckdb_mb_appender *a, moonbit_bytes_t *keys, moonbit_bytes_t *values,
                                                       int32_t entry_count) {                                            // :1860
  if (!a) return 0;
  std::vector<const uint8_t *> pk, pv;
  std::vector<int32_t> lk, lv;
  if (!bytes_array(a, "append_map_varchar_varchar", keys, entry_count, &pk, &lk) ||
      !bytes_array(a, "append_map_varchar_varchar", values, entry_count, &pv, &lv)) return 0;
  return duckdb_mb_gpu_append_map_varchar_varchar(a, pk.data(), lk.data(), pv.data(), lv.data(), entry_count);
}
