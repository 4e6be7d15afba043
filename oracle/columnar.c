/*
 * columnar.c — the strong CPU line of the benchmark: DataChunk -> Arrow, columnar, all host cores.
 * TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and tests/); the product never links it.
 *
 * SURVEY.md §8d / BASELINE.md §2.2 ask for two CPU numbers: the reference-equivalent one (oracle.c: the
 * reference's per-cell loops, src/duckdb_native.c:2357-2797, one thread like the reference) and "an -O3
 * -march=native OpenMP columnar version on all host cores as the honest strong-CPU line".  This file is the
 * second: the same conversions the GPU library does (chunk vectors -> dense Arrow buffers), written the way a
 * careful CPU implementation would — chunk-parallel memcpy of payloads, word copies of validity masks, a
 * two-pass (sum, scan, copy) utf8 gather — with no per-cell function calls.  Its outputs are checked against
 * oracle.c in tests/test_columnar_cpu.py, so the number it produces is for a correct conversion.
 *
 * Layouts: chunk k of a column is at data + data_off[k]; its mask (uint64[32]) at validity + val_off[k] or
 * val_off[k] < 0 for "all valid" (src/duckdb_native.c:530-533); duckdb_string_t as read at :597-603.
 */
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VS 2048

typedef struct {
  int32_t type_id, phys, dec_width, dec_scale;
  const uint8_t *data;
  const uint64_t *data_off;
  const uint64_t *validity;
  const int64_t *val_off;
  const char *name;
  const uint32_t *dict_offsets;
  const char *dict_data;
  uint32_t dict_size;
} col_column; /* = ora_column of oracle.c */

typedef struct {
  int64_t nchunks;
  const uint32_t *counts;
  int32_t ncols;
  const col_column *cols;
} col_batch; /* = ora_batch */

enum { COL_SAME = 0, COL_WIDEN128 = 6 };

static inline int bit(const uint64_t *m, uint32_t i) { return (int)((m[i >> 6] >> (i & 63)) & 1ull); }

/* validity of chunk k -> Arrow bitmap at row0; returns the chunk's null count */
static int64_t put_validity(uint8_t *bitmap, int64_t row0, const uint64_t *mask, uint32_t count) {
  int64_t nulls = 0;
  if ((row0 & 7) == 0 && (count & 7) == 0) { /* byte aligned: a copy (LSB-first on both sides) */
    if (!mask) memset(bitmap + (row0 >> 3), 0xff, count >> 3);
    else {
      memcpy(bitmap + (row0 >> 3), mask, count >> 3);
      for (uint32_t w = 0; w < (count + 63) / 64; w++) {
        uint64_t word = mask[w];
        uint32_t live = count - w * 64 < 64 ? count - w * 64 : 64;
        if (live < 64) word &= (1ull << live) - 1ull;
        nulls += live - (uint32_t)__builtin_popcountll(word);
      }
    }
    return nulls;
  }
  for (uint32_t i = 0; i < count; i++) { /* ragged chunk boundary: bit by bit (two chunks may share a byte: see the caller) */
    int v = mask ? bit(mask, i) : 1;
    int64_t r = row0 + i;
    if (v) __atomic_fetch_or(&bitmap[r >> 3], (uint8_t)(1u << (r & 7)), __ATOMIC_RELAXED);
    else nulls++;
  }
  return nulls;
}

/* fixed-width column -> dense values + bitmap.  dst COL_SAME (width w) or COL_WIDEN128 (int16/32/64 -> int128).
 * NULL slots are zeroed (the convention of every export here).  bitmap must be zeroed by the caller. */
int64_t col_fixed(const col_batch *b, int32_t c, int dst, const int64_t *row_off, uint8_t *out, uint8_t *bitmap, int threads) {
  const col_column *col = &b->cols[c];
  static const int W[] = {1, 1, 2, 4, 8, 1, 2, 4, 8, 4, 8, 16, 16, 16, 16};
  const int w = W[col->phys];
  int64_t nulls = 0;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : nulls)
  for (int64_t k = 0; k < b->nchunks; k++) {
    const uint32_t count = b->counts[k];
    if (!count) continue;
    const uint8_t *in = col->data + col->data_off[k];
    const uint64_t *mask = col->val_off[k] < 0 ? NULL : col->validity + col->val_off[k];
    const int64_t row0 = row_off[k];
    if (dst == COL_SAME) {
      uint8_t *o = out + (size_t)row0 * (size_t)w;
      memcpy(o, in, (size_t)count * (size_t)w);
      if (mask)
        for (uint32_t i = 0; i < count; i++)
          if (!bit(mask, i)) memset(o + (size_t)i * (size_t)w, 0, (size_t)w);
    } else {
      __int128 *o = (__int128 *)out + row0;
      if (w == 8) { const int64_t *v = (const int64_t *)in; for (uint32_t i = 0; i < count; i++) o[i] = (!mask || bit(mask, i)) ? (__int128)v[i] : 0; }
      else if (w == 4) { const int32_t *v = (const int32_t *)in; for (uint32_t i = 0; i < count; i++) o[i] = (!mask || bit(mask, i)) ? (__int128)v[i] : 0; }
      else { const int16_t *v = (const int16_t *)in; for (uint32_t i = 0; i < count; i++) o[i] = (!mask || bit(mask, i)) ? (__int128)v[i] : 0; }
    }
    if (bitmap) nulls += put_validity(bitmap, row0, mask, count);
  }
  return nulls;
}

/* VARCHAR column -> utf8 offsets (int32, or int64 when large) + data + bitmap.  chunk_base: scratch [nchunks + 1].
 * Returns the total data length, or -1 when int32 offsets overflow. */
int64_t col_string(const col_batch *b, int32_t c, int large, const int64_t *row_off, void *offsets, uint8_t *data,
                   uint8_t *bitmap, int64_t *chunk_base, int64_t *null_count, int threads) {
  const col_column *col = &b->cols[c];
  typedef struct { uint32_t length; char rest[12]; } str_t;
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int64_t k = 0; k < b->nchunks; k++) { /* pass 1: bytes per chunk */
    const str_t *v = (const str_t *)(col->data + col->data_off[k]);
    const uint64_t *mask = col->val_off[k] < 0 ? NULL : col->validity + col->val_off[k];
    int64_t sum = 0;
    for (uint32_t i = 0; i < b->counts[k]; i++)
      if (!mask || bit(mask, i)) sum += v[i].length;
    chunk_base[k + 1] = sum;
  }
  chunk_base[0] = 0;
  for (int64_t k = 0; k < b->nchunks; k++) chunk_base[k + 1] += chunk_base[k];
  const int64_t total = chunk_base[b->nchunks];
  if (!large && total > 0x7fffffffll) return -1;
  int64_t nulls = 0;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : nulls)
  for (int64_t k = 0; k < b->nchunks; k++) { /* pass 2: offsets + bytes */
    const uint32_t count = b->counts[k];
    const str_t *v = (const str_t *)(col->data + col->data_off[k]);
    const uint64_t *mask = col->val_off[k] < 0 ? NULL : col->validity + col->val_off[k];
    int64_t pos = chunk_base[k];
    const int64_t row0 = row_off[k];
    for (uint32_t i = 0; i < count; i++) {
      if (large) ((int64_t *)offsets)[row0 + i] = pos; else ((int32_t *)offsets)[row0 + i] = (int32_t)pos;
      if (mask && !bit(mask, i)) continue;
      const uint32_t len = v[i].length;
      const char *src = v[i].rest;
      if (len > 12) memcpy(&src, v[i].rest + 4, sizeof(src));
      memcpy(data + pos, src, len);
      pos += len;
    }
    if (bitmap) nulls += put_validity(bitmap, row0, mask, count);
  }
  int64_t nrows = b->nchunks ? row_off[b->nchunks] : 0;
  if (large) ((int64_t *)offsets)[nrows] = total; else ((int32_t *)offsets)[nrows] = (int32_t)total;
  if (null_count) *null_count = nulls;
  return total;
}

int col_max_threads(void) { return omp_get_max_threads(); }
