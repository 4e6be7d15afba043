"""DuckDB-shaped chunk batches (host side), deterministic by seed.

libduckdb is not available in this image, so inputs are built to the
layouts `duckdb.h` documents and the reference reads:

* a batch is a list of chunks of <= 2048 rows (``duckdb_data_chunk_get_size``,
  reference src/duckdb_native.c:510); every column has one flat vector per chunk;
* payload: dense array of the physical type (src/duckdb_native.c:553-662); vectors are allocated
  at full 2048-row capacity, so chunk k's payload sits at ``k * 2048 * width`` of the column slab;
* validity: ``uint64[32]`` per vector or a NULL pointer meaning all-valid
  (src/duckdb_native.c:530-533); payload under a NULL row is unspecified (filled with garbage);
* VARCHAR/BLOB: 16-byte ``duckdb_string_t`` — length <= 12 inlined (unused bytes zero), otherwise
  4-byte prefix + 8-byte *host pointer* into a string heap (src/duckdb_native.c:597-603).  The
  pointers here are real addresses into ``heap`` so the C oracle dereferences them exactly as the
  reference does.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

VECTOR_SIZE = 2048
VALIDITY_WORDS = 32

# DUCKDB_TYPE ids (include/duckdb_mb_gpu.h enum dmb_type; reference src/duckdb_parsing.mbt:8-52)
T_INVALID, T_BOOLEAN, T_TINYINT, T_SMALLINT, T_INTEGER, T_BIGINT = 0, 1, 2, 3, 4, 5
T_UTINYINT, T_USMALLINT, T_UINTEGER, T_UBIGINT, T_FLOAT, T_DOUBLE = 6, 7, 8, 9, 10, 11
T_TIMESTAMP, T_DATE, T_TIME, T_INTERVAL, T_HUGEINT, T_VARCHAR, T_BLOB, T_DECIMAL = 12, 13, 14, 15, 16, 17, 18, 19
T_TIMESTAMP_S, T_TIMESTAMP_MS, T_TIMESTAMP_NS = 20, 21, 22
T_ENUM = 23
T_LIST = 24
T_STRUCT, T_MAP = 25, 26
T_UUID, T_TIME_TZ, T_TIMESTAMP_TZ, T_UHUGEINT, T_TIME_NS = 27, 30, 31, 32, 39

# enum dmb_phys
P_BOOL, P_I8, P_I16, P_I32, P_I64, P_U8, P_U16, P_U32, P_U64, P_F32, P_F64 = range(11)
P_I128, P_U128, P_INTERVAL, P_STRING = 11, 12, 13, 14
PHYS_WIDTH = [1, 1, 2, 4, 8, 1, 2, 4, 8, 4, 8, 16, 16, 16, 16]
PHYS_NUMPY = {
    P_BOOL: np.uint8, P_I8: np.int8, P_I16: np.int16, P_I32: np.int32, P_I64: np.int64,
    P_U8: np.uint8, P_U16: np.uint16, P_U32: np.uint32, P_U64: np.uint64,
    P_F32: np.float32, P_F64: np.float64,
}

# enum dmb_dst
(D_SAME, D_I32_TRUNC, D_I64, D_F64, D_BOOL_BYTE, D_BOOL_BITS, D_I128, D_I32_SAT,
 D_TS_US_FROM_S, D_TS_US_FROM_MS, D_TS_US_FROM_NS, D_MONTH_DAY_NANO, D_DATE_REF,
 D_TS_REF, D_TS_REF_FROM_S, D_TS_REF_FROM_MS, D_TS_REF_FROM_NS,
 D_DEC_I64, D_DEC_I32_TRUNC, D_DEC_F64, D_DEC_BOOL_BYTE) = range(21)
OP_VALIDITY_ONLY = 0x7F00


def op(phys: int, dst: int) -> int:
    return (phys << 8) | dst


def phys_of_type(type_id: int, dec_width: int = 0) -> int:
    """Physical vector type of a logical DuckDB type."""
    if type_id == T_DECIMAL:
        return P_I16 if dec_width <= 4 else P_I32 if dec_width <= 9 else P_I64 if dec_width <= 18 else P_I128
    return {
        T_BOOLEAN: P_BOOL, T_TINYINT: P_I8, T_SMALLINT: P_I16, T_INTEGER: P_I32, T_BIGINT: P_I64,
        T_UTINYINT: P_U8, T_USMALLINT: P_U16, T_UINTEGER: P_U32, T_UBIGINT: P_U64,
        T_FLOAT: P_F32, T_DOUBLE: P_F64, T_TIMESTAMP: P_I64, T_DATE: P_I32, T_TIME: P_I64,
        T_INTERVAL: P_INTERVAL, T_HUGEINT: P_I128, T_VARCHAR: P_STRING, T_BLOB: P_STRING,
        T_TIMESTAMP_S: P_I64, T_TIMESTAMP_MS: P_I64, T_TIMESTAMP_NS: P_I64, T_UUID: P_U128,
        T_TIME_TZ: P_U64, T_TIMESTAMP_TZ: P_I64, T_UHUGEINT: P_U128, T_TIME_NS: P_I64,
    }[type_id]


@dataclass
class Column:
    """One column of a chunk batch: per-chunk flat vectors inside contiguous slabs."""
    name: str
    type_id: int
    phys: int
    data: np.ndarray                 # uint8 slab
    data_off: np.ndarray             # uint64 [nchunks] byte offset of each vector payload
    validity: Optional[np.ndarray]   # uint64 slab or None
    val_off: np.ndarray              # int64 [nchunks] word offset of each mask, -1 = NULL pointer
    dec_width: int = 0
    dec_scale: int = 0
    heap: Optional[np.ndarray] = None  # uint8 string heap (VARCHAR/BLOB)
    inline_only: bool = False          # VARCHAR/BLOB: every string is inlined (<= 12 bytes): DMB_HEAP_INLINE_ONLY
    dictionary: Optional[List[bytes]] = None  # ENUM: the type's labels (duckdb_enum_dictionary_value), indices in `data`
    # LIST: `data` holds duckdb_list_entry {uint64 offset, uint64 length}; per chunk the child vector the entries index
    list_child_type: int = 0
    list_child_dec_width: int = 0
    list_child_dec_scale: int = 0
    list_child_data: Optional[np.ndarray] = None      # uint8: the child vectors of all chunks back to back
    list_child_base: Optional[np.ndarray] = None      # uint64 [nchunks]: first element of chunk k's child vector
    list_child_sizes: Optional[np.ndarray] = None     # uint64 [nchunks]: duckdb_list_vector_get_size
    list_child_validity: Optional[np.ndarray] = None  # uint64 words
    list_child_val_off: Optional[np.ndarray] = None   # int64 [nchunks], -1 = NULL mask pointer
    # nested types (dmb_host_struct / dmb_host_list.child_col).  A nested child is itself a Column whose vector of chunk k
    # sits at data_off[k] / val_off[k] of its own slabs; under a LIST its chunk k vector holds list_child_sizes[k] elements.
    struct_fields: Optional[List["Column"]] = None    # STRUCT: one Column per field (same rows as the parent)
    list_child_col: Optional["Column"] = None         # LIST / MAP: the child described as a Column

    @property
    def width(self) -> int:
        return PHYS_WIDTH[self.phys]

    @property
    def heap_base(self) -> int:
        return 0 if self.heap is None else int(self.heap.ctypes.data)


@dataclass
class ChunkBatch:
    counts: np.ndarray               # uint32 [nchunks]
    columns: List[Column] = field(default_factory=list)

    @property
    def nchunks(self) -> int:
        return int(self.counts.shape[0])

    @property
    def nrows(self) -> int:
        return int(self.counts.sum(dtype=np.int64))

    @property
    def row_off(self) -> np.ndarray:
        ro = np.zeros(self.nchunks + 1, dtype=np.int64)
        np.cumsum(self.counts, dtype=np.int64, out=ro[1:])
        return ro


def chunk_counts(nrows: int, pattern: str = "full", rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """Row counts per chunk.  'full': 2048,...,remainder.  'ragged': filtered-scan style short
    chunks (including empty ones) that put chunk boundaries at arbitrary bit offsets."""
    if nrows <= 0:
        return np.zeros(0, dtype=np.uint32)
    if pattern == "full":
        n_full, rem = divmod(nrows, VECTOR_SIZE)
        c = [VECTOR_SIZE] * n_full + ([rem] if rem else [])
        return np.asarray(c, dtype=np.uint32)
    assert rng is not None
    out, left = [], nrows
    while left > 0:
        r = rng.random()
        if r < 0.08:
            c = 0
        elif r < 0.3:
            c = int(rng.integers(1, 70))
        elif r < 0.6:
            c = int(rng.integers(1, VECTOR_SIZE + 1))
        else:
            c = VECTOR_SIZE
        c = min(c, left)
        out.append(c)
        left -= c
    return np.asarray(out, dtype=np.uint32)


def _slab_offsets(nchunks: int, width: int) -> np.ndarray:
    return (np.arange(nchunks, dtype=np.uint64) * np.uint64(VECTOR_SIZE * width))


def _scatter_rows(values: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """Contiguous per-row values -> slab with one 2048-row slot per chunk (tail of a short chunk
    is left as zero, like unused vector capacity)."""
    nchunks = counts.shape[0]
    slab = np.zeros((nchunks * VECTOR_SIZE,) + values.shape[1:], dtype=values.dtype)
    if counts.size and np.all(counts[:-1] == VECTOR_SIZE):
        slab[: values.shape[0]] = values
        return slab
    ro = np.zeros(nchunks + 1, dtype=np.int64)
    np.cumsum(counts, dtype=np.int64, out=ro[1:])
    for k in range(nchunks):
        c = int(counts[k])
        if c:
            slab[k * VECTOR_SIZE: k * VECTOR_SIZE + c] = values[ro[k]: ro[k] + c]
    return slab


def make_validity(valid: Optional[np.ndarray], counts: np.ndarray, null_ptr_when_all_valid: bool = True):
    """Per-row bool validity -> (uint64 slab | None, val_off[nchunks]).  Chunks with no NULL get
    a NULL pointer (-1) like DuckDB vectors that never had their mask materialised."""
    nchunks = counts.shape[0]
    val_off = np.full(nchunks, -1, dtype=np.int64)
    if valid is None:
        return None, val_off
    bits = _scatter_rows(valid.astype(np.uint8), counts)  # capacity bits beyond count: 0
    packed = np.packbits(bits.reshape(nchunks, VECTOR_SIZE), axis=1, bitorder="little")  # [nchunks, 256]
    slab = np.ascontiguousarray(packed).view(np.uint64).reshape(-1)
    per_chunk_valid = bits.reshape(nchunks, VECTOR_SIZE).sum(axis=1, dtype=np.int64)
    for k in range(nchunks):
        if null_ptr_when_all_valid and per_chunk_valid[k] == int(counts[k]):
            continue
        val_off[k] = k * VALIDITY_WORDS
    return slab, val_off


def fixed_column(name: str, type_id: int, values: np.ndarray, counts: np.ndarray,
                 valid: Optional[np.ndarray] = None, dec_width: int = 0, dec_scale: int = 0,
                 garbage_rng: Optional[np.random.Generator] = None,
                 null_ptr_when_all_valid: bool = True) -> Column:
    """values: per-row array of the physical numpy dtype, or uint8[n,16] for 16-byte types."""
    phys = phys_of_type(type_id, dec_width)
    width = PHYS_WIDTH[phys]
    raw = np.ascontiguousarray(values).view(np.uint8).reshape(values.shape[0], width).copy()
    if valid is not None and garbage_rng is not None:
        nulls = ~valid
        k = int(nulls.sum())
        if k:
            raw[nulls] = garbage_rng.integers(0, 256, size=(k, width), dtype=np.uint8)
            if phys == P_BOOL:
                raw[nulls] &= 1
    slab = _scatter_rows(raw, counts).reshape(-1)
    vslab, val_off = make_validity(valid, counts, null_ptr_when_all_valid)
    return Column(name, type_id, phys, slab, _slab_offsets(counts.shape[0], width), vslab, val_off,
                  dec_width, dec_scale)


def enum_column(name: str, labels: Sequence[bytes], indices: np.ndarray, counts: np.ndarray,
                valid: Optional[np.ndarray] = None, garbage_rng: Optional[np.random.Generator] = None,
                null_ptr_when_all_valid: bool = True) -> Column:
    """ENUM column: indices in the width duckdb_enum_internal_type picks (uint8 up to 256 labels, uint16 up to
    65536, else uint32) + the type's dictionary."""
    n = len(labels)
    dt, type_phys = (np.uint8, T_UTINYINT) if n <= 256 else (np.uint16, T_USMALLINT) if n <= 65536 else (np.uint32, T_UINTEGER)
    col = fixed_column(name, type_phys, np.asarray(indices).astype(dt), counts, valid, garbage_rng=garbage_rng,
                       null_ptr_when_all_valid=null_ptr_when_all_valid)
    col.type_id = T_ENUM
    col.dictionary = [bytes(x) for x in labels]
    return col


def enum_dict_arrays(labels: Sequence[bytes]):
    """(uint32 offsets [n+1], uint8 data) of a dictionary: the dmb_enum_dict layout"""
    offs = np.zeros(len(labels) + 1, dtype=np.uint32)
    np.cumsum([len(x) for x in labels], out=offs[1:])
    data = np.frombuffer(b"".join(labels) + b"\0", dtype=np.uint8).copy()
    return offs, data


def string_column(name: str, strings: Sequence[Optional[bytes]], counts: np.ndarray,
                  type_id: int = T_VARCHAR, shuffle_heap: Optional[np.random.Generator] = None) -> Column:
    """Build real duckdb_string_t entries from python bytes (None = NULL). Small inputs only."""
    n = len(strings)
    lens = np.asarray([0 if s is None else len(s) for s in strings], dtype=np.int64)
    valid = np.asarray([s is not None for s in strings], dtype=bool)
    order = np.arange(n)
    if shuffle_heap is not None:
        order = shuffle_heap.permutation(n)
    heap_off = np.zeros(n, dtype=np.int64)
    pos = 0
    for i in order:
        if lens[i] > 12:
            heap_off[i] = pos
            pos += int(lens[i])
    heap = np.zeros(pos + 16, dtype=np.uint8)
    for i in range(n):
        if lens[i] > 12:
            heap[heap_off[i]: heap_off[i] + lens[i]] = np.frombuffer(strings[i], dtype=np.uint8)
    entries = np.zeros((n, 16), dtype=np.uint8)
    base = int(heap.ctypes.data)
    for i in range(n):
        s = strings[i]
        if s is None:
            # NULL row: DuckDB leaves the slot unspecified; use a recognisable non-empty garbage entry
            entries[i, 0:4] = np.frombuffer(np.uint32(7).tobytes(), dtype=np.uint8)
            entries[i, 4:11] = np.frombuffer(b"garbage", dtype=np.uint8)
            continue
        L = len(s)
        entries[i, 0:4] = np.frombuffer(np.uint32(L).tobytes(), dtype=np.uint8)
        if L <= 12:
            entries[i, 4:4 + L] = np.frombuffer(s, dtype=np.uint8)
        else:
            entries[i, 4:8] = np.frombuffer(s[:4], dtype=np.uint8)
            entries[i, 8:16] = np.frombuffer(np.uint64(base + int(heap_off[i])).tobytes(), dtype=np.uint8)
    slab = _scatter_rows(entries, counts).reshape(-1)
    vslab, val_off = make_validity(valid if not valid.all() else None, counts)
    col = Column(name, type_id, P_STRING, slab, _slab_offsets(counts.shape[0], 16), vslab, val_off)
    col.heap = heap
    return col


def string_column_bulk(name: str, lens: np.ndarray, valid: Optional[np.ndarray], counts: np.ndarray,
                       rng: np.random.Generator, utf8_fraction: float = 0.0,
                       alphabet: Optional[np.ndarray] = None) -> Column:
    """Vectorised VARCHAR column: random printable ASCII (optionally some 2-3 byte UTF-8 code
    points), lengths given, pointer strings laid out in row order in one contiguous heap."""
    n = lens.shape[0]
    lens = lens.astype(np.int64)
    total = int(lens.sum())
    if alphabet is None:
        body = rng.integers(0x20, 0x7F, size=total + 16, dtype=np.uint8)
    else:
        body = alphabet[rng.integers(0, alphabet.shape[0], size=total + 16)]
    starts = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=starts[1:])
    if utf8_fraction > 0 and n:
        # overwrite the head of some rows with a well-formed multi-byte sequence
        pick = np.nonzero((rng.random(n) < utf8_fraction) & (lens >= 3))[0]
        three = rng.random(pick.shape[0]) < 0.5
        p3, p2 = pick[three], pick[~three]
        body[starts[p3]] = 0xE3; body[starts[p3] + 1] = 0x81; body[starts[p3] + 2] = 0x82  # U+3042
        body[starts[p2]] = 0xC3; body[starts[p2] + 1] = 0xA9                                # U+00E9
    # heap holds only the non-inlined strings, row order
    is_ptr = lens > 12
    heap_lens = np.where(is_ptr, lens, 0)
    heap_starts = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(heap_lens, out=heap_starts[1:])
    heap = np.zeros(int(heap_starts[-1]) + 16, dtype=np.uint8)
    if is_ptr.any():
        # gather bytes of pointer strings: positions in body
        idx = np.repeat(starts[:-1][is_ptr] - heap_starts[:-1][is_ptr], lens[is_ptr]) + np.arange(int(heap_starts[-1]), dtype=np.int64)
        heap[: int(heap_starts[-1])] = body[idx]
    entries = np.zeros((n, 16), dtype=np.uint8)
    entries[:, 0:4] = lens.astype(np.uint32).view(np.uint8).reshape(n, 4)
    # first 12 bytes of each string (masked by length) for inline; first 4 for prefix
    for k in range(12):
        take = lens > k
        if k >= 4:
            take &= ~is_ptr
        entries[take, 4 + k] = body[starts[:-1][take] + k]
    base = int(heap.ctypes.data)
    ptrs = (np.uint64(base) + heap_starts[:-1].astype(np.uint64))
    entries[is_ptr, 8:16] = ptrs[is_ptr].view(np.uint8).reshape(-1, 8)
    if valid is not None:
        nulls = ~valid
        k = int(nulls.sum())
        if k:
            g = np.zeros((k, 16), dtype=np.uint8)
            g[:, 0] = 5
            g[:, 4:9] = rng.integers(0x41, 0x5B, size=(k, 5), dtype=np.uint8)
            entries[nulls] = g
    slab = _scatter_rows(entries, counts).reshape(-1)
    vslab, val_off = make_validity(valid, counts)
    col = Column(name, T_VARCHAR, P_STRING, slab, _slab_offsets(counts.shape[0], 16), vslab, val_off)
    col.heap = heap
    return col


def string_values(col: Column, counts: np.ndarray) -> List[Optional[bytes]]:
    """Decode a string column back to python bytes by following the real pointers (tests only)."""
    import ctypes
    out: List[Optional[bytes]] = []
    ent = col.data.reshape(-1, 16)
    for k in range(counts.shape[0]):
        for r in range(int(counts[k])):
            if col.val_off[k] >= 0:
                w = col.validity[col.val_off[k] + r // 64]
                if not (int(w) >> (r % 64)) & 1:
                    out.append(None)
                    continue
            e = ent[int(col.data_off[k]) // 16 + r]
            L = int(e[0:4].view(np.uint32)[0])
            if L <= 12:
                out.append(bytes(e[4:4 + L]))
            else:
                p = int(e[8:16].view(np.uint64)[0])
                out.append(ctypes.string_at(p, L))
    return out


# ----------------------------------------------------------------------------------------------
# BASELINE.json configs, seed = 20260101 + config number
# ----------------------------------------------------------------------------------------------

def config_c1(nrows: int = 1_000_000, variant_b: bool = False) -> ChunkBatch:
    """C1: SELECT i::INTEGER, i::DOUBLE, CASE WHEN i%7=0 THEN NULL END FROM range(n).
    Third column: every row NULL (typed INTEGER); variant_b = `... ELSE i END` (1/7 NULL)."""
    counts = chunk_counts(nrows)
    i = np.arange(nrows, dtype=np.int64)
    cols = [
        fixed_column("CAST(i AS INTEGER)", T_INTEGER, i.astype(np.int32), counts),
        fixed_column("CAST(i AS DOUBLE)", T_DOUBLE, i.astype(np.float64), counts),
    ]
    if variant_b:
        valid = (i % 7) != 0
        cols.append(fixed_column("CASE", T_BIGINT, i.copy(), counts, valid=valid))
    else:
        valid = np.zeros(nrows, dtype=bool)
        cols.append(fixed_column("CASE", T_INTEGER, np.zeros(nrows, dtype=np.int32), counts, valid=valid,
                                 null_ptr_when_all_valid=False))
    return ChunkBatch(counts, cols)


LINEITEM_SHIPINSTRUCT = [b"DELIVER IN PERSON", b"COLLECT COD", b"NONE", b"TAKE BACK RETURN"]
LINEITEM_SHIPMODE = [b"REG AIR", b"AIR", b"RAIL", b"SHIP", b"TRUCK", b"MAIL", b"FOB"]


def _dict_string_column(name, choices, picks, counts, rng):
    lens = np.asarray([len(c) for c in choices], dtype=np.int64)[picks]
    n = picks.shape[0]
    col = string_column_bulk(name, lens, None, counts, rng)
    # overwrite random bytes with the dictionary words (inline entries and heap)
    ent = _gather_entries(col, counts)
    maxlen = max(len(c) for c in choices)
    table = np.zeros((len(choices), maxlen), dtype=np.uint8)
    for j, c in enumerate(choices):
        table[j, : len(c)] = np.frombuffer(c, dtype=np.uint8)
    is_ptr = lens > 12
    for k in range(min(12, maxlen)):
        take = lens > k
        if k >= 4:
            take &= ~is_ptr
        ent[take, 4 + k] = table[picks[take], k]
    if is_ptr.any():
        heap_lens = np.where(is_ptr, lens, 0)
        hs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(heap_lens, out=hs[1:])
        rows = np.repeat(np.nonzero(is_ptr)[0], lens[is_ptr])
        within = np.arange(int(hs[-1]), dtype=np.int64) - np.repeat(hs[:-1][is_ptr], lens[is_ptr])
        col.heap[: int(hs[-1])] = table[picks[rows], within]
    _scatter_entries(col, ent, counts)
    return col


def _gather_entries(col: Column, counts: np.ndarray) -> np.ndarray:
    ent = col.data.reshape(-1, 16)
    if np.all(counts[:-1] == VECTOR_SIZE):
        return ent[: int(counts.sum())].copy()
    parts = [ent[k * VECTOR_SIZE: k * VECTOR_SIZE + int(c)] for k, c in enumerate(counts)]
    return np.concatenate(parts) if parts else ent[:0].copy()


def _scatter_entries(col: Column, ent: np.ndarray, counts: np.ndarray) -> None:
    col.data = _scatter_rows(ent, counts).reshape(-1)


def config_c2(nrows: int = 60_000_000, seed: int = 20260103) -> ChunkBatch:
    """C2: TPC-H lineitem shape (4 INTEGER, 4 DECIMAL(15,2), 3 DATE, 5 VARCHAR), no NULLs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    counts = chunk_counts(nrows)
    cols: List[Column] = []
    for nm in ("l_orderkey", "l_partkey", "l_suppkey", "l_linenumber"):
        cols.append(fixed_column(nm, T_INTEGER, rng.integers(0, 2**31 - 1, size=nrows, dtype=np.int32), counts))
    for nm in ("l_quantity", "l_extendedprice", "l_discount", "l_tax"):
        cols.append(fixed_column(nm, T_DECIMAL, rng.integers(0, 10**7, size=nrows, dtype=np.int64), counts,
                                 dec_width=15, dec_scale=2))
    for nm in ("l_shipdate", "l_commitdate", "l_receiptdate"):
        cols.append(fixed_column(nm, T_DATE, rng.integers(8035, 10592, size=nrows, dtype=np.int32), counts))
    cols.append(_dict_string_column("l_returnflag", [b"A", b"N", b"R"], rng.integers(0, 3, size=nrows), counts, rng))
    cols.append(_dict_string_column("l_linestatus", [b"F", b"O"], rng.integers(0, 2, size=nrows), counts, rng))
    cols.append(_dict_string_column("l_shipinstruct", LINEITEM_SHIPINSTRUCT, rng.integers(0, 4, size=nrows), counts, rng))
    cols.append(_dict_string_column("l_shipmode", LINEITEM_SHIPMODE, rng.integers(0, 7, size=nrows), counts, rng))
    lens = rng.integers(10, 44, size=nrows)
    cols.append(string_column_bulk("l_comment", lens, None, counts, rng,
                                   alphabet=np.frombuffer(b"abcdefghijklmnopqrstuvwxyz ", dtype=np.uint8)))
    return ChunkBatch(counts, cols)


def config_c3(nrows: int = 100_000_000, seed: int = 20260104, pattern: str = "full") -> ChunkBatch:
    """C3: one VARCHAR, len U[0,64], 10% NULL, 5% rows with multi-byte UTF-8."""
    rng = np.random.Generator(np.random.PCG64(seed))
    counts = chunk_counts(nrows, pattern, rng)
    lens = rng.integers(0, 65, size=nrows)
    valid = rng.random(nrows) >= 0.10
    return ChunkBatch(counts, [string_column_bulk("s", lens, valid, counts, rng, utf8_fraction=0.05)])


def config_c4(nrows: int = 10_000_000, seed: int = 20260105, ncols: int = 64, pattern: str = "full") -> ChunkBatch:
    """C4: 22 TIMESTAMP, 21 DECIMAL(18,3), 21 HUGEINT; 30% NULL i.i.d.; garbage under NULLs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    null_rng = np.random.Generator(np.random.PCG64(seed + 1000))
    counts = chunk_counts(nrows, pattern, rng)
    cols: List[Column] = []
    kinds = (["ts"] * 22 + ["dec"] * 21 + ["huge"] * 21)[:ncols] if ncols == 64 else \
        [("ts", "dec", "huge")[j % 3] for j in range(ncols)]
    for j, kind in enumerate(kinds):
        valid = null_rng.random(nrows) >= 0.30
        if kind == "ts":
            v = rng.integers(0, 2 * 10**15, size=nrows, dtype=np.int64)
            cols.append(fixed_column(f"ts{j}", T_TIMESTAMP, v, counts, valid=valid, garbage_rng=null_rng))
        elif kind == "dec":
            v = rng.integers(-(10**18 - 1), 10**18, size=nrows, dtype=np.int64)
            cols.append(fixed_column(f"dec{j}", T_DECIMAL, v, counts, valid=valid, dec_width=18, dec_scale=3,
                                     garbage_rng=null_rng))
        else:
            v = rng.integers(0, 256, size=(nrows, 16), dtype=np.uint8)
            cols.append(fixed_column(f"huge{j}", T_HUGEINT, v, counts, valid=valid, garbage_rng=null_rng))
    return ChunkBatch(counts, cols)
