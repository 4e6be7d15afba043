"""Host-side mirror of the reference's `duckdb_arrow_native` surface (src/duckdb_arrow_native.mbt),
served by libduckdb_mb_gpu.so.

`ArrowResult` has the reference's method names and semantics (`column_count`, `row_count`,
`get_schema`, `get_column_{int32,int64,double,string,bool}[_nullable]`, `close`); the blobs come
from the drop-in C symbols `duckdb_mb_arrow_*` (L2) and are decoded with the reference's own
decoder rules (count cap 1 000 000, int64 read as a 32-bit Int, NUL-scanned strings,
src/duckdb_arrow_native.mbt:430-822).  `to_arrow()` is the additive true-Arrow export (L1,
Arrow C Data Interface) imported zero-copy into pyarrow.

Where the reference runs SQL (`Connection::query_arrow`, :123-135) this mirror takes the
DataChunks the query produced (`ArrowResult.from_chunks`): libduckdb is not in this image.
No CPU fallback: every call goes through the CUDA library or raises.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import json
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import chunks as ch
from . import native as nat

DECODER_ROW_CAP = 1_000_000  # src/duckdb_arrow_native.mbt:435 (and every other decoder)


class DuckDBError(RuntimeError):
    """DuckDBError::Message of the reference (src/duckdb.mbt)."""


@dataclass
class ArrowField:  # src/duckdb_arrow_native.mbt:111-115
    name: str
    nullable: bool
    type_id: str


@dataclass
class ArrowSchemaInfo:  # :118-120
    fields: List[ArrowField]


class GpuContext:
    """One per GPU: streams, pinned staging ring, buffer pools (duckdb_mb_gpu_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = nat.lib()
        self.handle = self.lib.duckdb_mb_gpu_ctx_create(device)
        if not self.handle:
            raise DuckDBError(nat.last_error() or "duckdb_mb_gpu_ctx_create failed")
        self.device = device

    def sync(self) -> None:
        if not self.lib.duckdb_mb_gpu_ctx_sync(self.handle):
            raise DuckDBError(nat.last_error())

    def close(self) -> None:
        if self.handle:
            self.lib.duckdb_mb_gpu_ctx_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class HostBatch:
    """ctypes view (dmb_host_batch) of a host ChunkBatch: per column the per-chunk pointers that
    duckdb_vector_get_data / duckdb_vector_get_validity return (src/duckdb_native.c:529-530,547)."""

    def __init__(self, batch, pinned: bool = False, register_heap: bool = True):
        self.batch = batch  # keeps the numpy slabs alive
        ncols = len(batch.columns)
        nchunks = batch.nchunks
        self.counts = np.ascontiguousarray(batch.counts, dtype=np.uint32)
        self._cols = (nat.HostColumn * max(ncols, 1))()
        self._keep = []
        for j, col in enumerate(batch.columns):
            self._cols[j] = self._host_column(col, register_heap)
        self.struct = nat.HostBatch(ncols, nat_flags(pinned), nchunks, self.counts.ctypes.data, self._cols)


def _ptr_table(slab, offsets, scale: int) -> np.ndarray:
    base = slab.ctypes.data
    return (np.asarray(offsets, dtype=np.uint64) * np.uint64(scale) + np.uint64(base)).astype(np.uint64)


def _host_column(self, col, register_heap: bool = True) -> "nat.HostColumn":
    """dmb_host_column of a chunks.Column (recursively for STRUCT fields and LIST / MAP children)"""
    keep = self._keep
    name = col.name.encode()
    data_ptrs = None
    if col.data is not None and np.asarray(col.data_off).shape[0]:
        data_ptrs = _ptr_table(col.data, col.data_off, 1)
    elif col.data is not None:
        data_ptrs = np.zeros(1, dtype=np.uint64)
    val_ptrs = None
    if col.validity is not None and np.any(np.asarray(col.val_off) >= 0):
        vo = np.asarray(col.val_off, dtype=np.int64)
        val_ptrs = np.where(vo >= 0, col.validity.ctypes.data + 8 * vo, 0).astype(np.uint64)
    heap_base, heap_len = None, 0
    if getattr(col, "heap", None) is not None and register_heap:
        heap_base, heap_len = col.heap.ctypes.data, int(col.heap.shape[0])
    elif getattr(col, "inline_only", False):
        heap_base, heap_len = 1, 0  # DMB_HEAP_INLINE_ONLY
    dict_ptr = None
    if getattr(col, "dictionary", None) is not None:  # ENUM: dmb_enum_dict
        d_offs, d_data = ch.enum_dict_arrays(col.dictionary)
        ed = nat.EnumDict(len(col.dictionary), 0, d_offs.ctypes.data, d_data.ctypes.data)
        keep.append((d_offs, d_data, ed))
        dict_ptr = C.pointer(ed)
    list_ptr = None
    child_col = getattr(col, "list_child_col", None)
    if child_col is not None:  # LIST / MAP with a described child (VARCHAR, STRUCT, LIST)
        sizes = np.ascontiguousarray(col.list_child_sizes, dtype=np.uint64)
        cc = (nat.HostColumn * 1)()
        cc[0] = _host_column(self, child_col, register_heap=False)  # child vectors own their heaps: never one registered region
        hl = nat.HostList(child_col.type_id, child_col.phys, child_col.dec_width, child_col.dec_scale, None, None, sizes.ctypes.data,
                          C.cast(cc, C.POINTER(nat.HostColumn)))
        keep.append((sizes, cc, hl))
        list_ptr = C.pointer(hl)
    elif getattr(col, "list_child_data", None) is not None:  # LIST of a fixed-width child, flat form
        cphys = ch.phys_of_type(col.list_child_type, col.list_child_dec_width)
        cw = ch.PHYS_WIDTH[cphys]
        c_ptrs = _ptr_table(col.list_child_data, col.list_child_base, cw)
        cv_ptrs = None
        if col.list_child_validity is not None and np.any(np.asarray(col.list_child_val_off) >= 0):
            vo = np.asarray(col.list_child_val_off, dtype=np.int64)
            cv_ptrs = np.where(vo >= 0, col.list_child_validity.ctypes.data + 8 * vo, 0).astype(np.uint64)
        sizes = np.ascontiguousarray(col.list_child_sizes, dtype=np.uint64)
        hl = nat.HostList(col.list_child_type, cphys, col.list_child_dec_width, col.list_child_dec_scale, c_ptrs.ctypes.data,
                          cv_ptrs.ctypes.data if cv_ptrs is not None else None, sizes.ctypes.data, None)
        keep.append((c_ptrs, cv_ptrs, sizes, hl))
        list_ptr = C.pointer(hl)
    struct_ptr = None
    fields = getattr(col, "struct_fields", None)
    if fields is not None:
        fc = (nat.HostColumn * len(fields))()
        for i, f in enumerate(fields):
            fc[i] = _host_column(self, f, register_heap)
        hs = nat.HostStruct(len(fields), 0, C.cast(fc, C.POINTER(nat.HostColumn)))
        keep.append((fc, hs))
        struct_ptr = C.pointer(hs)
    keep.append((data_ptrs, val_ptrs, name))
    return nat.HostColumn(name, col.type_id, col.phys, col.dec_width, col.dec_scale,
                          C.cast(data_ptrs.ctypes.data, C.POINTER(C.c_void_p)) if data_ptrs is not None else None,
                          C.cast(val_ptrs.ctypes.data, C.POINTER(C.c_void_p)) if val_ptrs is not None else None,
                          heap_base, heap_len, dict_ptr, list_ptr, struct_ptr)


HostBatch._host_column = _host_column


def nat_flags(pinned: bool) -> int:
    return 1 if pinned else 0  # DMB_BATCH_PINNED


def _read_int32_le(data: bytes, offset: int) -> int:
    if offset + 4 > len(data):  # :452-455
        return 0
    return int(np.frombuffer(data, dtype="<i4", count=1, offset=offset)[0])


class ArrowResult:
    """`ArrowResult` of the reference (opaque handle, src/duckdb_arrow_native.mbt:3)."""

    def __init__(self, ctx: GpuContext, handle: int, host_batch: Optional[HostBatch]):
        self.ctx = ctx
        self.lib = ctx.lib
        self.handle = handle
        self._host_batch = host_batch

    # ---- construction (stands where Connection::query_arrow is, :123-135)
    @classmethod
    def from_chunks(cls, ctx: GpuContext, batch, pinned: bool = False, register_heap: bool = True) -> "ArrowResult":
        hb = HostBatch(batch, pinned=pinned, register_heap=register_heap)
        handle = ctx.lib.duckdb_mb_gpu_result_from_chunks(ctx.handle, C.byref(hb.struct))
        if ctx.lib.duckdb_mb_is_null_arrow_result(handle):
            raise DuckDBError(nat.last_error() or "arrow query failed")
        return cls(ctx, handle, hb)

    # ---- reference surface
    def column_count(self) -> int:  # :138-140
        return self.lib.duckdb_mb_arrow_column_count(self.handle)

    def row_count(self) -> int:  # :143-145
        return self.lib.duckdb_mb_arrow_row_count(self.handle)

    def get_schema(self) -> ArrowSchemaInfo:  # :148-158, JSON parser :310-418
        raw = nat.moonbit_bytes(self.lib.duckdb_mb_arrow_schema(self.handle)).decode("utf-8", errors="replace")
        try:
            items = json.loads(raw)
            fields = [ArrowField(str(it["name"]), bool(it["nullable"]), str(it["type_id"])) for it in items]
        except Exception as e:  # the reference's hand-written parser reports a message too
            raise DuckDBError(f"schema parse failed: {e}")
        return ArrowSchemaInfo(fields)

    def raw_column(self, kind: str, col: int, nullable: bool = False) -> bytes:
        """The packed blob exactly as the C getter returns it (src/duckdb_native.c:2357-2797)."""
        fn = getattr(self.lib, f"duckdb_mb_arrow_get_column_{kind}{'_nullable' if nullable else ''}")
        return nat.moonbit_bytes(fn(self.handle, col))

    def _blob(self, kind: str, col: int, nullable: bool = False) -> "nat.MoonbitBlob":
        """the getter's Bytes viewed in place (a numpy uint8 array inside a `with` block): decoded without a second copy"""
        if getattr(self, "lib", None) is None:  # a subclass that serves blobs itself (tests of the decoder rules)
            return contextlib.nullcontext(self.raw_column(kind, col, nullable))
        fn = getattr(self.lib, f"duckdb_mb_arrow_get_column_{kind}{'_nullable' if nullable else ''}")
        return nat.MoonbitBlob(fn(self.handle, col))

    @staticmethod
    def _count_ok(data: bytes, header: int) -> int:
        if len(data) < header:
            return 0
        count = _read_int32_le(data, 0)
        return 0 if (count <= 0 or count > DECODER_ROW_CAP) else count

    def _decode_fixed(self, data, width: int, dtype, nullable: bool, copy: bool = True):
        """data: the blob (bytes, or the in-place numpy view of _blob).  copy=False only when the caller derives a new
        array from the values before the blob is freed."""
        count = self._count_ok(data, 4)
        if not count or len(data) < 4 + count * width + (count if nullable else 0):
            return (np.zeros(0, dtype=dtype), np.zeros(0, dtype=bool)) if nullable else np.zeros(0, dtype=dtype)
        values = np.frombuffer(data, dtype=dtype, count=count, offset=4)
        if copy:
            values = values.copy()
        if not nullable:
            return values
        valid = np.frombuffer(data, dtype=np.uint8, count=count, offset=4 + count * width) != 0
        return values, valid

    def get_column_int32(self, col: int) -> np.ndarray:  # :421-449
        with self._blob("int32", col) as b:
            return self._decode_fixed(b, 4, np.dtype("<i4"), False)

    def get_column_int32_nullable(self, col: int):  # :619-652
        with self._blob("int32", col, True) as b:
            return self._decode_fixed(b, 4, np.dtype("<i4"), True)

    def get_column_int64(self, col: int) -> np.ndarray:
        """MoonBit `Int` is 32-bit on the native target: the byte-assembled value keeps the low 32
        bits of each int64 (:465-505)."""
        with self._blob("int64", col) as b:
            v = self._decode_fixed(b, 8, np.dtype("<i8"), False, copy=False)
            return v.astype(np.int32)  # wraps: low 32 bits

    def get_column_int64_nullable(self, col: int):  # :655-704
        with self._blob("int64", col, True) as b:
            v, valid = self._decode_fixed(b, 8, np.dtype("<i8"), True, copy=False)
            return v.astype(np.int32), valid

    def get_column_int64_exact(self, col: int, nullable: bool = False):
        """Additive: the full 64-bit values of the same blob."""
        return self._decode_fixed(self.raw_column("int64", col, nullable), 8, np.dtype("<i8"), nullable)

    def get_column_double(self, col: int) -> np.ndarray:  # :508-543
        with self._blob("double", col) as b:
            return self._decode_fixed(b, 8, np.dtype("<f8"), False)

    def get_column_double_nullable(self, col: int):  # :707-741
        with self._blob("double", col, True) as b:
            return self._decode_fixed(b, 8, np.dtype("<f8"), True)

    def get_column_bool(self, col: int) -> np.ndarray:  # :579-603
        with self._blob("bool", col) as b:
            return self._decode_fixed(b, 1, np.uint8, False, copy=False) != 0

    def get_column_bool_nullable(self, col: int):  # :790-822
        with self._blob("bool", col, True) as b:
            v, valid = self._decode_fixed(b, 1, np.uint8, True, copy=False)
            return v != 0, valid

    def _decode_strings(self, data: bytes, nullable: bool):
        """NUL scan from byte 8 (:559-574, :752-784); lossy UTF-8 decode like decode_lossy."""
        empty = ([], np.zeros(0, dtype=bool)) if nullable else []
        count = self._count_ok(data, 8)
        if not count:
            return empty
        total = _read_int32_le(data, 4)
        if nullable and len(data) < 8 + total + count:
            return empty
        buf = np.frombuffer(data, dtype=np.uint8)
        nul = np.flatnonzero(buf[8:] == 0) + 8
        out: List[str] = []
        pos = 8
        k = 0
        n_nul = nul.shape[0]
        for _ in range(count):
            start = pos
            while k < n_nul and nul[k] < pos:
                k += 1
            end = int(nul[k]) if k < n_nul else len(data)
            if start < len(data):
                out.append(data[start:end].decode("utf-8", errors="replace"))
            else:
                out.append("")
            pos = end + 1
        if not nullable:
            return out
        valid = np.frombuffer(data, dtype=np.uint8, count=count, offset=8 + total) != 0
        return out, valid

    def get_column_string_spans_nullable(self, col: int):
        """The nullable string decoder's scan (:752-784) without building a String per row: (starts, ends, valid, blob)
        with row i = blob[starts[i]:ends[i]].  Row i ends at the i-th NUL at or after byte 8 (or at the end of the blob)."""
        with self._blob("string", col, True) as view:
            return decode_string_spans(view, True)

    def get_column_string(self, col: int) -> List[str]:  # :546-575
        return self._decode_strings(self.raw_column("string", col), False)

    def get_column_string_nullable(self, col: int):  # :744-787
        return self._decode_strings(self.raw_column("string", col, True), True)

    def close(self) -> None:  # :606-611
        if self.handle:
            self.lib.duckdb_mb_arrow_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- additive: real Arrow (L1)
    def materialise(self) -> None:
        if not self.lib.duckdb_mb_gpu_result_materialise_arrow(self.handle):
            raise DuckDBError(nat.last_error())

    def timings(self) -> dict:
        t = (C.c_double * 4)()
        b = (C.c_uint64 * 2)()
        self.lib.duckdb_mb_gpu_result_timings(self.handle, t)
        self.lib.duckdb_mb_gpu_result_link_bytes(self.handle, b)
        return {"h2d_ms": t[0], "kernels_ms": t[1], "d2h_ms": t[2], "total_ms": t[3], "h2d_bytes": b[0], "d2h_bytes": b[1]}

    def export_c(self, col: int = -1):
        arr, sch = nat.ArrowArray(), nat.ArrowSchema()
        if not self.lib.duckdb_mb_gpu_result_export_arrow(self.handle, col, C.addressof(arr), C.addressof(sch)):
            raise DuckDBError(nat.last_error())
        return arr, sch

    def to_arrow(self, col: Optional[int] = None):
        """pyarrow arrays over the result's page-locked buffers (zero copy).  col=None: list of all
        columns; the arrays stay valid after close()."""
        import pyarrow as pa

        if col is not None:
            arr, sch = self.export_c(col)
            return pa.Array._import_from_c(C.addressof(arr), C.addressof(sch))
        arr, sch = self.export_c(-1)
        struct = pa.Array._import_from_c(C.addressof(arr), C.addressof(sch))
        return [struct.field(i) for i in range(struct.type.num_fields)]

    def to_stream(self, max_batch_rows: int = 0):
        """pyarrow.RecordBatchReader over the Arrow C stream interface: record batches of at most `max_batch_rows` rows
        (0: 16 M), converted one batch ahead of the consumer; a column with > 2^31 string bytes arrives as several utf8
        batches.  The chunk vectors must stay alive until the reader is closed."""
        import pyarrow as pa

        st = nat.ArrowArrayStream()
        if not self.lib.duckdb_mb_gpu_result_export_stream(self.handle, int(max_batch_rows), C.addressof(st)):
            raise DuckDBError(nat.last_error())
        return _keepalive_reader(pa.RecordBatchReader._import_from_c(C.addressof(st)), (self._host_batch,))

    def to_record_batch(self):
        import pyarrow as pa

        arr, sch = self.export_c(-1)
        return pa.RecordBatch._import_from_c(C.addressof(arr), C.addressof(sch))


def _keepalive_reader(inner, keep):
    """a RecordBatchReader over `inner` whose generator keeps the host chunk vectors alive until it is exhausted / closed"""
    import pyarrow as pa

    def batches(_keep=keep):
        for b in inner:
            yield b

    return pa.RecordBatchReader.from_batches(inner.schema, batches())


def _nul_positions(stream: np.ndarray, limit: int) -> np.ndarray:
    """positions of the first `limit` NUL bytes of a uint8 array"""
    return np.flatnonzero(stream == 0)[:limit].astype(np.int64)


def decode_string_spans(view, nullable: bool):
    """decoder mirror of the string blob [n:i32][total:i32][s0\\0 s1\\0 ...][valid bytes]: one copy of the stream + the spans
    into it.  Row i ends at the i-th NUL at or after byte 8; only the `total` stream bytes are scanned when they hold a NUL
    per row (every non-NULL layout), else the scan runs on into the validity bytes exactly like the reference's loop."""
    data = np.array(view, dtype=np.uint8, copy=True)
    z = np.zeros(0, dtype=np.int64)
    count = ArrowResult._count_ok(data, 8)
    if not count:
        return z, z, np.zeros(0, dtype=bool), data
    total = _read_int32_le(data, 4)
    if nullable and len(data) < 8 + total + count:
        return z, z, np.zeros(0, dtype=bool), data
    nul = _nul_positions(data[8: 8 + total], count)
    if nul.shape[0] < count:  # NULL rows' terminators are not counted in `total`: the reference's scan runs on
        nul = _nul_positions(data[8:], count)
    nul = nul + 8
    ends = np.full(count, len(data), dtype=np.int64)
    ends[: nul.shape[0]] = nul
    starts = np.empty(count, dtype=np.int64)
    starts[0] = 8
    starts[1:] = np.minimum(ends[:-1] + 1, len(data))
    valid = (data[8 + total: 8 + total + count] != 0) if nullable else np.ones(count, dtype=bool)
    return starts, ends, valid, data


class ShardedResult:
    """One table over N GPU contexts: contiguous chunk ranges, one independent record batch per part, no
    data-path collective; `string_bases` is the host exclusive scan that rebases utf8 offsets across the parts."""

    def __init__(self, ctxs: Sequence[GpuContext], batch, pinned: bool = False, register_heap: bool = True):
        self.lib = ctxs[0].lib
        self.ctxs = list(ctxs)
        self._hb = HostBatch(batch, pinned=pinned, register_heap=register_heap)
        arr = (C.c_void_p * len(ctxs))(*[c.handle for c in ctxs])
        self.handle = self.lib.duckdb_mb_gpu_result_from_chunks_sharded(arr, len(ctxs), C.byref(self._hb.struct))
        if not self.handle:
            raise DuckDBError(nat.last_error() or "sharded result failed")

    @property
    def part_count(self) -> int:
        return self.lib.duckdb_mb_gpu_sharded_part_count(self.handle)

    def first_row(self, i: int) -> int:
        return self.lib.duckdb_mb_gpu_sharded_first_row(self.handle, i)

    def materialise(self) -> None:
        if not self.lib.duckdb_mb_gpu_sharded_materialise_arrow(self.handle):
            raise DuckDBError(nat.last_error())

    def part(self, i: int) -> ArrowResult:
        """the i-th part as an ArrowResult (borrowed: do not close it)"""
        h = self.lib.duckdb_mb_gpu_sharded_part(self.handle, i)
        if not h:
            raise IndexError(i)
        r = ArrowResult(self.ctxs[0], h, self._hb)
        r.close = lambda: None  # owned by the sharded handle
        return r

    def string_bases(self, col: int) -> List[int]:
        out = (C.c_uint64 * (self.part_count + 1))()
        if not self.lib.duckdb_mb_gpu_sharded_string_bases(self.handle, col, out):
            raise DuckDBError(nat.last_error())
        return [int(v) for v in out]

    def to_stream(self):
        """hands the parts over to a pyarrow.RecordBatchReader (one record batch per part)"""
        import pyarrow as pa

        st = nat.ArrowArrayStream()
        if not self.lib.duckdb_mb_gpu_sharded_export_stream(self.handle, C.addressof(st)):
            raise DuckDBError(nat.last_error())
        self.handle = None  # the stream owns it now
        return _keepalive_reader(pa.RecordBatchReader._import_from_c(C.addressof(st)), (self._hb,))

    def close(self) -> None:
        if self.handle:
            self.lib.duckdb_mb_gpu_sharded_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
