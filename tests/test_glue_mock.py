"""The glue as code (round-1 verdict, items 2, 3 and 5): glue/duckdb_gpu_glue.c compiled against a mock DuckDB C API
(glue/mock/) that serves canned DataChunks, and driven SQL-less:  mock libduckdb -> glue (duckdb_mb_query_arrow /
duckdb_mb_query / duckdb_mb_query_stream / duckdb_mb_appender_create) -> the library's drop-in symbols -> oracle.

CPU part: the glue builds and links, and glue + library together export every symbol SURVEY.md §8b lists.
GPU part: the reference's arrow tests replayed through the glue; a mixed batch (fixed types, DECIMAL, VARCHAR with
per-vector heaps, ENUM, LIST) bit-exact against the oracle through getters / per-cell / streaming symbols; the appender
path (reference row protocol with MoonBit Bytes arguments) into duckdb_append_data_chunk of the mock.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "glue"))

import oracle  # noqa: E402
from duckdb_mbt_b200 import chunks as ch  # noqa: E402
from duckdb_mbt_b200 import native as nat  # noqa: E402

import build as glue_build  # noqa: E402


class MockColumn(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type_id", C.c_int32), ("width", C.c_int32), ("dec_width", C.c_int32), ("dec_scale", C.c_int32),
                ("data", C.c_void_p), ("validity", C.c_void_p), ("dict_size", C.c_uint32), ("dict_values", C.c_void_p),
                ("child_type_id", C.c_int32), ("child_width", C.c_int32), ("child_data", C.c_void_p), ("child_validity", C.c_void_p),
                ("child_sizes", C.c_void_p)]


class Conn(C.Structure):  # duckdb_mb_connection {duckdb_database db; duckdb_connection conn;}
    _fields_ = [("db", C.c_void_p), ("conn", C.c_void_p)]


class Glue:
    def __init__(self):
        import __graft_entry__ as ge
        if not os.path.exists(ge.LIB):
            ge.build()
        self.lib = nat.lib()
        mock_path, glue_path = glue_build.build()
        self.mock = C.CDLL(mock_path, mode=C.RTLD_GLOBAL)
        self.glue = C.CDLL(glue_path)
        vp = C.c_void_p
        for name in ("duckdb_mb_query_arrow", "duckdb_mb_query", "duckdb_mb_query_stream"):
            f = getattr(self.glue, name)
            f.restype, f.argtypes = vp, [C.POINTER(Conn), vp]
        self.glue.duckdb_mb_appender_create.restype = vp
        self.glue.duckdb_mb_appender_create.argtypes = [C.POINTER(Conn), vp, vp]
        self.glue.duckdb_mb_last_error.restype = vp
        self.mock.duckdb_mock_register_table.argtypes = [C.c_char_p, C.c_int32, C.POINTER(MockColumn), C.c_int64, vp]
        self.mock.duckdb_mock_register_append_table.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, vp, vp, vp]
        for name in ("duckdb_mock_appended_rows", "duckdb_mock_appended_chunks", "duckdb_mock_append_flushes"):
            getattr(self.mock, name).restype = C.c_int64
            getattr(self.mock, name).argtypes = [C.c_char_p]
        self.mock.duckdb_mock_appended_chunk.restype = C.c_int64
        self.mock.duckdb_mock_appended_chunk.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.POINTER(vp), C.POINTER(vp)]
        self.conn = Conn(1, 1)  # the mock ignores the handles
        self.keep = []
        self.nsql = 0

    def bytes_(self, b: bytes):
        """a MoonBit Bytes (stand-in header of include/moonbit_standin.h): [rc:i32][len:u32] payload"""
        buf = C.create_string_buffer(8 + len(b) + 1)
        C.memmove(buf, np.asarray([1], dtype=np.int32).tobytes() + np.asarray([len(b)], dtype=np.uint32).tobytes(), 8)
        C.memmove(C.addressof(buf) + 8, b, len(b))
        self.keep.append(buf)
        return C.addressof(buf) + 8

    def last_error(self) -> str:
        return nat.moonbit_bytes(self.glue.duckdb_mb_last_error()).decode()

    def register(self, batch) -> bytes:
        """serve `batch` as the result of a fresh SQL text"""
        self.nsql += 1
        sql = f"SELECT * FROM canned_{self.nsql}".encode()
        cols = (MockColumn * max(len(batch.columns), 1))()
        nch = batch.nchunks
        for j, col in enumerate(batch.columns):
            base = col.data.ctypes.data
            data = (np.asarray(col.data_off, dtype=np.uint64) + np.uint64(base)).astype(np.uint64)
            val = None
            if col.validity is not None and np.any(np.asarray(col.val_off) >= 0):
                vo = np.asarray(col.val_off, dtype=np.int64)
                val = np.where(vo >= 0, col.validity.ctypes.data + 8 * vo, 0).astype(np.uint64)
            mc = MockColumn(col.name.encode(), col.type_id, col.width, col.dec_width, col.dec_scale, data.ctypes.data if nch else None,
                            val.ctypes.data if val is not None else None, 0, None, 0, 0, None, None, None)
            self.keep += [data, val]
            if getattr(col, "dictionary", None) is not None:
                labels = [C.create_string_buffer(x) for x in col.dictionary]
                arr = (C.c_char_p * len(labels))(*[C.cast(x, C.c_char_p) for x in labels])
                mc.dict_size, mc.dict_values = len(labels), C.cast(arr, C.c_void_p)
                self.keep += [labels, arr]
            if getattr(col, "list_child_data", None) is not None:
                cphys = ch.phys_of_type(col.list_child_type, col.list_child_dec_width)
                cw = ch.PHYS_WIDTH[cphys]
                cptr = (np.asarray(col.list_child_base, dtype=np.uint64) * np.uint64(cw) + np.uint64(col.list_child_data.ctypes.data)).astype(np.uint64)
                cval = None
                if col.list_child_validity is not None and np.any(np.asarray(col.list_child_val_off) >= 0):
                    vo = np.asarray(col.list_child_val_off, dtype=np.int64)
                    cval = np.where(vo >= 0, col.list_child_validity.ctypes.data + 8 * vo, 0).astype(np.uint64)
                sizes = np.ascontiguousarray(col.list_child_sizes, dtype=np.uint64)
                mc.child_type_id, mc.child_width = col.list_child_type, cw
                mc.child_data, mc.child_validity, mc.child_sizes = cptr.ctypes.data, (cval.ctypes.data if cval is not None else None), sizes.ctypes.data
                self.keep += [cptr, cval, sizes]
            cols[j] = mc
        counts = np.ascontiguousarray(batch.counts, dtype=np.uint32)
        self.keep += [cols, counts, batch]
        assert self.mock.duckdb_mock_register_table(sql, len(batch.columns), cols, nch, counts.ctypes.data if nch else None)
        return sql

    def query_arrow(self, batch):
        from duckdb_mbt_b200 import arrow_result as ar
        h = self.glue.duckdb_mb_query_arrow(C.byref(self.conn), self.bytes_(self.register(batch)))
        assert h, self.last_error()
        return ar.ArrowResult(self, h, None)  # (`self` stands for the context object: ArrowResult only needs `.lib`)


@pytest.fixture(scope="module")
def glue():
    return Glue()


# ------------------------------------------------------------------------------------------- CPU
SURVEY_8B_SYMBOLS = (
    ["duckdb_mb_query_arrow", "duckdb_mb_arrow_column_count", "duckdb_mb_arrow_row_count", "duckdb_mb_arrow_schema", "duckdb_mb_arrow_destroy",
     "duckdb_mb_is_null_arrow_result", "duckdb_mb_bytes_to_double"]
    + [f"duckdb_mb_arrow_get_column_{k}{s}" for k in ("int32", "int64", "double", "string", "bool") for s in ("", "_nullable")]
    + ["duckdb_mb_query"] + [f"duckdb_mb_result_{x}" for x in ("destroy", "column_count", "row_count", "column_name", "column_type", "is_null", "value")]
    + ["duckdb_mb_query_stream"] + [f"duckdb_mb_stream_{x}" for x in ("destroy", "column_count", "column_name", "fetch_chunk")]
    + [f"duckdb_mb_chunk_{x}" for x in ("destroy", "row_count", "column_count", "is_null", "value")]
    + [f"duckdb_mb_appender_{x}" for x in ("create", "destroy", "error")]
    + [f"duckdb_mb_{x}" for x in ("begin_row", "append_int", "append_bigint", "append_double", "append_varchar", "append_bool", "append_null",
                                  "end_row", "flush")]
    + [f"duckdb_mb_append_{x}" for x in ("date", "timestamp", "blob", "decimal", "interval", "list_varchar", "struct_varchar", "map_varchar_varchar")]
)


def test_glue_builds_and_every_survey_8b_symbol_is_exported(glue):
    lib = C.CDLL(nat.LIB_PATH)
    missing = [s for s in SURVEY_8B_SYMBOLS if not hasattr(lib, s) and not hasattr(glue.glue, s)]
    assert not missing, missing
    # the four entry points that need libduckdb are the glue's, everything else the library's own
    for s in ("duckdb_mb_query_arrow", "duckdb_mb_query", "duckdb_mb_query_stream", "duckdb_mb_appender_create"):
        assert hasattr(glue.glue, s) and not hasattr(lib, s)


def test_glue_query_failure_keeps_the_reference_convention(glue):
    """unknown SQL -> NULL handle + duckdb_mb_last_error() (src/duckdb_native.c:160-170)"""
    h = glue.glue.duckdb_mb_query_arrow(C.byref(glue.conn), glue.bytes_(b"SELEC nonsense"))
    assert not h
    assert glue.last_error() != ""
    assert not glue.glue.duckdb_mb_query_arrow(None, glue.bytes_(b"x"))
    assert glue.last_error() == "invalid connection handle"
    assert not glue.glue.duckdb_mb_query(None, glue.bytes_(b"x")) and glue.last_error() == "connection is null"


# ------------------------------------------------------------------------------------------- GPU
torch = pytest.importorskip("torch")
needs_gpu = pytest.mark.gpu


def _mixed_with_everything(n, seed):
    from test_gpu_l0_parity import _mixed_batch
    import list_cases
    b = _mixed_batch(n, "ragged", seed)
    rng = np.random.default_rng(seed + 1)
    b.columns.append(ch.string_column_bulk("s", rng.integers(0, 50, n), rng.random(n) > 0.15, b.counts, rng, utf8_fraction=0.1))
    labels = [b"AIR", b"RAIL", b"a label longer than twelve bytes", b"", b"TRUCK"]
    b.columns.append(ch.enum_column("e", labels, rng.integers(0, len(labels), n), b.counts, valid=rng.random(n) > 0.2, garbage_rng=None))
    return b


@needs_gpu
def test_reference_arrow_tests_through_the_glue(glue):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from test_gpu_host_api import reference_arrow_cases
    reference_arrow_cases(glue.query_arrow)


@needs_gpu
def test_mock_to_glue_to_getters_equals_the_oracle(glue):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    batch = _mixed_with_everything(20_011, 41)
    ora = oracle.OracleResult(batch)
    with glue.query_arrow(batch) as res:
        assert res.column_count() == len(batch.columns) and res.row_count() == batch.nrows
        assert nat.moonbit_bytes(res.lib.duckdb_mb_arrow_schema(res.handle)) == ora.schema()
        for col, c in enumerate(batch.columns):
            if c.phys == ch.P_STRING or c.type_id == ch.T_ENUM:
                for nullable in (False, True):
                    assert res.raw_column("string", col, nullable) == ora.get_column("string", col, nullable), (c.name, nullable)
                continue
            if c.phys in (ch.P_U128, ch.P_INTERVAL):
                continue
            for kind in ("int32", "int64", "double", "bool"):
                assert res.raw_column(kind, col, True) == ora.get_column(kind, col, True), (c.name, kind)
        arrays = res.to_arrow()
        s_col = len(batch.columns) - 2
        eo, ed = ora.arrow_string(s_col, 0)
        assert np.array_equal(np.frombuffer(arrays[s_col].buffers()[1], dtype=np.int32)[: batch.nrows + 1], eo)
        assert bytes(arrays[s_col].buffers()[2])[: ed.shape[0]] == ed.tobytes()


@needs_gpu
def test_query_and_query_stream_cells_through_the_glue(glue):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from test_oracle_golden import batch_of
    L = nat.lib()
    batch = batch_of(("i", ch.T_INTEGER, [1, None, 3]), ("s", ch.T_VARCHAR, ["alpha", "a string longer than twelve", None]),
                     ("d", ch.T_DOUBLE, [1.5, 2.25, None]), ("dt", ch.T_DATE, [19877, None, -1]))
    ora = oracle.OracleResult(batch)
    vp = C.c_void_p
    L.duckdb_mb_result_value.restype, L.duckdb_mb_result_value.argtypes = vp, [vp, C.c_int32, C.c_int32]
    L.duckdb_mb_result_is_null.argtypes = [vp, C.c_int32, C.c_int32]
    h = glue.glue.duckdb_mb_query(C.byref(glue.conn), glue.bytes_(glue.register(batch)))
    assert h, glue.last_error()
    assert L.duckdb_mb_result_column_count(vp(h)) == 4 and L.duckdb_mb_result_row_count(vp(h)) == 3
    for c in range(4):
        for r in range(3):
            assert bool(L.duckdb_mb_result_is_null(vp(h), c, r)) == ora.cell_is_null(c, r)
            assert nat.moonbit_bytes(L.duckdb_mb_result_value(vp(h), c, r)) == ora.cell_value(c, r)
    L.duckdb_mb_result_destroy(vp(h))
    # streaming: the stream owns the result (src/duckdb_native.c:426-438)
    s = glue.glue.duckdb_mb_query_stream(C.byref(glue.conn), glue.bytes_(glue.register(batch)))
    assert s, glue.last_error()
    L.duckdb_mb_stream_fetch_chunk.restype, L.duckdb_mb_stream_fetch_chunk.argtypes = vp, [vp]
    L.duckdb_mb_chunk_value.restype, L.duckdb_mb_chunk_value.argtypes = vp, [vp, C.c_int32, C.c_int32]
    L.duckdb_mb_chunk_is_null.argtypes = [vp, C.c_int32, C.c_int32]
    L.duckdb_mb_chunk_row_count.argtypes = [vp]
    ck = L.duckdb_mb_stream_fetch_chunk(vp(s))
    assert ck and L.duckdb_mb_chunk_row_count(vp(ck)) == 3
    for c in range(4):
        for r in range(3):
            assert bool(L.duckdb_mb_chunk_is_null(vp(ck), c, r)) == ora.cell_is_null(c, r)
            assert nat.moonbit_bytes(L.duckdb_mb_chunk_value(vp(ck), c, r)) == ora.cell_value(c, r)
    L.duckdb_mb_chunk_destroy(vp(ck))
    assert not L.duckdb_mb_stream_fetch_chunk(vp(s))  # end of stream
    L.duckdb_mb_stream_destroy(vp(s))
    # a LIST column is rejected by the stream whitelist with the reference's message (:335-338)
    import list_cases
    lc = list_cases.make_list_column(100, 4, "full", 3, "contiguous")
    lb = ch.ChunkBatch(lc.counts, [list_cases.as_column(lc, "l", ch.T_INTEGER)])
    assert not glue.glue.duckdb_mb_query_stream(C.byref(glue.conn), glue.bytes_(glue.register(lb)))
    assert glue.last_error() == "streaming query has unsupported column type"


@needs_gpu
def test_reference_appender_symbols_into_duckdb_append_data_chunk(glue):
    """duckdb_mb_appender_create(conn, schema, table) -> duckdb_mb_begin_row / append_* (MoonBit Bytes arguments) /
    end_row / flush -> the sink fills duckdb_data_chunks -> duckdb_append_data_chunk (the bulk door the reference leaves
    unbound, src/duckdb_native.c:2109-2132)"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    L = nat.lib()
    vp = C.c_void_p
    types = np.asarray([ch.T_INTEGER, ch.T_BIGINT, ch.T_DOUBLE, ch.T_VARCHAR, ch.T_BOOLEAN, ch.T_DATE, ch.T_DECIMAL], dtype=np.int32)
    dw = np.asarray([0, 0, 0, 0, 0, 0, 18], dtype=np.int32)
    ds = np.asarray([0, 0, 0, 0, 0, 0, 2], dtype=np.int32)
    assert glue.mock.duckdb_mock_register_append_table(b"main", b"t_app", len(types), types.ctypes.data, dw.ctypes.data, ds.ctypes.data)
    assert not glue.glue.duckdb_mb_appender_create(C.byref(glue.conn), glue.bytes_(b"main"), glue.bytes_(b"no_such_table"))
    a = glue.glue.duckdb_mb_appender_create(C.byref(glue.conn), glue.bytes_(b"main"), glue.bytes_(b"t_app"))
    assert a
    a = vp(a)
    for f, at in (("duckdb_mb_append_int", [vp, C.c_int32]), ("duckdb_mb_append_bigint", [vp, C.c_int64]), ("duckdb_mb_append_double", [vp, C.c_double]),
                  ("duckdb_mb_append_varchar", [vp, vp]), ("duckdb_mb_append_bool", [vp, C.c_bool]), ("duckdb_mb_append_date", [vp, C.c_int32]),
                  ("duckdb_mb_append_decimal", [vp, C.c_uint8, C.c_uint8, C.c_int64, C.c_int64]), ("duckdb_mb_begin_row", [vp]), ("duckdb_mb_end_row", [vp]),
                  ("duckdb_mb_append_null", [vp]), ("duckdb_mb_flush", [vp]), ("duckdb_mb_appender_destroy", [vp])):
        getattr(L, f).argtypes = at
    L.duckdb_mb_appender_error.restype, L.duckdb_mb_appender_error.argtypes = vp, [vp]
    n = 5000
    rng = np.random.default_rng(8)
    strs = [None if i % 11 == 0 else bytes(rng.integers(0x61, 0x7B, int(rng.integers(0, 40)), dtype=np.uint8)) for i in range(n)]
    for i in range(n):
        assert L.duckdb_mb_begin_row(a)
        assert L.duckdb_mb_append_int(a, i)
        assert L.duckdb_mb_append_bigint(a, i * 10**10) if i % 7 else L.duckdb_mb_append_null(a)
        assert L.duckdb_mb_append_double(a, i / 4.0)
        assert L.duckdb_mb_append_varchar(a, glue.bytes_(strs[i])) if strs[i] is not None else L.duckdb_mb_append_null(a)
        assert L.duckdb_mb_append_bool(a, i % 3 == 0)
        assert L.duckdb_mb_append_date(a, 19877 + i)
        assert L.duckdb_mb_append_decimal(a, 18, 2, 100 * i + 5, 0)
        assert L.duckdb_mb_end_row(a), nat.moonbit_bytes(L.duckdb_mb_appender_error(a))
    assert L.duckdb_mb_flush(a), nat.moonbit_bytes(L.duckdb_mb_appender_error(a))
    assert glue.mock.duckdb_mock_appended_rows(b"t_app") == n and glue.mock.duckdb_mock_append_flushes(b"t_app") == 1
    assert glue.mock.duckdb_mock_appended_chunks(b"t_app") == (n + 2047) // 2048
    # read the table back chunk by chunk
    got = {c: [] for c in range(7)}
    valid = {c: [] for c in range(7)}
    widths = [4, 8, 8, 16, 1, 4, 8]
    for k in range((n + 2047) // 2048):
        for c in range(7):
            dptr, mptr = vp(), vp()
            cnt = glue.mock.duckdb_mock_appended_chunk(b"t_app", k, c, C.byref(dptr), C.byref(mptr))
            raw = np.frombuffer(C.string_at(dptr.value, cnt * widths[c]), dtype=np.uint8)
            mask = np.ones(cnt, dtype=bool) if not mptr.value else np.unpackbits(np.frombuffer(C.string_at(mptr.value, 256), dtype=np.uint8), bitorder="little")[:cnt].astype(bool)
            valid[c].append(mask)
            if c == 3:
                ent = raw.reshape(-1, 16)
                out = []
                for i in range(cnt):
                    if not mask[i]:
                        out.append(None)
                        continue
                    ln = int(ent[i, 0:4].view(np.uint32)[0])
                    out.append(bytes(ent[i, 4:4 + ln]) if ln <= 12 else C.string_at(int(ent[i, 8:16].view(np.uint64)[0]), ln))
                got[c] += out
            else:
                got[c].append(raw)
    idx = np.arange(n)
    assert np.array_equal(np.concatenate(got[0]).view(np.int32), idx.astype(np.int32))
    v1 = np.concatenate(valid[1])
    assert np.array_equal(v1, idx % 7 != 0)
    assert np.array_equal(np.concatenate(got[1]).view(np.int64)[v1], (idx * 10**10)[v1])
    assert np.array_equal(np.concatenate(got[2]).view(np.float64), idx / 4.0)
    assert got[3] == strs
    assert np.array_equal(np.concatenate(got[4]), (idx % 3 == 0).astype(np.uint8))
    assert np.array_equal(np.concatenate(got[5]).view(np.int32), (19877 + idx).astype(np.int32))
    assert np.array_equal(np.concatenate(got[6]).view(np.int64), 100 * idx + 5)
    # protocol errors keep the reference's 1/0 + per-handle message convention
    assert L.duckdb_mb_end_row(a) == 0
    assert nat.moonbit_bytes(L.duckdb_mb_appender_error(a)) != b""
    L.duckdb_mb_appender_destroy(a)
