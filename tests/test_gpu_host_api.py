"""GPU parity through the C ABI with HOST buffers: the drop-in symbols `duckdb_mb_arrow_*` (L2) and
the Arrow / typed exports (L1) against the CPU oracle, bit-exact, on the same seeded DuckDB-shaped
chunk batches.  Also the reference's own arrow-test vectors (src/duckdb_arrow_test.mbt) replayed
through the GPU path."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
pa = pytest.importorskip("pyarrow")

import oracle  # noqa: E402
from duckdb_mbt_b200 import chunks as ch  # noqa: E402

from test_gpu_l0_parity import _mixed_batch  # noqa: E402
from test_oracle_golden import batch_of  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from duckdb_mbt_b200 import arrow_result as ar
    c = ar.GpuContext(0)
    yield c
    c.close()


def _result(ctx, batch, **kw):
    from duckdb_mbt_b200 import arrow_result as ar
    return ar.ArrowResult.from_chunks(ctx, batch, **kw)


KINDS = ("int32", "int64", "double", "bool")


def _with_strings(n, pattern, seed):
    b = _mixed_batch(n, pattern, seed)
    rng = np.random.default_rng(seed + 1)
    lens = rng.integers(0, 50, n)
    b.columns.append(ch.string_column_bulk("s", lens, rng.random(n) > 0.15, b.counts, rng, utf8_fraction=0.1))
    b.columns.append(ch.string_column_bulk("s_nonull", rng.integers(0, 20, n), None, b.counts, rng))
    return b


@pytest.mark.parametrize("n,pattern", [(1, "full"), (2049, "full"), (10_000, "ragged"), (70_001, "ragged")])
def test_reference_getters_bit_exact(ctx, n, pattern):
    batch = _with_strings(n, pattern, 300 + n)
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        assert res.column_count() == len(batch.columns)
        assert res.row_count() == n
        for col, c in enumerate(batch.columns):
            if c.phys == ch.P_STRING:
                for nullable in (False, True):
                    assert res.raw_column("string", col, nullable) == ora.get_column("string", col, nullable), (col, nullable)
                continue
            if c.phys in (ch.P_U128, ch.P_INTERVAL):
                continue  # libduckdb casts of UUID/INTERVAL to numbers are errors; not on the path
            for kind in KINDS:
                for nullable in (False, True):
                    got = res.raw_column(kind, col, nullable)
                    exp = ora.get_column(kind, col, nullable)
                    assert got == exp, f"col={c.name} kind={kind} nullable={nullable}"


def test_getters_bad_arguments_return_empty_bytes(ctx):
    batch = ch.config_c1(100)
    with _result(ctx, batch) as res:
        for kind in KINDS + ("string",):
            assert res.raw_column(kind, -1) == b""
            assert res.raw_column(kind, 3) == b""
            assert res.raw_column(kind, 99, True) == b""
    empty = ch.ChunkBatch(np.zeros(0, dtype=np.uint32), [ch.fixed_column("x", ch.T_INTEGER, np.zeros(0, np.int32), np.zeros(0, np.uint32))])
    with _result(ctx, empty) as res:
        assert res.row_count() == 0 and res.column_count() == 1
        assert res.raw_column("int32", 0) == b""  # row_count <= 0 -> empty Bytes (src/duckdb_native.c:2366-2368)
        assert res.get_column_int32(0).shape == (0,)
        arrays = res.to_arrow()
        assert len(arrays) == 1 and len(arrays[0]) == 0


def test_unsupported_cast_is_zero_filled_with_validity(ctx):
    rng = np.random.default_rng(3)
    n = 5000
    counts = ch.chunk_counts(n)
    valid = rng.random(n) > 0.3
    batch = ch.ChunkBatch(counts, [ch.string_column_bulk("s", rng.integers(0, 9, n), valid, counts, rng)])
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        for kind in ("int32", "int64", "double"):
            assert res.raw_column(kind, 0, True) == ora.get_column(kind, 0, True)


def test_schema_json_matches_reference_format(ctx):
    batch = _with_strings(10, "full", 1)
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        from duckdb_mbt_b200 import native as nat
        raw = nat.moonbit_bytes(res.lib.duckdb_mb_arrow_schema(res.handle))
        assert raw == ora.schema()
        fields = res.get_schema().fields
        assert [f.type_id for f in fields][:5] == ["bool", "int32", "int32", "int32", "int64"]
        assert all(f.nullable for f in fields)


# ------------------------------------------------------------------ src/duckdb_arrow_test.mbt replayed
def reference_arrow_cases(make_result):
    """src/duckdb_arrow_test.mbt:128-518 replayed; make_result(batch) -> ArrowResult (host API directly, or SQL-less
    through the glue: tests/test_glue_mock.py)"""
    # :210-228 RANGE(5) via the int32 getter on a BIGINT column
    with make_result(batch_of(("range", ch.T_BIGINT, [0, 1, 2, 3, 4]))) as r:
        assert r.get_column_int32(0).tolist() == [0, 1, 2, 3, 4]
    # :231-247
    with make_result(batch_of(("x", ch.T_BIGINT, [100]))) as r:
        assert r.get_column_int64(0).tolist() == [100]
    # :250-269
    with make_result(batch_of(("x", ch.T_DOUBLE, [3.14]))) as r:
        assert abs(r.get_column_double(0)[0] - 3.14) < 0.01
    # :272-296
    with make_result(batch_of(("t", ch.T_BOOLEAN, [1]), ("f", ch.T_BOOLEAN, [0]))) as r:
        assert r.get_column_bool(0).tolist() == [True] and r.get_column_bool(1).tolist() == [False]
    # :299-315
    with make_result(batch_of(("s", ch.T_VARCHAR, ["hello"]))) as r:
        assert r.get_column_string(0) == ["hello"]
    # :318-336
    with make_result(batch_of(("range", ch.T_BIGINT, list(range(100))))) as r:
        v = r.get_column_int32(0)
        assert len(v) == 100 and v[0] == 0 and v[99] == 99
    # :343-371 [1,NULL,3,NULL,5]
    with make_result(batch_of(("x", ch.T_INTEGER, [1, None, 3, None, 5]))) as r:
        v, valid = r.get_column_int32_nullable(0)
        assert valid.tolist() == [True, False, True, False, True]
        assert v.tolist() == [1, 0, 3, 0, 5]
    # all null / no null
    with make_result(batch_of(("x", ch.T_INTEGER, [None, None, None]))) as r:
        v, valid = r.get_column_int32_nullable(0)
        assert valid.tolist() == [False] * 3 and v.tolist() == [0, 0, 0]
    # strings ['a',NULL,'c',NULL,'e']
    with make_result(batch_of(("s", ch.T_VARCHAR, ["a", None, "c", None, "e"]))) as r:
        s, valid = r.get_column_string_nullable(0)
        assert valid.tolist() == [True, False, True, False, True]
        # the reference test checks values[0] and values[2] only (:450-453): `total` omits the NULL rows'
        # terminators (src/duckdb_native.c:2719-2729), so the tail of the stream ('e') is overwritten by validity bytes
        assert s[0] == "a" and s[2] == "c" and len(s) == 5
    # doubles / bools with nulls
    with make_result(batch_of(("d", ch.T_DOUBLE, [1.5, None, 3.5]), ("b", ch.T_BOOLEAN, [1, None, 0]))) as r:
        d, dv = r.get_column_double_nullable(0)
        assert d.tolist() == [1.5, 0.0, 3.5] and dv.tolist() == [True, False, True]
        b, bv = r.get_column_bool_nullable(1)
        assert b.tolist() == [True, False, False] and bv.tolist() == [True, False, True]
    # schema type ids :128-203
    with make_result(batch_of(("a", ch.T_INTEGER, [1]), ("b", ch.T_BIGINT, [1]), ("c", ch.T_DOUBLE, [1.0]),
                               ("d", ch.T_BOOLEAN, [1]), ("e", ch.T_VARCHAR, ["x"]))) as r:
        assert [f.type_id for f in r.get_schema().fields] == ["int32", "int64", "double", "bool", "string"]


def test_reference_arrow_tests_through_gpu(ctx):
    reference_arrow_cases(lambda batch: _result(ctx, batch))


def test_decoder_row_cap(ctx):
    # decoders return [] above 1,000,000 rows (src/duckdb_arrow_native.mbt:435); the blob itself is complete
    n = 1_000_001
    batch = ch.config_c1(n)
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        blob = res.raw_column("int32", 0)
        assert blob == ora.get_column("int32", 0)
        assert res.get_column_int32(0).shape == (0,)
        assert oracle.decode_int32(blob)[0].shape == (0,)


# ------------------------------------------------------------------ Arrow C Data export (L1)
ARROW_DST = {ch.T_BOOLEAN: ch.D_BOOL_BITS, ch.T_DECIMAL: ch.D_I128, ch.T_HUGEINT: ch.D_I128, ch.T_INTERVAL: ch.D_MONTH_DAY_NANO}


def _check_arrow(ctx, batch, **kw):
    ora = oracle.OracleResult(batch)
    n = batch.nrows
    with _result(ctx, batch, **kw) as res:
        arrays = res.to_arrow()
        t = res.timings()
        assert t["h2d_bytes"] > 0 and t["d2h_bytes"] > 0
    # the arrays own their pinned buffers: still valid after close()
    assert len(arrays) == len(batch.columns)
    for j, (arr, col) in enumerate(zip(arrays, batch.columns)):
        assert len(arr) == n
        # random 128-bit HUGEINT payloads exceed 38 digits (DuckDB's own decimal128(38,0) export has the same limit)
        arr.validate(full=col.type_id != ch.T_HUGEINT)
        bufs = arr.buffers()
        if col.phys == ch.P_STRING:
            eo, ed = ora.arrow_string(j, 0)
            _, ebm, _, enc = ora.arrow_fixed(j, ch.D_SAME, 16, want_values=False)
            assert np.array_equal(np.frombuffer(bufs[1], dtype=np.int32)[: n + 1], eo)
            assert bytes(bufs[2])[: ed.shape[0]] == ed.tobytes()
        else:
            dst = ARROW_DST.get(col.type_id, ch.D_SAME)
            w = {ch.D_BOOL_BITS: 0, ch.D_I128: 16, ch.D_MONTH_DAY_NANO: 16}.get(dst, col.width)
            ev, ebm, _, enc = ora.arrow_fixed(j, dst, w)
            assert bytes(bufs[1])[: ev.shape[0]] == ev.tobytes(), f"values differ col={col.name}"
        assert arr.null_count == enc, col.name
        assert bytes(bufs[0])[: (n + 7) // 8] == ebm.tobytes()[: (n + 7) // 8], f"bitmap differs col={col.name}"
    return arrays


@pytest.mark.parametrize("n,pattern", [(1, "full"), (4096, "full"), (33_333, "ragged")])
def test_arrow_export_all_types(ctx, n, pattern):
    arrays = _check_arrow(ctx, _with_strings(n, pattern, 500 + n))
    # spot-check the logical types pyarrow sees
    types = [str(a.type) for a in arrays]
    assert types[0] == "bool" and types[3] == "int32" and "decimal128(38, 0)" in types[11]
    assert types[12] == "decimal128(4, 1)" and types[15] == "date32[day]" and types[16] == "timestamp[s]"
    assert types[19] == "month_day_nano_interval" and types[21] == "string"


def test_arrow_export_values_semantics(ctx):
    import datetime
    import decimal
    batch = batch_of(("d", ch.T_DATE, [19877, None, -1]),
                     ("ts", ch.T_TIMESTAMP, [1717418096789123, 0, None]),
                     ("dec", ch.T_DECIMAL, [123456, -99999999, None], 10, 3),
                     ("s", ch.T_VARCHAR, ["héllo", "", None]))
    with _result(ctx, batch) as res:
        a = res.to_arrow()
    assert a[0].to_pylist() == [datetime.date(2024, 6, 3), None, datetime.date(1969, 12, 31)]
    assert a[1].to_pylist()[0] == datetime.datetime(2024, 6, 3, 12, 34, 56, 789123)
    assert a[2].to_pylist() == [decimal.Decimal("123.456"), decimal.Decimal("-99999.999"), None]
    assert a[3].to_pylist() == ["héllo", "", None]


def test_arrow_export_configs(ctx):
    _check_arrow(ctx, ch.config_c1(100_000, variant_b=True))
    _check_arrow(ctx, ch.config_c1(100_000))
    _check_arrow(ctx, ch.config_c2(50_000))
    _check_arrow(ctx, ch.config_c3(50_000, pattern="ragged"))
    _check_arrow(ctx, ch.config_c4(20_000, ncols=12))


def test_scattered_heap_is_compacted_by_the_stager(ctx):
    # no registered heap: the stager gathers the pointed-to bytes itself (real DuckDB string heaps)
    rng = np.random.default_rng(9)
    many = [None if rng.random() < 0.1 else bytes(rng.integers(1, 255, int(rng.integers(0, 120)), dtype=np.uint8)) for _ in range(9000)]
    counts = ch.chunk_counts(len(many), "ragged", rng)
    batch = ch.ChunkBatch(counts, [ch.string_column("s", many, counts, type_id=ch.T_BLOB, shuffle_heap=rng)])
    ora = oracle.OracleResult(batch)
    eo, ed = ora.arrow_string(0, 0)
    with _result(ctx, batch, register_heap=False) as res:
        arr = res.to_arrow(0)
        assert str(arr.type) == "binary"
        assert np.array_equal(np.frombuffer(arr.buffers()[1], dtype=np.int32)[: len(many) + 1], eo)
        assert bytes(arr.buffers()[2])[: ed.shape[0]] == ed.tobytes()
        assert res.raw_column("string", 0, True) == ora.get_column("string", 0, True)


def test_inline_only_columns_skip_the_heap(ctx):
    # DMB_HEAP_INLINE_ONLY: the caller vouches that every string is inlined; nothing is staged or compacted,
    # the heap-less kernel runs (flags / codes: l_returnflag, l_shipmode shapes), NULLs and ragged chunks included
    rng = np.random.default_rng(21)
    n = 50_000
    counts = ch.chunk_counts(n, "ragged", rng)
    cols = [ch.string_column_bulk("flag", rng.integers(1, 2, n), None, counts, rng),
            ch.string_column_bulk("mode", rng.integers(0, 13, n), rng.random(n) > 0.2, counts, rng, utf8_fraction=0.1)]
    for c in cols:
        c.heap = None
        c.inline_only = True
    batch = ch.ChunkBatch(counts, cols)
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        for j in range(2):
            eo, ed = ora.arrow_string(j, 0)
            arr = res.to_arrow(j)
            assert np.array_equal(np.frombuffer(arr.buffers()[1], dtype=np.int32)[: n + 1], eo)
            assert bytes(arr.buffers()[2])[: ed.shape[0]] == ed.tobytes()
            arr.validate(full=True)
            assert res.raw_column("string", j, True) == ora.get_column("string", j, True)


def test_inline_only_contract_violation_is_an_error(ctx):
    # a pointer string in a column declared inline-only must fail loudly, never read through the pointer
    rng = np.random.default_rng(22)
    n = 5000
    counts = ch.chunk_counts(n, "full", rng)
    lens = rng.integers(0, 13, n)
    lens[1234] = 40
    col = ch.string_column_bulk("s", lens, None, counts, rng)
    col.heap = None
    col.inline_only = True
    with _result(ctx, ch.ChunkBatch(counts, [col])) as res:
        with pytest.raises(Exception, match="heap"):
            res.to_arrow(0)


def test_pinned_inputs_take_the_direct_dma_path(ctx):
    from duckdb_mbt_b200 import pinned
    batch = ch.config_c2(40_000)
    pb = pinned.pin_batch(batch)
    try:
        _check_arrow(ctx, pb, pinned=True)
    finally:
        pinned.free_batch(pb)


def test_record_batch_export(ctx):
    batch = ch.config_c1(5000, variant_b=True)
    with _result(ctx, batch) as res:
        rb = res.to_record_batch()
    assert rb.num_rows == 5000 and rb.num_columns == 3
    assert rb.column(0).to_numpy().tolist()[:3] == [0, 1, 2]
    assert rb.column(2).null_count == (5000 + 6) // 7


# ------------------------------------------------------------------ typed columns
def test_typed_columns_match_the_reference_text_round_trip(ctx):
    from duckdb_mbt_b200 import typed_result as tr
    n = 20_000
    rng = np.random.default_rng(77)
    counts = ch.chunk_counts(n, "ragged", rng)
    v = lambda p=0.2: rng.random(n) >= p  # noqa: E731
    cols = [
        ch.fixed_column("i32", ch.T_INTEGER, rng.integers(-2**31, 2**31, n).astype(np.int32), counts, valid=v()),
        ch.fixed_column("i64", ch.T_BIGINT, rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64), counts, valid=v()),
        ch.fixed_column("u64", ch.T_UBIGINT, rng.integers(0, 2**64 - 1, n, dtype=np.uint64), counts, valid=v()),
        ch.fixed_column("i8", ch.T_TINYINT, rng.integers(-128, 128, n).astype(np.int8), counts),
        ch.fixed_column("b", ch.T_BOOLEAN, rng.integers(0, 2, n).astype(np.uint8), counts, valid=v()),
        ch.fixed_column("d", ch.T_DOUBLE, rng.integers(-10**6, 10**6, n).astype(np.float64), counts, valid=v()),
        ch.fixed_column("date", ch.T_DATE, rng.integers(-700000, 2900000, n).astype(np.int32), counts, valid=v()),
        ch.fixed_column("ts", ch.T_TIMESTAMP, rng.integers(-10**15, 4 * 10**15, n, dtype=np.int64), counts, valid=v()),
        ch.fixed_column("ts_s", ch.T_TIMESTAMP_S, rng.integers(-10**9, 4 * 10**9, n, dtype=np.int64), counts, valid=v()),
        ch.fixed_column("ts_ms", ch.T_TIMESTAMP_MS, rng.integers(-10**12, 4 * 10**12, n, dtype=np.int64), counts),
        ch.fixed_column("ts_ns", ch.T_TIMESTAMP_NS, rng.integers(-10**18, 4 * 10**18, n, dtype=np.int64), counts, valid=v()),
    ]
    lens = rng.integers(0, 30, n)
    cols.append(ch.string_column_bulk("s", lens, v(), counts, rng, utf8_fraction=0.2))
    batch = ch.ChunkBatch(counts, cols)
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        typed = tr.to_typed(res)
        assert typed.row_count() == n and typed.column_count() == len(cols)
        for j, col in enumerate(cols[:-1]):
            tags, iv, dv = ora.typed_fixed(j)
            tc = typed.data[j]
            non_null = tags != tr.NULL
            assert np.array_equal(tc.valid, non_null), col.name
            if col.name in ("date",):
                # cells the reference's parse_date rejects (years outside 0001..9999) stay Value::String there;
                # compare where it yields a Date
                m = tags == tr.DATE
                assert m.sum() > n // 2
                assert np.array_equal(tc.values[m].astype(np.int64), iv[m]), col.name
                continue
            if col.name.startswith("ts"):
                m = tags == tr.TIMESTAMP
                assert m.sum() > n // 2
                assert np.array_equal(tc.values[m].astype(np.int64), iv[m]), col.name
                continue
            assert (tags[non_null] == tc.tag).all(), col.name
            if tc.tag == tr.DOUBLE:
                assert np.array_equal(tc.values[non_null], dv[non_null])
            else:
                assert np.array_equal(tc.values[non_null].astype(np.int64), iv[non_null]), col.name
        s = typed.data[len(cols) - 1]
        eo, ed = ora.arrow_string(len(cols) - 1, 0)
        assert np.array_equal(s.offsets, eo) and s.data == ed.tobytes()
        # reference accessors
        assert typed.get_int(0, 0) == (int(typed.data[0].values[0]) if typed.data[0].valid[0] else None)
        assert typed.get_value(-1, 0) is None and typed.is_null(n, 0)
        col0 = typed.get_int_column(0)
        assert len(col0) == n and (col0[5] is None) == (not typed.data[0].valid[5])
        assert typed.get_double_column(0) == [None] * n  # wrong variant -> None, like the reference


def test_typed_reference_vectors(ctx):
    # src/duckdb_test.mbt:1251-1287 BIGINT extremes saturate to Int32 min/max; :1175-1195 exact micros
    from duckdb_mbt_b200 import typed_result as tr
    batch = batch_of(("big", ch.T_BIGINT, [9223372036854775807, -9223372036854775808, 42, None]),
                     ("ts", ch.T_TIMESTAMP, [1717418096789123, None, 0, 1]),
                     ("d", ch.T_DATE, [19877, 0, -1, None]),
                     ("s", ch.T_VARCHAR, ["123", None, "", "x"]))
    with _result(ctx, batch) as res:
        t = tr.to_typed(res)
        assert t.get_int(0, 0) == 2147483647 and t.get_int(1, 0) == -2147483648 and t.get_int(2, 0) == 42
        assert t.is_null(3, 0)
        assert t.get_timestamp(0, 1) == 1717418096789123
        assert t.get_date(0, 2) == oracle.parse_date("2024-06-03") == 19877
        assert t.get_date(2, 2) == -1
        assert t.get_string(0, 3) == "123" and t.get_int(0, 3) is None  # VARCHAR that looks numeric stays String
        assert t.get_string(2, 3) == ""
    # DECIMAL stays a text-rendered Value::String in the reference (src/duckdb_parsing.mbt:124-127)
    with _result(ctx, batch_of(("dec", ch.T_DECIMAL, [1], 10, 3))) as res:
        assert tr.typed_column(res, 0).value(0).as_string() == "0.001"
    # UUID stays a text-rendered Value::String too (src/duckdb_parsing.mbt:122)
    one = ch.chunk_counts(1)
    zero = np.zeros((1, 16), np.uint8)
    zero[0, 15] = 0x80  # stored with the top bit flipped
    with _result(ctx, ch.ChunkBatch(one, [ch.fixed_column("h", ch.T_UUID, zero, one)])) as res:
        assert tr.typed_column(res, 0).value(0).as_string() == "00000000-0000-0000-0000-000000000000"


def test_aliased_string_pointers_are_sized_by_a_second_launch(ctx):
    """string_t entries of a flattened dictionary / constant vector point at the SAME heap bytes, so the column's bytes
    are not bounded by 12 n + heap_len: the kernels must not write past the first-guess buffer (round-1 advice) and the
    host sizes a second launch from the total the first one reports."""
    n = 50_000
    counts = ch.chunk_counts(n)
    words = [b"the quick brown fox jumps over the lazy dog", b"lorem ipsum dolor sit amet, consectetur", b"short"]
    col = ch.string_column("s", [words[i % 3] for i in range(3)] + [b"x"] * (n - 3), counts)
    ent = col.data.reshape(-1, 16)
    ent[:n] = ent[np.arange(n) % 3]  # every row is one of the three entries: 2/3 of them pointers into 82 heap bytes
    batch = ch.ChunkBatch(counts, [col])
    ora = oracle.OracleResult(batch)
    eo, ed = ora.arrow_string(0, 0)
    assert ed.shape[0] > 12 * n + col.heap.shape[0]
    with _result(ctx, batch) as res:
        (a,) = res.to_arrow()
        assert np.array_equal(np.frombuffer(a.buffers()[1], dtype=np.int32)[: n + 1], eo)
        assert bytes(a.buffers()[2])[: ed.shape[0]] == ed.tobytes()
        assert res.raw_column("string", 0, True) == ora.get_column("string", 0, True)
    # L0: the capacity check itself -- flag 8, nothing written past the capacity, exact total reported
    from duckdb_mbt_b200 import device
    db = device.DeviceBatch(batch)
    cap = 12 * n + col.heap.shape[0]
    so = db.plan_string(0, 0, data_capacity=cap)
    so.data[cap:] = 0xEE
    so.job.out_data_cap = cap
    db.run_string(so)
    assert db.string_error(so) & 8
    assert int(device.to_numpy(so.total, np.uint64)[0]) == ed.shape[0]
    assert bool((so.data[cap:] == 0xEE).all().item())


def test_string_spans_decoder_equals_the_string_decoder(ctx):
    batch = _with_strings(5000, "ragged", 77)
    col = len(batch.columns) - 2
    with _result(ctx, batch) as res:
        strs, valid = res.get_column_string_nullable(col)
        starts, ends, valid2, blob = res.get_column_string_spans_nullable(col)
        assert np.array_equal(valid, valid2)
        assert [bytes(blob[s:e]).decode("utf-8", errors="replace") for s, e in zip(starts, ends)] == strs


def test_blob_cells_read_as_their_varchar_cast(ctx):
    """BLOB through the text surfaces (string getter, per-cell value, typed Value::String): escaped like DuckDB's
    Blob::ToString; the Arrow export keeps the raw bytes (binary)."""
    from duckdb_mbt_b200 import native as nat
    from duckdb_mbt_b200 import typed_result as tr
    rng = np.random.default_rng(13)
    n = 6000
    vals = [None if rng.random() < 0.15 else bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8)) for _ in range(n)]
    vals[:3] = [b"abc", b"\x00\xff'q\"\\z", b""]
    counts = ch.chunk_counts(n, "ragged", rng)
    batch = ch.ChunkBatch(counts, [ch.string_column("b", vals, counts, type_id=ch.T_BLOB),
                                   ch.fixed_column("i", ch.T_INTEGER, np.arange(n, dtype=np.int32), counts)])
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        for nullable in (True, False):
            assert res.raw_column("string", 0, nullable) == ora.get_column("string", 0, nullable)
        L = res.lib
        L.duckdb_mb_result_value.restype = C.c_void_p
        for row in (0, 1, 2, 17, n - 1):
            assert nat.moonbit_bytes(L.duckdb_mb_result_value(C.c_void_p(res.handle), 0, row)) == ora.cell_value(0, row)
        arr = res.to_arrow(0)
        assert str(arr.type) == "binary" and arr.to_pylist() == vals


def test_a_lookback_that_gives_up_is_an_error_not_a_dead_context(ctx):
    """round-1 verdict (weak 13): the watchdog used to end in __trap(), which poisons the CUDA context for every result on
    that GPU.  With the wait limit forced to 0 a look-back that meets an unpublished predecessor gives up at once: the
    launch still ends, the host reports an error, and the SAME context converts the next result correctly."""
    from duckdb_mbt_b200 import arrow_result as ar
    L = ctx.lib
    L.dmb_dev_set_lookback_limit_ns.argtypes = [C.c_uint64]
    batch = ch.config_c3(400_000, seed=3)
    ora = oracle.OracleResult(batch)
    eo, ed = ora.arrow_string(0, 0)
    assert L.dmb_dev_set_lookback_limit_ns(0) == 0
    try:
        failures = 0
        for _ in range(5):  # (a look-back only waits when a predecessor is late: try a few times)
            try:
                with _result(ctx, batch) as res:
                    arr = res.to_arrow(0)
                assert np.array_equal(np.frombuffer(arr.buffers()[1], dtype=np.int32)[: eo.shape[0]], eo)  # no wait was needed: still exact
            except ar.DuckDBError as e:
                assert "look-back gave up" in str(e)
                failures += 1
        assert failures >= 1, "with a zero limit some of ~800 look-backs must have met an unpublished predecessor"
    finally:
        assert L.dmb_dev_set_lookback_limit_ns(4_000_000_000) == 0
    with _result(ctx, batch) as res:  # the context is alive and exact again
        arr = res.to_arrow(0)
        assert np.array_equal(np.frombuffer(arr.buffers()[1], dtype=np.int32)[: eo.shape[0]], eo)
        assert bytes(arr.buffers()[2])[: ed.shape[0]] == ed.tobytes()
