"""GPU parity of the reverse path (Arrow -> DataChunk vectors behind the appender) through the
C ABI with host Arrow buffers, against the CPU oracle; the appender protocol against the model of
src/duckdb_appender_state_machine.mbt; forward->reverse round trips."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
pa = pytest.importorskip("pyarrow")

import oracle  # noqa: E402
from duckdb_mbt_b200 import chunks as ch  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from duckdb_mbt_b200 import arrow_result as ar
    c = ar.GpuContext(0)
    yield c
    c.close()


def _c5_batch(n, seed, null_frac=0.1, max_len=24):
    """BASELINE.json configs[4] shape: int32 id, int64 v, float64 x, bool flag, utf8 s."""
    rng = np.random.default_rng(seed)
    mask = lambda: rng.random(n) < null_frac  # noqa: E731
    ids = pa.array(np.arange(n, dtype=np.int32))
    v = pa.array(rng.integers(-2**62, 2**62, n, dtype=np.int64), mask=mask())
    x = pa.array(rng.standard_normal(n), mask=mask())
    flag = pa.array(rng.random(n) < 0.5, mask=mask())
    lens = rng.integers(0, max_len + 1, n)
    body = rng.integers(0x20, 0x7F, int(lens.sum()) + 1, dtype=np.uint8).tobytes()
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    smask = mask()
    strs = [None if smask[i] else body[offs[i]:offs[i + 1]].decode() for i in range(n)]
    s = pa.array(strs, type=pa.string())
    return pa.record_batch([ids, v, x, flag, s], names=["id", "v", "x", "flag", "s"])


C5_TYPES = [ch.T_INTEGER, ch.T_BIGINT, ch.T_DOUBLE, ch.T_BOOLEAN, ch.T_VARCHAR]
REV_COPY = {1: 0, 2: 1, 4: 2, 8: 3, 16: 4}


def _oracle_column(arr: pa.Array, type_id: int):
    """expected (vector slab bytes, validity words, null count) for one Arrow array"""
    n = len(arr)
    bufs = arr.buffers()
    bitmap = None if (bufs[0] is None or arr.null_count == 0) else np.frombuffer(bufs[0], dtype=np.uint8)
    off = arr.offset
    if pa.types.is_string(arr.type) or pa.types.is_binary(arr.type):
        offsets = np.frombuffer(bufs[1], dtype=np.int32)[off: off + n + 1]
        data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(1, np.uint8)
        base = bufs[2].address if bufs[2] is not None and bufs[2].size else 0
        return oracle.rev_string(np.ascontiguousarray(offsets), data, base, bitmap, off, n), 16
    if pa.types.is_boolean(arr.type):
        vals = np.frombuffer(bufs[1], dtype=np.uint8)
        return oracle.rev_fixed(vals, bitmap, off, n, 5, 1), 1
    w = arr.type.bit_width // 8
    if pa.types.is_decimal(arr.type):
        p = arr.type.precision
        w_out, op = (2, 8) if p <= 4 else (4, 7) if p <= 9 else (8, 6) if p <= 18 else (16, 4)
        vals = np.frombuffer(bufs[1], dtype=np.uint8)[off * 16:]
        return oracle.rev_fixed(np.ascontiguousarray(vals), bitmap, off, n, op, w_out), w_out
    vals = np.frombuffer(bufs[1], dtype=np.uint8)[off * w:]
    return oracle.rev_fixed(np.ascontiguousarray(vals), bitmap, off, n, REV_COPY[w], w), w


def _append_and_check(ctx, rb: pa.RecordBatch, type_ids, dec_widths=None, env_rows=None, monkeypatch=None):
    from duckdb_mbt_b200 import appender as ap
    if env_rows is not None:
        monkeypatch.setenv("DMB_REV_BATCH_ROWS", str(env_rows))
    sink = ap.CollectedChunks(type_ids, dec_widths)
    a = ap.Appender(ctx, type_ids, sink)
    assert a.state == ap.READY
    struct = rb.to_struct_array()  # keep the exact buffers alive: pointer string_t refer to them in place
    a.append_arrow(struct)
    assert a.state == ap.READY and a.row_count == rb.num_rows
    a.flush()
    assert a.state == ap.FLUSHED and a.flushed_row_count == rb.num_rows
    n = rb.num_rows
    assert sink.nrows == n
    assert all(c == ch.VECTOR_SIZE for c in sink.counts[:-1])
    for c in range(rb.num_columns):
        (exp_out, exp_val, _), w = _oracle_column(struct.field(c), type_ids[c])
        got = sink.column_bytes(c)
        assert got == exp_out.tobytes()[: n * w], f"vector payload differs col={c}"
        got_val = np.concatenate(sink.validity[c]) if sink.validity[c] else np.zeros(0, np.uint64)
        assert np.array_equal(got_val, exp_val[: got_val.shape[0]]), f"validity masks differ col={c}"
    a.close()
    return sink


@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 10_000, 100_003])
def test_append_arrow_c5_shape(ctx, n):
    _append_and_check(ctx, _c5_batch(n, 40 + n), C5_TYPES)


@pytest.mark.parametrize("offset", [1, 5, 64, 2047, 2051])
def test_append_arrow_sliced_batches_have_bit_offsets(ctx, offset):
    rb = _c5_batch(20_000, 9).slice(offset, 20_000 - offset - 3)
    _append_and_check(ctx, rb, C5_TYPES)


def test_append_arrow_in_sub_batches(ctx, monkeypatch):
    # several pipelined sub-batches (double-buffered slots) must give the same chunks
    _append_and_check(ctx, _c5_batch(50_000, 3).slice(3, 49_990), C5_TYPES, env_rows=4096, monkeypatch=monkeypatch)


def test_append_arrow_more_types(ctx):
    import decimal
    rng = np.random.default_rng(8)
    n = 9000
    m = lambda: rng.random(n) < 0.2  # noqa: E731
    cols = [
        pa.array(rng.integers(-128, 128, n).astype(np.int8), mask=m()),
        pa.array(rng.integers(0, 2**16, n).astype(np.uint16), mask=m()),
        pa.array(rng.standard_normal(n).astype(np.float32), mask=m()),
        pa.array(rng.integers(-10**5, 10**5, n).astype(np.int32), mask=m()).cast(pa.date32()),
        pa.array(rng.integers(0, 2 * 10**15, n, dtype=np.int64), mask=m()).cast(pa.timestamp("us")),
        pa.array([None if rng.random() < 0.2 else decimal.Decimal(int(rng.integers(-10**17, 10**17))) / 1000 for _ in range(n)],
                 type=pa.decimal128(18, 3)),
        pa.array([decimal.Decimal(int(rng.integers(-10**8, 10**8))) / 100 for _ in range(n)], type=pa.decimal128(9, 2)),
        pa.array([None if rng.random() < 0.2 else decimal.Decimal(int(rng.integers(-2**62, 2**62))) * 10**15 for _ in range(n)],
                 type=pa.decimal128(38, 0)),
        pa.array([None if rng.random() < 0.3 else bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8)) for _ in range(n)],
                 type=pa.binary()),
    ]
    types = [ch.T_TINYINT, ch.T_USMALLINT, ch.T_FLOAT, ch.T_DATE, ch.T_TIMESTAMP, ch.T_DECIMAL, ch.T_DECIMAL, ch.T_HUGEINT, ch.T_BLOB]
    rb = pa.record_batch(cols, names=[f"c{i}" for i in range(len(cols))])
    _append_and_check(ctx, rb.slice(11, n - 20), types, dec_widths=[0, 0, 0, 0, 0, 18, 9, 0, 0])


def test_forward_then_reverse_round_trip(ctx):
    """DataChunks -> Arrow (forward kernels) -> DataChunks (reverse kernels): payload of valid rows,
    validity masks and strings survive; NULL payloads come back zeroed."""
    from duckdb_mbt_b200 import appender as ap
    from duckdb_mbt_b200 import arrow_result as ar
    n = 150_000
    rng = np.random.default_rng(21)
    counts = ch.chunk_counts(n)
    valid = rng.random(n) > 0.25
    i64 = rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64)
    lens = rng.integers(0, 40, n)
    batch = ch.ChunkBatch(counts, [
        ch.fixed_column("i", ch.T_BIGINT, i64, counts, valid=valid, garbage_rng=rng),
        ch.fixed_column("b", ch.T_BOOLEAN, rng.integers(0, 2, n).astype(np.uint8), counts, valid=valid, garbage_rng=rng),
        ch.string_column_bulk("s", lens, valid, counts, rng),
    ])
    with ar.ArrowResult.from_chunks(ctx, batch) as res:
        rb = res.to_record_batch()
    types = [ch.T_BIGINT, ch.T_BOOLEAN, ch.T_VARCHAR]
    sink = ap.CollectedChunks(types)
    a = ap.Appender(ctx, types, sink)
    struct = rb.to_struct_array()
    a.append_arrow(struct)
    a.close()
    assert sink.nrows == n
    back = np.frombuffer(sink.column_bytes(0), dtype=np.int64)
    assert np.array_equal(back, np.where(valid, i64, 0))
    assert np.array_equal(sink.valid_bits(0), valid)
    assert np.array_equal(sink.valid_bits(2), valid)
    # strings: decode the returned string_t (inline or pointer into the Arrow data buffer)
    orig = ch.string_values(batch.columns[2], counts)
    ent = np.frombuffer(sink.column_bytes(2), dtype=np.uint8).reshape(-1, 16)
    import ctypes
    for i in rng.integers(0, n, 3000):
        e = ent[i]
        L = int(e[0:4].view(np.uint32)[0])
        if orig[i] is None:
            assert not e.any()
            continue
        got = bytes(e[4:4 + L]) if L <= 12 else ctypes.string_at(int(e[8:16].view(np.uint64)[0]), L)
        assert got == orig[i]
    del struct


# ------------------------------------------------------------------ protocol
def test_row_api_matches_bulk_path_and_reference_tests(ctx):
    """src/duckdb_test.mbt:478-799 style: int/varchar/double/bool/null/bigint/multi-row."""
    from duckdb_mbt_b200 import appender as ap
    types = [ch.T_INTEGER, ch.T_VARCHAR, ch.T_DOUBLE, ch.T_BOOLEAN, ch.T_BIGINT, ch.T_DATE, ch.T_TIMESTAMP]
    sink = ap.CollectedChunks(types)
    a = ap.Appender(ctx, types, sink)
    rows = [(1, "Alice", 1.5, True, 2**40, 19877, 1717418096789123), (2, None, None, False, None, None, None),
            (None, "a much longer string than twelve bytes", -0.0, None, -1, -1, 0)] * 900
    for r in rows:
        a.begin_row()
        a.append_int(r[0]) if r[0] is not None else a.append_null()
        a.append_varchar(r[1]) if r[1] is not None else a.append_null()
        a.append_double(r[2]) if r[2] is not None else a.append_null()
        a.append_bool(r[3]) if r[3] is not None else a.append_null()
        a.append_bigint(r[4]) if r[4] is not None else a.append_null()
        a.append_date(r[5]) if r[5] is not None else a.append_null()
        a.append_timestamp(r[6]) if r[6] is not None else a.append_null()
        a.end_row()
    assert a.state == ap.READY and a.row_count == len(rows) and sink.nrows == 0  # still buffered
    a.flush()
    assert a.state == ap.FLUSHED and sink.nrows == len(rows)
    ints = np.frombuffer(sink.column_bytes(0), dtype=np.int32)
    assert ints[:3].tolist() == [1, 2, 0] and sink.valid_bits(0)[:3].tolist() == [True, True, False]
    dbl = np.frombuffer(sink.column_bytes(2), dtype=np.float64)
    assert dbl[0] == 1.5 and sink.valid_bits(2)[:3].tolist() == [True, False, True]
    assert np.frombuffer(sink.column_bytes(3), dtype=np.uint8)[:3].tolist() == [1, 0, 0]
    assert np.frombuffer(sink.column_bytes(4), dtype=np.int64)[:3].tolist() == [2**40, 0, -1]
    assert np.frombuffer(sink.column_bytes(5), dtype=np.int32)[0] == 19877  # exact days (not the reference's approximate string path)
    assert np.frombuffer(sink.column_bytes(6), dtype=np.int64)[0] == 1717418096789123
    ent = np.frombuffer(sink.column_bytes(1), dtype=np.uint8).reshape(-1, 16)
    assert bytes(ent[0, 4:9]) == b"Alice" and int(ent[0, 0:4].view(np.uint32)[0]) == 5
    assert int(ent[2, 0:4].view(np.uint32)[0]) == 38 and bytes(ent[2, 4:8]) == b"a mu"
    a.close()


def test_row_api_blob_decimal_interval(ctx):
    """src/duckdb_native.c:1397-1415 (blob), :1447-1481 (decimal from hugeint parts), :1511-1533 (interval):
    the cells land in the vectors as DuckDB stores them (DECIMAL narrowed to the precision's physical width)."""
    from duckdb_mbt_b200 import appender as ap
    from duckdb_mbt_b200.arrow_result import DuckDBError
    types = [ch.T_BLOB, ch.T_DECIMAL, ch.T_DECIMAL, ch.T_DECIMAL, ch.T_INTERVAL]
    sink = ap.CollectedChunks(types, [0, 15, 38, 4, 0])
    a = ap.Appender(ctx, types, sink)
    a.set_decimal(1, 15, 2)
    a.set_decimal(2, 38, 0)
    rng = np.random.default_rng(5)
    rows = []
    for i in range(3000):
        blob = None if i % 11 == 0 else bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8))
        d15 = None if i % 7 == 0 else int(rng.integers(-10**15 + 1, 10**15))
        d38 = None if i % 5 == 0 else (int.from_bytes(rng.bytes(16), "little") % 10**38) * (1 if i % 2 else -1)
        d4 = int(rng.integers(-9999, 10000))
        iv = None if i % 13 == 0 else (int(rng.integers(-100, 100)), int(rng.integers(-400, 400)), int(rng.integers(-2**50, 2**50)))
        rows.append((blob, d15, d38, d4, iv))
        a.begin_row()
        a.append_blob(blob) if blob is not None else a.append_null()
        a.append_decimal(15, 2, d15) if d15 is not None else a.append_null()
        a.append_decimal(38, 0, d38) if d38 is not None else a.append_null()
        a.append_decimal(4, 1, d4)  # the first value declares DECIMAL(4,1): physical int16
        a.append_interval(*iv) if iv is not None else a.append_null()
        a.end_row()
    a.flush()
    n = len(rows)
    assert sink.nrows == n
    ent = np.frombuffer(sink.column_bytes(0), dtype=np.uint8).reshape(-1, 16)
    vb = sink.valid_bits(0)
    for i in (1, 2, 3, 500, 2999):
        blob = rows[i][0]
        assert vb[i] == (blob is not None)
        if blob is not None:
            assert int(ent[i, 0:4].view(np.uint32)[0]) == len(blob)
            assert bytes(ent[i, 4:4 + min(len(blob), 12 if len(blob) <= 12 else 4)]) == blob[: 12 if len(blob) <= 12 else 4]
    d15 = np.frombuffer(sink.column_bytes(1), dtype=np.int64)
    assert d15.tolist() == [0 if r[1] is None else r[1] for r in rows]
    assert sink.valid_bits(1).tolist() == [r[1] is not None for r in rows]
    d38 = np.frombuffer(sink.column_bytes(2), dtype=np.uint64).reshape(-1, 2)
    got38 = [(int(lo) | (int(hi) << 64)) - ((1 << 128) if int(hi) >> 63 else 0) for lo, hi in d38]
    assert got38 == [0 if r[2] is None else r[2] for r in rows]
    assert np.frombuffer(sink.column_bytes(3), dtype=np.int16).tolist() == [r[3] for r in rows]
    iv = np.frombuffer(sink.column_bytes(4), dtype=np.dtype([("m", "<i4"), ("d", "<i4"), ("us", "<i8")]))
    assert [tuple(int(x) for x in v) for v in iv] == [(0, 0, 0) if r[4] is None else r[4] for r in rows]
    # decimal -> decimal cast on the way in: scale up exactly, scale down rounding half away from zero, range checked
    a.begin_row()
    a.append_null()
    a.append_decimal(5, 1, 123)       # 12.3 -> DECIMAL(15,2): 1230
    a.append_decimal(10, 3, -12345)   # -12.345 -> DECIMAL(38,0): -12
    a.append_decimal(6, 2, 1255)      # 12.55 -> DECIMAL(4,1): 126 (12.6)
    a.append_null()
    a.end_row()
    a.flush()
    assert np.frombuffer(sink.column_bytes(1), dtype=np.int64)[n] == 1230
    assert np.frombuffer(sink.column_bytes(2), dtype=np.int64).reshape(-1, 2)[n].tolist() == [-12, -1]
    assert np.frombuffer(sink.column_bytes(3), dtype=np.int16)[n] == 126
    a.begin_row()
    a.append_null()
    with pytest.raises(DuckDBError, match="out of range"):
        a.append_decimal(18, 2, 10**17)  # 10^15 does not fit DECIMAL(15,2)
    assert a.state == ap.ERROR
    a.close()
    b = ap.Appender(ctx, [ch.T_INTEGER], discard=True)
    b.begin_row()
    with pytest.raises(DuckDBError):
        b.append_interval(1, 2, 3)  # type mismatch: 0 + per-handle error, state Error
    assert b.state == ap.ERROR
    b.close()


def test_protocol_follows_the_reference_model(ctx):
    from duckdb_mbt_b200 import appender as ap
    from duckdb_mbt_b200.arrow_result import DuckDBError
    rng = np.random.default_rng(2026)
    cmds = ["begin_row", "append_int", "append_int", "append_int", "end_row", "end_row", "flush", "close", "begin_row"]
    for trial in range(150):
        a = ap.Appender(ctx, [ch.T_INTEGER, ch.T_INTEGER], discard=True)
        model = ap.AppenderModel(2).execute("create")
        for _ in range(int(rng.integers(1, 25))):
            cmd = cmds[int(rng.integers(0, len(cmds)))]
            before = model.state
            model.execute(cmd)
            try:
                if cmd == "close":
                    a.lib.duckdb_mb_gpu_appender_close(a.handle)
                elif cmd == "append_int":
                    a.append_int(7)
                else:
                    getattr(a, cmd)()
                ok = True
            except DuckDBError:
                ok = False
            assert a.state == model.state, f"trial {trial}: {ap.STATE_NAMES[before]} --{cmd}--> C={ap.STATE_NAMES[a.state]} model={ap.STATE_NAMES[model.state]}"
            if model.state not in (ap.ERROR, ap.CLOSED):
                assert ok and a.row_count == model.row_count
            if not ok:
                assert a.error() != ""
        a.lib.duckdb_mb_gpu_appender_destroy(a.handle)
        a.handle = None


def test_append_arrow_errors_keep_the_reference_convention(ctx):
    from duckdb_mbt_b200 import appender as ap
    from duckdb_mbt_b200.arrow_result import DuckDBError
    a = ap.Appender(ctx, [ch.T_INTEGER, ch.T_VARCHAR], discard=True)
    rb = pa.record_batch([pa.array([1, 2, 3], type=pa.int64()), pa.array(["a", "b", "c"])], names=["a", "b"])
    with pytest.raises(DuckDBError, match="cannot be appended"):
        a.append_arrow(rb)  # int64 into INTEGER: type mismatch -> 0 + per-handle error string, state Error
    assert a.state == ap.ERROR
    with pytest.raises(DuckDBError):
        a.flush()
    a.close()
    b = ap.Appender(ctx, [ch.T_INTEGER], discard=True)
    with pytest.raises(DuckDBError, match="columns"):
        b.append_arrow(rb)
    b.close()
    c = ap.Appender(ctx, [ch.T_INTEGER], discard=True)
    c.begin_row()
    with pytest.raises(DuckDBError):  # a batch inside a row is as illegal as BeginRow there
        c.append_arrow(pa.record_batch([pa.array([1], type=pa.int32())], names=["a"]))
    assert c.state == ap.ERROR
    c.close()


class _StringSink:
    """sink that resolves every VARCHAR cell while the chunk is handed over (pointer string_t refer to the appender's
    own row buffer, valid during the call -- DuckDB copies them inside duckdb_append_data_chunk)"""

    def __init__(self, ncols):
        self.cells = [[] for _ in range(ncols)]

    def __call__(self, count, vec_data, vec_validity):
        import ctypes as C
        for c in range(len(self.cells)):
            ent = np.frombuffer(C.string_at(vec_data[c], count * 16), dtype=np.uint8).reshape(-1, 16)
            mask = np.frombuffer(C.string_at(vec_validity[c], 8 * ch.VALIDITY_WORDS), dtype=np.uint64)
            for i in range(count):
                if not (int(mask[i >> 6]) >> (i & 63)) & 1:
                    self.cells[c].append(None)
                    continue
                ln = int(ent[i, 0:4].view(np.uint32)[0])
                if ln <= 12:
                    self.cells[c].append(bytes(ent[i, 4:4 + ln]))
                else:
                    self.cells[c].append(C.string_at(int(ent[i, 8:16].view(np.uint64)[0]), ln))
        return True


def test_list_struct_map_cells_are_the_reference_text_forms(ctx):
    """Appender::append_list_varchar / append_struct / append_map (src/duckdb_native.mbt:1703-1755 over
    src/duckdb_native.c:1735-1926) and append_list_varchar_value (:1764-1795, the reference's own test
    src/duckdb_test.mbt:1391-1425): one VARCHAR cell holding the serialised text, against the oracle's restatement."""
    from duckdb_mbt_b200 import appender as ap
    rng = np.random.default_rng(21)
    sink = _StringSink(3)
    a = ap.Appender(ctx, [ch.T_VARCHAR, ch.T_VARCHAR, ch.T_VARCHAR], sink)
    word = lambda: bytes(rng.integers(0x20, 0x7F, int(rng.integers(0, 9)), dtype=np.uint8)).decode()  # noqa: E731
    rows = []
    for i in range(2500):
        items = [word() for _ in range(int(rng.integers(0, 5)))]
        keys = [word() for _ in range(int(rng.integers(0, 4)))]
        vals = [word() for _ in keys]
        rows.append((items, keys, vals))
        a.begin_row()
        a.append_list_varchar(items)
        a.append_struct(keys, vals) if i % 2 else a.append_map(keys, vals)
        a.append_list_varchar_value(items)
        a.end_row()
    # known answers + the C-string cut at an embedded NUL
    a.begin_row()
    a.append_list_varchar(["a", "b", "c"])
    a.append_struct(["name", "age"], ["duck", "3"])
    a.append_list_varchar_value(["a", "it's", "c"])
    a.end_row()
    a.begin_row()
    a.append_list_varchar([b"ab\0cd", b"x"])
    a.append_map([], [])
    a.append_list_varchar_value([])
    a.end_row()
    assert a.state == ap.READY
    a.flush()
    n = len(rows)
    assert len(sink.cells[0]) == n + 2
    for i, (items, keys, vals) in enumerate(rows):
        assert sink.cells[0][i] == oracle.list_varchar_text(items), i
        assert sink.cells[1][i] == oracle.pairs_varchar_text(keys, vals), i
        assert sink.cells[2][i] == ("[" + ", ".join("'" + v.replace("'", "''") + "'" for v in items) + "]").encode(), i
    assert [c[n] for c in sink.cells] == [b'["a", "b", "c"]', b'{"name": "duck", "age": "3"}', b"['a', 'it''s', 'c']"]
    assert [c[n + 1] for c in sink.cells] == [b'["ab', b"{}", b"[]"]
    # wrong column type: the reference's mutator convention (0 + per-handle error), state -> Error
    b = ap.Appender(ctx, [ch.T_INTEGER], _StringSink(1))
    b.begin_row()
    with pytest.raises(Exception):
        b.append_list_varchar(["x"])
    b.close()
    a.close()
