"""GPU parity of the reference's per-cell paths through the drop-in C symbols:
`duckdb_mb_result_{column_count,row_count,column_name,column_type,is_null,value}` (Connection::query,
src/duckdb_native.mbt:454-501 over src/duckdb_native.c:174-238) and the streaming set
`duckdb_mb_stream_* / duckdb_mb_chunk_*` (ResultStream::next, src/duckdb_native.mbt:504-582 over
src/duckdb_native.c:260-667), against the oracle's per-cell restatement and the reference's own
stream test (src/duckdb_test.mbt:97-113)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from duckdb_mbt_b200 import chunks as ch  # noqa: E402

from test_gpu_l0_parity import _mixed_batch  # noqa: E402
from test_oracle_golden import batch_of  # noqa: E402

RENDERED = {"b", "i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64", "f32", "f64", "huge", "dec4", "dec9", "dec18", "date", "ts_s", "ts_ms", "ts_ns", "iv", "uuid"}


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


@pytest.mark.parametrize("n,pattern", [(1, "full"), (2500, "full"), (6000, "ragged")])
def test_result_cells_match_the_reference_loop(ctx, n, pattern):
    from duckdb_mbt_b200 import native as nat
    batch = _mixed_batch(n, pattern, 40 + n)
    rng = np.random.default_rng(n)
    strings = [None if rng.random() < 0.1 else bytes(rng.integers(0x20, 0x7F, int(l), dtype=np.uint8)) for l in rng.integers(0, 30, n)]
    if n > 3:
        strings[1] = b"in\0ner"  # strlen truncation of duckdb_value_varchar's C string
        strings[2] = b""
    batch.columns.append(ch.string_column("s", strings, batch.counts))
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        L, h = res.lib, res.handle
        assert L.duckdb_mb_result_column_count(h) == len(batch.columns) and L.duckdb_mb_result_row_count(h) == n
        assert not L.duckdb_mb_is_null_result(h) and L.duckdb_mb_is_null_result(None)
        rows = sorted(set([0, n - 1, n // 2] + rng.integers(0, n, 60).tolist()))
        for j, col in enumerate(batch.columns):
            assert nat.moonbit_bytes(L.duckdb_mb_result_column_name(h, j)).decode() == col.name
            assert L.duckdb_mb_result_column_type(h, j) == col.type_id
            if col.name not in RENDERED and col.name != "s":
                continue
            for r in rows:
                assert bool(L.duckdb_mb_result_is_null(h, j, r)) == ora.cell_is_null(j, r), (col.name, r)
                assert nat.moonbit_bytes(L.duckdb_mb_result_value(h, j, r)) == ora.cell_value(j, r), (col.name, r)
        # out of range: NULL / empty Bytes / INVALID, like the reference's guards
        assert L.duckdb_mb_result_is_null(h, 0, n) == 1 and L.duckdb_mb_result_is_null(h, -1, 0) == 1
        assert nat.moonbit_bytes(L.duckdb_mb_result_value(h, 0, n)) == b"" and nat.moonbit_bytes(L.duckdb_mb_result_value(h, 99, 0)) == b""
        assert L.duckdb_mb_result_column_type(h, 99) == 0 and nat.moonbit_bytes(L.duckdb_mb_result_column_name(h, 99)) == b""


def test_query_per_cell_equals_the_columnar_form(ctx):
    from duckdb_mbt_b200.query_result import QueryResult, query_per_cell
    batch = batch_of(("id", ch.T_INTEGER, [1, 2, None]), ("name", ch.T_VARCHAR, ["a", None, "ccc"]),
                     ("d", ch.T_DATE, [19877, None, 0]), ("p", ch.T_DECIMAL, [1050, 99, None], 10, 2))
    with _result(ctx, batch) as res:
        q = query_per_cell(res)
        assert q.rows == [["1", "a", "2024-06-03", "10.50"], ["2", "", "", "0.99"], ["", "ccc", "1970-01-01", ""]]
        assert q.nulls == [[False, False, False, False], [False, True, True, False], [True, False, False, True]]
        assert q.columns == ["id", "name", "d", "p"] and q.column_types == [ch.T_INTEGER, ch.T_VARCHAR, ch.T_DATE, ch.T_DECIMAL]
        c = QueryResult.from_result(res, q.column_types)
        assert c.rows == q.rows and c.nulls == q.nulls


def test_stream_large_range(ctx):
    # src/duckdb_test.mbt:97-113: SELECT i FROM RANGE(1000000) tbl(i) streamed, rows counted per chunk
    from duckdb_mbt_b200 import native as nat
    n = 1_000_000
    counts = ch.chunk_counts(n)
    batch = ch.ChunkBatch(counts, [ch.fixed_column("i", ch.T_BIGINT, np.arange(n, dtype=np.int64), counts)])
    with _result(ctx, batch) as res:
        L = res.lib
        s = L.duckdb_mb_gpu_stream_from_result(res.handle)
        assert not L.duckdb_mb_is_null_stream(s)
        assert L.duckdb_mb_stream_column_count(s) == 1 and nat.moonbit_bytes(L.duckdb_mb_stream_column_name(s, 0)) == b"i"
        total, k = 0, 0
        while True:
            c = L.duckdb_mb_stream_fetch_chunk(s)
            if L.duckdb_mb_is_null_chunk(c):
                assert nat.last_error() == ""  # end of stream, not an error
                break
            rows = L.duckdb_mb_chunk_row_count(c)
            assert rows == int(counts[k]) and L.duckdb_mb_chunk_column_count(c) == 1
            if k in (0, 17, len(counts) - 1):
                assert L.duckdb_mb_chunk_is_null(c, 0, rows - 1) == 0
                assert nat.moonbit_bytes(L.duckdb_mb_chunk_value(c, 0, rows - 1)) == str(total + rows - 1).encode()
            total += rows
            k += 1
            L.duckdb_mb_chunk_destroy(c)
        assert total == n and k == len(counts)
        L.duckdb_mb_stream_destroy(s)


def test_stream_chunks_nulls_and_whitelist(ctx):
    from duckdb_mbt_b200.arrow_result import DuckDBError
    from duckdb_mbt_b200.query_result import ResultStream
    rng = np.random.default_rng(5)
    n = 5000
    counts = ch.chunk_counts(n, "ragged", rng)
    valid = rng.random(n) > 0.3
    vals = rng.integers(-1000, 1000, n).astype(np.int32)
    lens = rng.integers(0, 20, n)
    batch = ch.ChunkBatch(counts, [ch.fixed_column("v", ch.T_INTEGER, vals, counts, valid=valid, garbage_rng=rng),
                                   ch.string_column_bulk("s", lens, rng.random(n) > 0.2, counts, rng)])
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        st = ResultStream(res)
        assert st.columns() == ["v", "s"]
        row0 = 0
        seen = 0
        while True:
            chunk = st.next()
            if chunk is None:
                break
            for r in (0, chunk.row_count() - 1):
                if chunk.row_count() == 0:
                    continue
                for c in range(2):
                    assert chunk.nulls[r][c] == ora.cell_is_null(c, row0 + r)
                    assert chunk.rows[r][c].encode() == ora.cell_value(c, row0 + r)
            row0 += chunk.row_count()
            seen += 1
        assert row0 == n and seen == len(counts)
        st.close()
    # DECIMAL is not on the reference's streaming whitelist (src/duckdb_native.c:271-303)
    with _result(ctx, batch_of(("dec", ch.T_DECIMAL, [1], 10, 3))) as res:
        with pytest.raises(DuckDBError, match="unsupported column type"):
            ResultStream(res)


def test_c1_string_form_and_typed(ctx):
    # BASELINE.json configs[0]: SELECT i::INTEGER, i::DOUBLE, CASE WHEN i%7=0 THEN NULL [ELSE i] END FROM range(n):
    # the string form Connection::query builds (every column rendered on the device) and the typed form
    from duckdb_mbt_b200 import typed_result as tr
    from duckdb_mbt_b200.query_result import QueryResult
    n = 30_000
    for variant_b in (False, True):
        batch = ch.config_c1(n, variant_b=variant_b)
        ora = oracle.OracleResult(batch)
        with _result(ctx, batch) as res:
            q = QueryResult.from_result(res, [c.type_id for c in batch.columns])
            assert q.row_count() == n and q.column_count() == 3
            for r in (0, 1, 6, 7, 14, 999, n - 1):
                for c in range(3):
                    assert q.cell(r, c) == (None if ora.cell_is_null(c, r) else ora.cell_value(c, r).decode()), (r, c)
            assert q.cell(7, 0) == "7" and q.cell(7, 1) == "7.0" and q.cell(7, 2) is None
            t = q.to_typed()
            assert t.get_int(n - 1, 0) == n - 1 and t.get_double(n - 1, 1) == float(n - 1) and t.is_null(7, 2)
