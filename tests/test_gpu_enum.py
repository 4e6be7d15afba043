"""ENUM columns: uint8/16/32 index vectors + the type's dictionary.
The reference keeps an ENUM cell as Value::String of its label (src/duckdb_parsing.mbt:119-122), gets the label
through duckdb_value_varchar (src/duckdb_native.c:215-238, :2474-2510) and names the column "string" in the Arrow
schema JSON (:2314-2339).  Here: one lookup kernel (kernels_enum.cu) + the string kernels with the dictionary as
the heap; the Arrow C Data export is dictionary-encoded.  Checked against the oracle's restatement (UNPINNED in
the reference: no reference test holds an ENUM column) and with pyarrow."""
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


def _result(ctx, batch, **kw):
    from duckdb_mbt_b200 import arrow_result as ar
    return ar.ArrowResult.from_chunks(ctx, batch, **kw)


def _labels(k, rng, max_len):
    out = []
    for i in range(k):
        ln = int(rng.integers(1, max_len + 1))
        body = rng.integers(0x61, 0x7B, max_len, dtype=np.uint8).tobytes()
        out.append((b"%d_" % i + body)[:ln])
    if k > 2:
        out[1] = b""  # an empty label is legal bytes-wise and exercises zero-length entries
    return out


def _enum_batch(n, k, max_len, pattern, seed, null_frac=0.2):
    rng = np.random.default_rng(seed)
    counts = ch.chunk_counts(n, pattern, rng)
    labels = _labels(k, rng, max_len)
    idx = rng.integers(0, k, n)
    valid = rng.random(n) >= null_frac if null_frac else None
    # garbage under NULL rows stays inside the index type but may point past the dictionary
    col = ch.enum_column("e", labels, idx, counts, valid=valid, garbage_rng=rng if valid is not None else None)
    other = ch.fixed_column("i", ch.T_INTEGER, np.arange(n, dtype=np.int32), counts)
    return ch.ChunkBatch(counts, [col, other]), labels, idx, valid


@pytest.mark.parametrize("n,k,max_len,pattern", [
    (1, 1, 3, "full"), (2049, 3, 8, "full"), (10_000, 7, 12, "ragged"),      # every label inlined: the heap-less kernel
    (40_001, 5, 30, "ragged"), (100_000, 300, 20, "full"),                      # uint16 indices, labels in the heap
    (30_000, 70_000, 9, "ragged"), (20_000, 4, 200, "full")])                   # uint32 indices; labels longer than a render slot
def test_enum_string_getter_and_text_column(ctx, n, k, max_len, pattern):
    from duckdb_mbt_b200 import typed_result as tr
    batch, labels, idx, valid = _enum_batch(n, k, max_len, pattern, 77 + n + k)
    assert batch.columns[0].phys == (ch.P_U8 if k <= 256 else ch.P_U16 if k <= 65536 else ch.P_U32)
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        for nullable in (False, True):
            assert res.raw_column("string", 0, nullable) == ora.get_column("string", 0, nullable), f"nullable={nullable}"
        for fn in (tr.typed_column, tr.text_column):   # to_typed keeps ENUM as Value::String
            tc = fn(res, 0)
            assert tc.tag == tr.STRING
            assert np.array_equal(tc.valid, np.ones(n, bool) if valid is None else valid)
            for i in range(0, n, max(1, n // 97)):
                v = tc.value(i)
                if valid is None or valid[i]:
                    assert v.as_string().encode() == labels[idx[i]]
                else:
                    assert v.is_null()
        schema = res.get_schema()
        assert [f.type_id for f in schema.fields][0] == "string"


@pytest.mark.parametrize("n,k,max_len", [(5000, 3, 6), (70_001, 1000, 25), (4097, 66_000, 5)])
def test_enum_arrow_export_is_dictionary_encoded(ctx, n, k, max_len):
    batch, labels, idx, valid = _enum_batch(n, k, max_len, "ragged", 5 + k)
    with _result(ctx, batch) as res:
        arr = res.to_arrow(0)
        arr.validate(full=True)
        assert pa.types.is_dictionary(arr.type)
        assert arr.type.index_type == (pa.uint8() if k <= 256 else pa.uint16() if k <= 65536 else pa.uint32())
        assert arr.type.value_type == pa.string()
        assert arr.dictionary.to_pylist() == [x.decode() for x in labels]
        got_idx = np.asarray(arr.indices.fill_null(0))
        exp_idx = np.where(valid, idx, 0)
        assert np.array_equal(got_idx, exp_idx)           # NULL slots zeroed, like every fixed-width export
        assert arr.null_count == int((~valid).sum())
        exp = [labels[idx[i]].decode() if valid[i] else None for i in range(n)]
        assert arr.to_pylist() == exp
        rb = res.to_record_batch()
        rb.validate(full=True)
        assert rb.column(0).to_pylist() == exp
    del arr, rb  # exported arrays outlive the result: releasing them afterwards must be clean


def test_enum_index_past_the_dictionary_is_an_error(ctx):
    from duckdb_mbt_b200 import typed_result as tr
    from duckdb_mbt_b200.typed_result import DuckDBError
    counts = ch.chunk_counts(3000, "full")
    idx = np.zeros(3000, dtype=np.int64)
    idx[1234] = 9
    col = ch.enum_column("e", [b"a", b"bb", b"a-long-label-in-the-heap"], idx, counts)
    with _result(ctx, ch.ChunkBatch(counts, [col])) as res:
        with pytest.raises(DuckDBError, match="dictionary"):
            tr.text_column(res, 0)
        assert res.raw_column("string", 0) == b""     # getters never fail: empty Bytes (src/duckdb_native.c:2361-2368)


def test_enum_column_without_dictionary_is_rejected(ctx):
    from duckdb_mbt_b200 import arrow_result as ar
    from duckdb_mbt_b200 import native as nat
    counts = ch.chunk_counts(10, "full")
    col = ch.enum_column("e", [b"x"], np.zeros(10, np.int64), counts)
    col.dictionary = None
    with pytest.raises(Exception, match="dictionary"):
        ar.ArrowResult.from_chunks(ctx, ch.ChunkBatch(counts, [col]))
    assert "dictionary" in nat.last_error()


def test_enum_cells_through_connection_query_symbols(ctx):
    """Connection::query's per-cell loop (duckdb_mb_result_is_null / _value, src/duckdb_native.mbt:477-497) and the
    columnar string form give the labels; to_typed keeps them as Value::String (src/duckdb_parsing.mbt:119-122)."""
    from duckdb_mbt_b200 import native as nat
    from duckdb_mbt_b200.query_result import QueryResult, query_per_cell
    batch, labels, idx, valid = _enum_batch(5000, 6, 30, "ragged", 123)
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        L, h = res.lib, res.handle
        assert L.duckdb_mb_result_column_type(h, 0) == ch.T_ENUM
        for r in range(0, 5000, 61):
            assert bool(L.duckdb_mb_result_is_null(h, 0, r)) == ora.cell_is_null(0, r) == (not valid[r])
            assert nat.moonbit_bytes(L.duckdb_mb_result_value(h, 0, r)) == ora.cell_value(0, r)
            if valid[r]:
                assert ora.cell_value(0, r) == labels[idx[r]]
        q = query_per_cell(res)
        c = QueryResult.from_result(res, q.column_types)
        assert c.rows == q.rows and c.nulls == q.nulls
        assert [row[0] for row in q.rows[:50]] == [labels[idx[i]].decode() if valid[i] else "" for i in range(50)]
        t = c.to_typed()
        first = int(np.argmax(valid))
        assert t.get_string(first, 0) == labels[idx[first]].decode()
