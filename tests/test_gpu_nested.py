"""Nested types through the host API (round-1 verdict item 8): STRUCT, LIST<VARCHAR>,
MAP = LIST<STRUCT<key, value>>, LIST<STRUCT>, LIST<LIST<x>>.  The reference rejects them on its chunk path
(src/duckdb_native.c:271-303) and has no Arrow mapping, so the contract is the Arrow format: every exported array is
validated in full by pyarrow and must read back as exactly the Python values the DuckDB-shaped vectors stand for."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
pa = pytest.importorskip("pyarrow")

from duckdb_mbt_b200 import chunks as ch  # noqa: E402

import nested_cases  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from duckdb_mbt_b200 import arrow_result as ar
    c = ar.GpuContext(0)
    yield c
    c.close()


def _export(ctx, counts, cols):
    from duckdb_mbt_b200 import arrow_result as ar
    batch = ch.ChunkBatch(counts, cols)
    with ar.ArrowResult.from_chunks(ctx, batch) as res:
        arrays = res.to_arrow()
    for a in arrays:
        a.validate(full=True)
    return arrays


@pytest.mark.parametrize("kind,arrow_type", [
    ("varchar", pa.list_(pa.string())),
    ("map", pa.map_(pa.string(), pa.int32())),
    ("struct", pa.list_(pa.struct([("a", pa.int32()), ("s", pa.string())]))),
    ("list", pa.list_(pa.list_(pa.int32()))),
    ("list_varchar", pa.list_(pa.list_(pa.string()))),
])
@pytest.mark.parametrize("n,pattern,layout", [(1, "full", "contiguous"), (2049, "full", "contiguous"), (9000, "ragged", "shuffled"),
                                              (40_001, "ragged", "contiguous")])
def test_list_children_that_are_not_one_fixed_vector(ctx, kind, arrow_type, n, pattern, layout):
    counts, col, expected = nested_cases.make_list_of(kind, n, pattern, 300 + n, layout)
    other = ch.fixed_column("i", ch.T_INTEGER, np.arange(n, dtype=np.int32), counts)
    arr, i = _export(ctx, counts, [col, other])
    assert arr.type.equals(arrow_type, check_metadata=False) or str(arr.type).replace("not null", "").replace(" ", "") == str(arrow_type).replace(" ", ""), arr.type
    assert len(arr) == n and arr.null_count == sum(1 for e in expected if e is None)
    assert arr.to_pylist() == expected
    assert i.to_pylist() == list(range(n))


@pytest.mark.parametrize("n,pattern", [(1, "full"), (5000, "ragged"), (30_011, "full")])
def test_struct_columns(ctx, n, pattern):
    counts, col, expected = nested_cases.make_struct(n, pattern, 40 + n)
    (arr,) = _export(ctx, counts, [col])
    assert pa.types.is_struct(arr.type) and [f.name for f in arr.type] == ["i", "s", "d", "inner"]
    assert arr.type.field("d").type == pa.decimal128(9, 2) and pa.types.is_struct(arr.type.field("inner").type)
    assert arr.null_count == sum(1 for e in expected if e is None)
    assert arr.to_pylist() == expected


def test_nested_columns_in_a_sharded_stream(ctx):
    """the slices of a nested column (shards / stream batches) carry their child levels with them"""
    from duckdb_mbt_b200 import arrow_result as ar
    n = 20_000
    counts, col, expected = nested_cases.make_list_of("map", n, "full", 77)
    counts2, st, exp_st = nested_cases.make_struct(n, "full", 78)
    batch = ch.ChunkBatch(counts, [col, st])
    with ar.ArrowResult.from_chunks(ctx, batch) as res:
        reader = res.to_stream(max_batch_rows=6000)
        got_m, got_s = [], []
        for b in reader:
            b.validate(full=True)
            got_m += b.column(0).to_pylist()
            got_s += b.column(1).to_pylist()
    assert got_m == expected and got_s == exp_st


def test_malformed_nested_entries_are_refused(ctx):
    from duckdb_mbt_b200 import arrow_result as ar
    counts, col, expected = nested_cases.make_list_of("varchar", 3000, "full", 5, null_frac=0.0)
    ent = col.data.view(np.uint64).reshape(-1, 2)
    ent[7] = (int(col.list_child_sizes[0]) - 1, 9)
    with ar.ArrowResult.from_chunks(ctx, ch.ChunkBatch(counts, [col])) as res:
        with pytest.raises(ar.DuckDBError, match="outside its chunk's child vector"):
            res.to_arrow(0)
