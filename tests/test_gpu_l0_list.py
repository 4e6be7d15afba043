"""L0 parity of K9 (LIST vectors -> Arrow list<child>, kernels_list.cu) against the oracle's restatement
(oracle.c ora_list_arrow, itself pinned on pyarrow in tests/test_oracle_golden.py): offsets, gathered child values
(NULL elements zeroed), child bitmap, totals -- bit for bit, for contiguous, shuffled and shared entry layouts."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from duckdb_mbt_b200 import chunks as ch  # noqa: E402
from duckdb_mbt_b200 import native as nat  # noqa: E402

import list_cases  # noqa: E402


def _dev(a: np.ndarray, pad: int = 64):
    t = torch.zeros(a.nbytes + pad, dtype=torch.uint8, device="cuda:0")
    if a.nbytes:
        t[: a.nbytes] = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to("cuda:0")
    return t


def _run(lc, large, with_sizes=False):
    L = nat.lib()
    n = int(lc.counts.sum())
    nch = lc.counts.shape[0]
    vecs = np.zeros(nch, dtype=[("data_off", "<u8"), ("val_off", "<i8")])
    vecs["data_off"], vecs["val_off"] = lc.data_off, lc.val_off
    row_off = np.zeros(nch + 1, dtype=np.int64)
    np.cumsum(lc.counts, out=row_off[1:])
    d = {k: _dev(v) for k, v in dict(entries=lc.entries, validity=lc.validity if lc.validity is not None else np.zeros(1, np.uint64),
                                     vecs=vecs, child_base=lc.child_base, child_data=lc.child_data, child_validity=lc.child_validity,
                                     child_val_off=lc.child_val_off, counts=lc.counts.astype(np.uint32), row_off=row_off).items()}
    cap = lc.capacity
    ow = 8 if large else 4
    out_off = torch.full(((n + 1) * ow + 64,), 0xAB, dtype=torch.uint8, device="cuda:0")
    out_child = torch.full((max(cap, 1) * lc.width + 64,), 0xAB, dtype=torch.uint8, device="cuda:0")
    out_bm = torch.full((((cap + 63) // 64 + 1) * 8,), 0xAB, dtype=torch.uint8, device="cuda:0")
    ctr = torch.zeros(2, dtype=torch.int64, device="cuda:0")
    scratch = torch.zeros(L.dmb_dev_list_scratch_bytes(nch) + 64, dtype=torch.uint8, device="cuda:0")
    job = nat.ListJob(d["entries"].data_ptr(), d["validity"].data_ptr(), d["vecs"].data_ptr(), d["child_base"].data_ptr(),
                      d["child_data"].data_ptr(), d["child_validity"].data_ptr(), d["child_val_off"].data_ptr(),
                      out_off.data_ptr(), out_child.data_ptr(), out_bm.data_ptr(), ctr.data_ptr(), ctr.data_ptr() + 8,
                      lc.width, 1 if large else 0)
    if with_sizes:
        d["child_sizes"] = _dev(lc.child_sizes)
        job.child_sizes = d["child_sizes"].data_ptr()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(L.dmb_dev_list_batch(C.byref(job), d["counts"].data_ptr(), d["row_off"].data_ptr(), nch, n, cap, scratch.data_ptr(), st), "list")
    torch.cuda.synchronize()
    flags = int(scratch[:8].cpu().numpy().view(np.uint64)[0])
    offs = out_off.cpu().numpy()[: (n + 1) * ow].view(np.int64 if large else np.int32)
    total, nulls = [int(x) for x in ctr.cpu().numpy()]
    return offs, out_child.cpu().numpy()[: total * lc.width], out_bm.cpu().numpy(), total, nulls, flags


@pytest.mark.parametrize("layout", ["contiguous", "shuffled", "shared"])
@pytest.mark.parametrize("n,width,pattern,large", [(1, 4, "full", False), (2048, 1, "full", False), (2049, 8, "full", True),
                                                   (10_000, 4, "ragged", False), (60_001, 2, "ragged", False), (150_000, 16, "full", True)])
def test_list_to_arrow_matches_the_oracle(layout, n, width, pattern, large):
    lc = list_cases.make_list_column(n, width, pattern, 500 + n + width, layout)
    exp_off, exp_child, exp_bm, exp_total, exp_nulls = oracle.list_arrow(
        lc.entries, lc.data_off, lc.validity, lc.val_off, lc.counts, lc.child_base, lc.child_data, lc.child_validity, lc.child_val_off,
        width, large, lc.capacity)
    offs, child, bm, total, nulls, flags = _run(lc, large)
    assert flags == 0
    assert total == exp_total == lc.capacity and nulls == exp_nulls
    assert np.array_equal(offs, exp_off)
    assert child.tobytes() == exp_child.tobytes()
    nb = (total + 7) // 8
    assert bm[:nb].tobytes() == exp_bm[:nb].tobytes()


def test_list_without_nulls_and_empty_lists_only():
    lc = list_cases.make_list_column(30_000, 4, "ragged", 9, "contiguous", null_frac=0.0, child_null_frac=0.0)
    exp = oracle.list_arrow(lc.entries, lc.data_off, lc.validity, lc.val_off, lc.counts, lc.child_base, lc.child_data, lc.child_validity,
                            lc.child_val_off, 4, False, lc.capacity)
    offs, child, bm, total, nulls, flags = _run(lc, False)
    assert flags == 0 and nulls == 0 and total == exp[3]
    assert np.array_equal(offs, exp[0]) and child.tobytes() == exp[1].tobytes()
    assert bm[: (total + 7) // 8].tobytes() == exp[2][: (total + 7) // 8].tobytes()
    lc0 = list_cases.make_list_column(5000, 8, "full", 10, "contiguous", max_len=0)
    offs, child, bm, total, nulls, flags = _run(lc0, False)
    assert flags == 0 and total == 0 and not offs.any()


def test_list_arrow_array_reads_back_in_pyarrow():
    pa = pytest.importorskip("pyarrow")
    lc = list_cases.make_list_column(7000, 4, "ragged", 77, "shuffled")
    offs, child, bm, total, nulls, flags = _run(lc, False)
    values = pa.Array.from_buffers(pa.binary(4), total, [pa.py_buffer(bm.tobytes()), pa.py_buffer(child.tobytes() + b"\0")], null_count=nulls)
    pbits = np.packbits(lc.valid.astype(np.uint8), bitorder="little").tobytes() + b"\0"
    arr = pa.Array.from_buffers(pa.list_(pa.binary(4)), len(lc.valid), [pa.py_buffer(pbits), pa.py_buffer(offs.tobytes())], children=[values])
    arr.validate(full=True)
    assert arr.to_pylist() == lc.expected


@pytest.mark.parametrize("layout", ["contiguous", "shuffled"])
@pytest.mark.parametrize("n,child_type,pattern", [(1, ch.T_INTEGER, "full"), (5000, ch.T_INTEGER, "ragged"), (40_001, ch.T_BIGINT, "ragged"),
                                                  (9000, ch.T_SMALLINT, "full"), (3000, ch.T_UUID, "full")])
def test_list_column_through_the_host_api(layout, n, child_type, pattern):
    """dmb_host_column + dmb_host_list -> duckdb_mb_gpu_result_export_arrow: a nested Arrow list<child> array (child vectors
    of different sizes gathered by the stager), validated in full by pyarrow and equal to the lists the chunks describe."""
    pa = pytest.importorskip("pyarrow")
    from duckdb_mbt_b200 import arrow_result as ar
    w = ch.PHYS_WIDTH[ch.phys_of_type(child_type)]
    lc = list_cases.make_list_column(n, w, pattern, 900 + n, layout)
    other = ch.fixed_column("i", ch.T_INTEGER, np.arange(n, dtype=np.int32), lc.counts)
    batch = ch.ChunkBatch(lc.counts, [list_cases.as_column(lc, "l", child_type), other])
    exp = [None if row is None else [None if v is None else int.from_bytes(v, "little", signed=True) for v in row] for row in lc.expected]
    with ar.GpuContext(0) as ctx, ar.ArrowResult.from_chunks(ctx, batch) as res:
        arr = res.to_arrow(0)
        arr.validate(full=True)
        assert pa.types.is_list(arr.type)
        assert arr.null_count == int((~lc.valid).sum())
        got = arr.to_pylist()
        if child_type == ch.T_UUID:  # fixed_size_binary(16), bytes as stored
            exp = lc.expected
        assert got == exp
        rb = res.to_record_batch()
        rb.validate(full=True)
        assert rb.column(1).to_pylist() == list(range(n))
        assert rb.column(0).null_count == arr.null_count
    del arr, rb


def test_list_children_that_need_a_conversion():
    """BOOLEAN children (Arrow bit-packed) and DECIMAL children (decimal128): the gathered dense child takes a second pass
    through the fixed-width conversion kernel"""
    import decimal
    pa = pytest.importorskip("pyarrow")
    from duckdb_mbt_b200 import arrow_result as ar
    with ar.GpuContext(0) as ctx:
        lc = list_cases.make_list_column(20_000, 4, "ragged", 41, "shuffled")
        batch = ch.ChunkBatch(lc.counts, [list_cases.as_column(lc, "l", ch.T_DECIMAL, 9, 2)])
        with ar.ArrowResult.from_chunks(ctx, batch) as res:
            arr = res.to_arrow(0)
            assert arr.type == pa.list_(pa.field("item", pa.decimal128(9, 2)))
            # validate(full=True) would reject random 32-bit payloads beyond 9 digits: the values are checked directly
            arr.validate()
            exp = [None if row is None else [None if v is None else decimal.Decimal(int.from_bytes(v, "little", signed=True)).scaleb(-2) for v in row] for row in lc.expected]
            assert arr.to_pylist() == exp
        lc = list_cases.make_list_column(30_001, 1, "ragged", 42, "contiguous")
        lc.child_data &= 1  # DuckDB bool vectors hold 0 / 1
        lc.expected = [None if row is None else [None if v is None else bool(v[0] & 1) for v in row] for row in lc.expected]
        batch = ch.ChunkBatch(lc.counts, [list_cases.as_column(lc, "l", ch.T_BOOLEAN)])
        with ar.ArrowResult.from_chunks(ctx, batch) as res:
            arr = res.to_arrow(0)
            arr.validate(full=True)
            assert arr.type == pa.list_(pa.field("item", pa.bool_()))
            assert arr.to_pylist() == lc.expected


def test_list_of_varchar_is_not_exported_yet():
    from duckdb_mbt_b200 import arrow_result as ar
    from duckdb_mbt_b200 import native as nat
    lc = list_cases.make_list_column(100, 16, "full", 3, "contiguous")
    col = list_cases.as_column(lc, "l", ch.T_VARCHAR)
    with ar.GpuContext(0) as ctx:
        with pytest.raises(Exception, match="fixed-width"):
            ar.ArrowResult.from_chunks(ctx, ch.ChunkBatch(lc.counts, [col]))
        assert "LIST child" in nat.last_error()


def test_list_and_enum_edge_cases_through_the_host_api():
    """no chunks at all, all-NULL lists, only empty lists, an empty ENUM dictionary behind all-NULL rows"""
    pa = pytest.importorskip("pyarrow")
    from duckdb_mbt_b200 import arrow_result as ar
    with ar.GpuContext(0) as ctx:
        # all rows NULL: no child element is ever read (the entries are garbage)
        lc = list_cases.make_list_column(3000, 4, "ragged", 5, "contiguous", null_frac=1.0)
        with ar.ArrowResult.from_chunks(ctx, ch.ChunkBatch(lc.counts, [list_cases.as_column(lc, "l", ch.T_INTEGER)])) as res:
            arr = res.to_arrow(0)
            arr.validate(full=True)
            assert arr.null_count == 3000 and len(arr.values) == 0
        # only empty lists
        lc = list_cases.make_list_column(2500, 8, "full", 6, "contiguous", null_frac=0.0, max_len=0)
        with ar.ArrowResult.from_chunks(ctx, ch.ChunkBatch(lc.counts, [list_cases.as_column(lc, "l", ch.T_BIGINT)])) as res:
            arr = res.to_arrow(0)
            arr.validate(full=True)
            assert arr.to_pylist() == [[]] * 2500
        # a result without chunks
        lc = list_cases.make_list_column(0, 4, "full", 7, "contiguous")
        en = ch.enum_column("e", [b"a", b"bb"], np.zeros(0, np.int64), lc.counts)
        with ar.ArrowResult.from_chunks(ctx, ch.ChunkBatch(lc.counts, [list_cases.as_column(lc, "l", ch.T_INTEGER), en])) as res:
            assert res.row_count() == 0
            rb = res.to_record_batch()
            rb.validate(full=True)
            assert rb.num_rows == 0 and pa.types.is_list(rb.column(0).type) and pa.types.is_dictionary(rb.column(1).type)
            assert res.get_column_string(1) == []
        # ENUM: all rows NULL over an empty dictionary
        counts = ch.chunk_counts(100, "full")
        en = ch.enum_column("e", [], np.zeros(100, np.int64), counts, valid=np.zeros(100, bool))
        with ar.ArrowResult.from_chunks(ctx, ch.ChunkBatch(counts, [en])) as res:
            arr = res.to_arrow(0)
            arr.validate(full=True)
            assert arr.null_count == 100 and len(arr.dictionary) == 0
            strs, valid = res.get_column_string_nullable(0)
            assert not any(valid)


def test_list_entry_outside_its_child_vector_is_flagged_not_followed():
    """round-1 advice: a malformed / stale entry (offset + length past duckdb_list_vector_get_size) must raise an error,
    never read outside the staged child slab; host API: an error, not a result"""
    from duckdb_mbt_b200 import arrow_result as ar
    lc = list_cases.make_list_column(6000, 4, "full", 31, "contiguous", null_frac=0.0)
    offs, child, bm, total, nulls, flags = _run(lc, False, with_sizes=True)
    assert flags == 0 and total == lc.capacity
    ent = lc.entries.view(np.uint64).reshape(-1, 2)
    ent[2048 + 5] = (int(lc.child_sizes[1]) - 1, 7)  # chunk 1, row 5: reaches 6 elements past its child vector
    offs, child, bm, total, nulls, flags = _run(lc, False, with_sizes=True)
    assert flags & 8
    batch = ch.ChunkBatch(lc.counts, [list_cases.as_column(lc, "l", ch.T_INTEGER)])
    with ar.GpuContext(0) as ctx, ar.ArrowResult.from_chunks(ctx, batch) as res:
        with pytest.raises(ar.DuckDBError, match="outside its chunk's child vector"):
            res.to_arrow(0)
