"""CPU-only: host-side logic of the package — the MoonBit decoder rules in arrow_result.py against
the oracle's decoder loops, the typed Value surface, the appender model against the reference's
transition table, chunk-batch construction, row-group sharding."""
import ctypes as C

import numpy as np
import pytest

import oracle
from duckdb_mbt_b200 import appender as ap
from duckdb_mbt_b200 import arrow_result as ar
from duckdb_mbt_b200 import chunks as ch
from duckdb_mbt_b200 import shard
from duckdb_mbt_b200 import typed_result as tr

from test_oracle_golden import batch_of


class _Blob(ar.ArrowResult):
    """ArrowResult whose getters return canned blobs (decoder tests need no GPU)."""

    def __init__(self, blobs):
        self.blobs = blobs

    def raw_column(self, kind, col, nullable=False):
        return self.blobs[(kind, nullable)]


def _oracle_blobs(batch, col):
    ora = oracle.OracleResult(batch)
    return {(k, n): ora.get_column(k, col, n) for k in ("int32", "int64", "double", "bool", "string") for n in (False, True)}


def test_fixed_decoders_follow_the_moonbit_rules():
    rng = np.random.default_rng(1)
    n = 5000
    counts = ch.chunk_counts(n, "ragged", rng)
    valid = rng.random(n) > 0.3
    batch = ch.ChunkBatch(counts, [ch.fixed_column("x", ch.T_BIGINT, rng.integers(-2**62, 2**62, n, dtype=np.int64), counts, valid=valid)])
    blobs = _oracle_blobs(batch, 0)
    r = _Blob(blobs)
    v, ok = r.get_column_int32_nullable(0)
    ev, eok = oracle.decode_int32(blobs[("int32", True)], True)
    assert np.array_equal(v, ev) and np.array_equal(ok, eok)
    assert np.array_equal(r.get_column_int32(0), oracle.decode_int32(blobs[("int32", False)])[0])
    # int64 column read into a 32-bit MoonBit Int keeps the low half (src/duckdb_arrow_native.mbt:494-501)
    v64, ok64 = r.get_column_int64_nullable(0)
    e64, eok64 = oracle.decode_int64_as_int(blobs[("int64", True)], True)
    assert np.array_equal(v64, e64) and np.array_equal(ok64, eok64)
    d, okd = r.get_column_double_nullable(0)
    ed, eokd = oracle.decode_double(blobs[("double", True)], True)
    assert np.array_equal(d, ed) and np.array_equal(okd, eokd)
    b, okb = r.get_column_bool_nullable(0)
    eb, eokb = oracle.decode_bool(blobs[("bool", True)], True)
    assert np.array_equal(b, eb) and np.array_equal(okb, eokb)


def test_string_decoder_follows_the_moonbit_rules():
    strings = [b"a", None, b"c", None, b"e", b"", b"hello", b"exactly12byt", b"thirteen byte", "héllo ✓".encode(), None]
    counts = ch.chunk_counts(len(strings))
    batch = ch.ChunkBatch(counts, [ch.string_column("s", strings, counts)])
    blobs = _oracle_blobs(batch, 0)
    r = _Blob(blobs)
    got, ok = r.get_column_string_nullable(0)
    exp, eok = oracle.decode_string(blobs[("string", True)], True)
    # the reference drops NULL rows' terminators from `total`, so the tail of the stream is overwritten by the
    # validity bytes (src/duckdb_native.c:2719-2755); decode_lossy then sees a cut multi-byte sequence
    assert got == [e.decode("utf-8", errors="replace") for e in exp] and np.array_equal(ok, eok)
    got_plain = r.get_column_string(0)
    exp_plain, _ = oracle.decode_string(blobs[("string", False)], False)
    assert got_plain == [e.decode("utf-8", errors="replace") for e in exp_plain]


def test_decoders_reject_short_and_oversized_blobs():
    r = _Blob({("int32", False): b"", ("int32", True): b"\x01\x00", ("string", False): b"\x01\x00\x00\x00",
               ("string", True): b"", ("int64", False): b"", ("int64", True): b"", ("double", False): b"", ("double", True): b"",
               ("bool", False): b"", ("bool", True): b""})
    assert r.get_column_int32(0).shape == (0,)
    assert r.get_column_int32_nullable(0)[0].shape == (0,)
    assert r.get_column_string(0) == []
    assert r.get_column_string_nullable(0)[0] == []
    # count > 1,000,000 -> [] (src/duckdb_arrow_native.mbt:435)
    big = np.int32(1_000_001).tobytes() + b"\0" * (4 * 1_000_001)
    assert _Blob({("int32", False): big}).get_column_int32(0).shape == (0,)
    assert oracle.decode_int32(big)[0].shape == (0,)
    # truncated payload -> []
    short = np.int32(10).tobytes() + b"\0" * 39
    assert _Blob({("int32", False): short}).get_column_int32(0).shape == (0,)


def test_host_batch_pointers_address_the_chunk_vectors():
    rng = np.random.default_rng(3)
    n = 7000
    counts = ch.chunk_counts(n, "ragged", rng)
    valid = rng.random(n) > 0.5
    col = ch.fixed_column("x", ch.T_INTEGER, np.arange(n, dtype=np.int32), counts, valid=valid)
    s = ch.string_column_bulk("s", rng.integers(0, 30, n), None, counts, rng)
    hb = ar.HostBatch(ch.ChunkBatch(counts, [col, s]))
    st = hb.struct
    assert st.ncols == 2 and st.nchunks == counts.shape[0] and st.flags == 0
    c0 = st.cols[0]
    ro = np.concatenate([[0], np.cumsum(counts)])
    for k in (0, 1, counts.shape[0] - 1):
        if counts[k]:
            first = C.c_int32.from_address(c0.data[k]).value
            assert first == ro[k]
        if col.val_off[k] >= 0:
            assert c0.validity[k] == col.validity.ctypes.data + 8 * int(col.val_off[k])
        else:
            assert not c0.validity[k]
    c1 = st.cols[1]
    assert not c1.validity and c1.heap_base == s.heap.ctypes.data and c1.heap_len == s.heap.shape[0]
    assert ar.HostBatch(ch.ChunkBatch(counts, [col]), pinned=True).struct.flags == 1


def test_typed_value_surface():
    col = tr.TypedColumn(tr.INT, np.asarray([1, 2, 3], dtype=np.int32), np.asarray([True, False, True]))
    scol = tr.TypedColumn(tr.STRING, None, np.asarray([True, True, False]), np.asarray([0, 2, 2, 2], dtype=np.int32), b"hi")
    t = tr.TypedQueryResult(["a", "s"], [col, scol])
    assert t.row_count() == 3 and t.column_count() == 2
    assert t.get_value(0, 0) == tr.Value(tr.INT, 1) and t.get_value(1, 0).is_null()
    assert t.get_value(3, 0) is None and t.get_value(0, 2) is None
    assert t.get_int(0, 0) == 1 and t.get_int(1, 0) is None and t.get_double(0, 0) is None
    assert t.get_string(0, 1) == "hi" and t.get_string(1, 1) == "" and t.get_string(2, 1) is None
    assert t.is_null(1, 0) and not t.is_null(0, 0) and t.is_null(99, 0)
    assert t.get_int_column(0) == [1, None, 3]
    assert t.get_string_column(0) == [None, None, None]
    assert t.get_column(5) is None and t.get_int_column(-1) is None
    assert repr(t.get_column(1)[0]) == "String('hi')"
    assert tr.Value(tr.DATE, 5).as_date() == 5 and tr.Value(tr.DATE, 5).as_int() is None


def test_appender_model_matches_the_reference_transition_table():
    # src/duckdb_appender_state_machine.mbt:54-239
    m = ap.AppenderModel(2)
    assert m.state == ap.NOT_CREATED
    assert m.execute("begin_row").state == ap.ERROR  # NotCreated can only Create
    m = ap.AppenderModel(2).execute("create")
    assert m.state == ap.READY
    m.execute("begin_row")
    assert m.state == ap.ROW_IN_PROGRESS
    m.execute("append_int").execute("append_double")
    assert m.column_count == 2 and m.state == ap.ROW_IN_PROGRESS
    m.execute("end_row")
    assert m.state == ap.READY and m.row_count == 1 and m.flushed_row_count == 0
    m.execute("flush")
    assert m.state == ap.FLUSHED and m.flushed_row_count == 1
    m.execute("flush")
    assert m.state == ap.FLUSHED
    m.execute("begin_row").execute("append_int").execute("end_row")  # under-filled row
    assert m.state == ap.ERROR
    assert m.execute("flush").state == ap.ERROR  # sticky
    assert m.execute("close").state == ap.CLOSED
    assert m.execute("begin_row").state == ap.CLOSED  # terminal
    over = ap.AppenderModel(1).execute("create").execute("begin_row").execute("append_int").execute("append_int")
    assert over.state == ap.ERROR  # over-filled row
    assert ap.AppenderModel(1).execute("create").execute("append_int").state == ap.ERROR  # append outside a row
    assert ap.AppenderModel(1).execute("create").execute("begin_row").execute("flush").state == ap.ERROR
    assert ap.AppenderModel(1).execute("create").execute("begin_row").execute("close").state == ap.CLOSED


def test_shard_chunks_cover_the_batch_without_overlap():
    for nchunks in (0, 1, 7, 8, 9, 1000, 29297):
        for world in (1, 2, 4, 8):
            ranges = [shard.shard_chunks(nchunks, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nchunks
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            per = -(-nchunks // world) if nchunks else 0
            assert all(c1 - c0 <= per for c0, c1 in ranges)
    assert shard.string_bases([5, 0, 7]) == [0, 5, 5]
    with pytest.raises(ValueError):
        shard.shard_chunks(10, 2, 2)


def test_sharded_string_column_stitches_to_the_whole():
    b = ch.config_c3(30_000, pattern="ragged", seed=5)
    whole_o, whole_d = oracle.OracleResult(b).arrow_string(0, 1)
    parts = []
    for r in range(4):
        c0, c1 = shard.shard_chunks(b.nchunks, 4, r)
        sub = shard.slice_batch(b, c0, c1)
        o, d = oracle.OracleResult(sub).arrow_string(0, 0)
        assert o[0] == 0  # every shard is an independent record batch
        parts.append((o, d.tobytes()))
    offsets, data = shard.concat_utf8(parts)
    assert np.array_equal(offsets, whole_o) and data == whole_d.tobytes()


def test_reference_golden_vectors_decode_through_the_mirror():
    # src/duckdb_arrow_test.mbt:343-371 through oracle blobs + the package's decoders
    b = batch_of(("x", ch.T_INTEGER, [1, None, 3, None, 5]))
    r = _Blob(_oracle_blobs(b, 0))
    v, ok = r.get_column_int32_nullable(0)
    assert v.tolist() == [1, 0, 3, 0, 5] and ok.tolist() == [True, False, True, False, True]
