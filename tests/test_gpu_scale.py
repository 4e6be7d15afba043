"""Parity at the sizes bench.py runs (round-1 verdict, item 1): the persistent grids, the look-back across
tens of thousands of tiles and the > 2 GiB outputs are compared with the CPU oracle here, not only timed.

The oracle converts WINDOWS of whole chunks (about a million rows each: the first chunks, the ragged tail and
seeded random ranges); the device converts the whole column and the window is cut out of its output (utf8
offsets rebased by the offset of the window's first row).  The size-independent properties are checked over
the whole output: monotonic offsets, last offset == data length, null counts.

  C2  lineitem shape, 12 M rows per column        L0 (the DeviceStep bench.py times) and L1 / L2 (host API)
  C3  100 M VARCHAR rows, > 2^31 string bytes     L0 (int32 overflow flag, int64 result), L1 with a registered
                                                  heap (`surely_large`) and with scattered pointers (the
                                                  utf8 -> large_utf8 retry, src/duckdb_native.c:2488 is the
                                                  reference's own int32 overflow at this size)
  C4  all 64 columns, 10 M rows, 30 % NULL        L0; L1 at 2 M rows
  C5  reverse, 50 M rows, bit offset 3            L0 kernels and the appender host API
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
pa = pytest.importorskip("pyarrow")

import oracle  # noqa: E402
from duckdb_mbt_b200 import chunks as ch  # noqa: E402
from duckdb_mbt_b200 import shard  # noqa: E402

WINDOW_CHUNKS = 489  # ~1 M rows


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(scope="module")
def ctx():
    _need_gpu()
    from duckdb_mbt_b200 import arrow_result as ar
    c = ar.GpuContext(0)
    yield c
    c.close()


def windows(nchunks: int, seed: int, k: int = 2, size: int = WINDOW_CHUNKS):
    """chunk ranges [c0, c1): the head, the tail (ragged last chunk) and k seeded random ones"""
    size = min(size, nchunks)
    out = [(0, size), (nchunks - size, nchunks)]
    rng = np.random.default_rng(seed)
    for _ in range(k):
        c0 = int(rng.integers(0, nchunks - size + 1))
        out.append((c0, c0 + size))
    return out


def _np(t, dtype=np.uint8):
    a = t.cpu().numpy()
    return a[: a.shape[0] // np.dtype(dtype).itemsize * np.dtype(dtype).itemsize].view(dtype)


def check_fixed_window(ora, col, dst, width, values_u8, bitmap_u8, row0, nrows_w):
    """device outputs of a whole column (numpy uint8 views) against the oracle on a chunk window"""
    ev, bm, _, nc = ora.arrow_fixed(col, dst, width)
    if dst == ch.D_BOOL_BITS:
        got = values_u8[row0 // 8: row0 // 8 + ev.shape[0]]
        tail_bits = nrows_w % 8
        if tail_bits:  # the window's last byte is shared with the next chunk's rows on the device side
            got = got.copy()
            got[-1] &= (1 << tail_bits) - 1
    else:
        got = values_u8[row0 * width: row0 * width + ev.shape[0]]
    assert np.array_equal(got, ev), f"values differ col={col} dst={dst} window row0={row0}"
    nb = (nrows_w + 7) // 8
    gb = bitmap_u8[row0 // 8: row0 // 8 + nb].copy()
    eb = bm[:nb].copy()
    if nrows_w % 8:
        gb[-1] &= (1 << (nrows_w % 8)) - 1
        eb[-1] &= (1 << (nrows_w % 8)) - 1
    assert np.array_equal(gb, eb), f"bitmap differs col={col} window row0={row0}"
    return nc


def check_string_window(ora, col, offsets, data_u8, row0, nrows_w):
    """offsets: the device's whole offsets array (int32 or int64 numpy), data_u8: its data buffer"""
    eo, ed = ora.arrow_string(col, 1)
    base = int(offsets[row0])
    got_o = offsets[row0: row0 + nrows_w + 1].astype(np.int64) - base
    assert np.array_equal(got_o, eo), f"offsets differ col={col} window row0={row0}"
    assert np.array_equal(data_u8[base: base + ed.shape[0]], ed), f"utf8 data differs col={col} window row0={row0}"


# ------------------------------------------------------------------------------------------- C2
C2_ROWS = 12_000_000 + 1234  # ragged last chunk


@pytest.fixture(scope="module")
def c2():
    """the C2 table bench.py builds (same generator, same DeviceStep), at 12 M rows, + its host copy"""
    _need_gpu()
    import bench
    heaps = []

    def host_heap_alloc(nb):
        a = np.zeros(max(int(nb), 1), dtype=np.uint8)
        heaps.append(a)
        return a

    db = bench.build_c2_device(C2_ROWS, 20260103, torch.device("cuda", 0), host_heap_alloc=host_heap_alloc)
    hb = db.to_host_batch()
    yield db, hb
    del db, hb, heaps
    torch.cuda.empty_cache()


def test_c2_bench_step_l0_parity(c2):
    import bench
    db, hb = c2
    step = bench.DeviceStep(db)
    step.run()
    step.run()  # a second pass over the same scratch / outputs, like the timed loop
    torch.cuda.synchronize()
    step.check()
    checked = bench.parity_check_step(step, seed=11, k=2)  # every output column, 4 windows of ~1 M rows
    assert checked >= 3 * WINDOW_CHUNKS * 2048 * len(hb.columns)
    # whole-column properties
    n = db.nrows
    for so in step.strings:
        offs = _np(so.offsets, np.int32)[: n + 1]
        assert offs[0] == 0 and np.all(np.diff(offs) >= 0)
        assert int(offs[-1]) == int(_np(so.total, np.uint64)[0]) == db.meta[so.col]["total_len"]


def test_c2_host_api_l1_l2_parity(ctx, c2):
    from duckdb_mbt_b200 import arrow_result as ar
    db, hb = c2
    n = hb.nrows
    with ar.ArrowResult.from_chunks(ctx, hb) as res:
        arrays = res.to_arrow()
        assert len(arrays) == 16 and all(len(a) == n for a in arrays)
        assert all(a.null_count == 0 for a in arrays)
        blobs = {
            "int32": (0, res.raw_column("int32", 0, True)),
            "int64": (5, res.raw_column("int64", 5, True)),
            "string": (15, res.raw_column("string", 15, True)),
        }
        for c0, c1 in windows(hb.nchunks, seed=5, k=1):
            sub = shard.slice_batch(hb, c0, c1)
            ora = oracle.OracleResult(sub)
            row0, nw = c0 * 2048, sub.nrows
            for j, col in enumerate(hb.columns):
                a = arrays[j]
                bufs = a.buffers()
                if col.phys == ch.P_STRING:
                    offs = np.frombuffer(bufs[1], dtype=np.int32)[: n + 1]
                    check_string_window(ora, j, offs, np.frombuffer(bufs[2], dtype=np.uint8), row0, nw)
                else:
                    dst, w = (ch.D_I128, 16) if col.type_id == ch.T_DECIMAL else (ch.D_SAME, col.width)
                    check_fixed_window(ora, j, dst, w, np.frombuffer(bufs[1], dtype=np.uint8),
                                       np.frombuffer(bufs[0], dtype=np.uint8), row0, nw)
            # L2 getters: [n:i32][values][validity bytes] of the whole column; the window's part of each section
            for kind, (j, blob) in blobs.items():
                exp = ora.get_column(kind, j, True)
                assert blob[:4] == np.int32(n).tobytes()
                if kind == "string":
                    offs = np.frombuffer(arrays[j].buffers()[1], dtype=np.int32)
                    # no NULLs and no embedded NULs in C2: stream position of a row = utf8 offset + one terminator per earlier row
                    p0, p1 = int(offs[row0]) + row0, int(offs[row0 + nw]) + row0 + nw
                    total = int(np.frombuffer(blob[4:8], dtype=np.int32)[0])
                    assert total == int(offs[n]) + n
                    assert blob[8 + p0: 8 + p1] == exp[8: 8 + (p1 - p0)]
                    assert blob[8 + total + row0: 8 + total + row0 + nw] == exp[8 + (p1 - p0):]
                else:
                    w = 4 if kind == "int32" else 8
                    assert blob[4 + row0 * w: 4 + (row0 + nw) * w] == exp[4: 4 + nw * w], kind
                    assert blob[4 + n * w + row0: 4 + n * w + row0 + nw] == exp[4 + nw * w:], kind
            ora.close()


# ------------------------------------------------------------------------------------------- C3
C3_ROWS = 100_000_000


@pytest.fixture(scope="module")
def c3():
    _need_gpu()
    from duckdb_mbt_b200 import devgen
    heaps = []

    def host_heap_alloc(nb):
        a = np.zeros(max(int(nb), 1), dtype=np.uint8)
        heaps.append(a)
        return a

    gen = torch.Generator(device="cuda:0")
    gen.manual_seed(20260104)
    db = devgen.GeneratedBatch(C3_ROWS, "cuda:0")
    db.add_string(gen, 0.10, 0, 64, name="s", host_heap_alloc=host_heap_alloc)
    assert db.meta[0]["total_len"] > 2**31, "C3's defining property: more string bytes than int32 offsets hold"
    hb = db.to_host_batch()
    yield db, hb
    del db, hb, heaps
    torch.cuda.empty_cache()


def test_c3_l0_int32_overflow_flag_and_int64_result(c3):
    db, hb = c3
    n = db.nrows
    total = db.meta[0]["total_len"]
    so32 = db.plan_string(0, 0, data_capacity=total)
    db.run_string(so32)
    assert db.string_error(so32) & 2, "int32 offsets must raise the overflow flag (the reference overflows total_size here)"
    del so32
    so = db.plan_string(0, 1, data_capacity=total)
    db.run_string(so)
    assert db.string_error(so) == 0
    offs = _np(so.offsets, np.int64)[: n + 1]
    assert offs[0] == 0 and int(offs[-1]) == total == int(_np(so.total, np.uint64)[0])
    assert np.all(np.diff(offs) >= 0)
    data = _np(so.data)
    for c0, c1 in windows(hb.nchunks, seed=3):
        sub = shard.slice_batch(hb, c0, c1)
        ora = oracle.OracleResult(sub)
        check_string_window(ora, 0, offs, data, c0 * 2048, sub.nrows)
        ora.close()


@pytest.mark.parametrize("register_heap", [True, False], ids=["registered_heap_surely_large", "scattered_heap_retry"])
def test_c3_host_api_large_utf8(ctx, c3, register_heap):
    """register_heap: heap_len > 2^31 - 1 picks 64-bit offsets up front; scattered pointers (heap_len = 0, the
    layout the glue hands over) start with int32 offsets, overflow, and are redone as large_utf8"""
    from duckdb_mbt_b200 import arrow_result as ar
    db, hb = c3
    n = hb.nrows
    with ar.ArrowResult.from_chunks(ctx, hb, register_heap=register_heap) as res:
        (a,) = res.to_arrow()
        assert a.type == pa.large_string()
        assert len(a) == n
        bufs = a.buffers()
        offs = np.frombuffer(bufs[1], dtype=np.int64)[: n + 1]
        assert offs[0] == 0 and int(offs[-1]) == db.meta[0]["total_len"]
        valid = db.meta[0]["valid"]
        assert a.null_count == n - int(valid.sum().item())
        data = np.frombuffer(bufs[2], dtype=np.uint8)
        for c0, c1 in windows(hb.nchunks, seed=4, k=1):
            sub = shard.slice_batch(hb, c0, c1)
            ora = oracle.OracleResult(sub)
            check_string_window(ora, 0, offs, data, c0 * 2048, sub.nrows)
            _, bm, _, _ = ora.arrow_fixed(0, ch.D_SAME, 16, want_values=False)
            nb = sub.nrows // 8
            assert np.array_equal(np.frombuffer(bufs[0], dtype=np.uint8)[c0 * 256: c0 * 256 + nb], bm[:nb])
            ora.close()
        del a, bufs, offs, data


# ------------------------------------------------------------------------------------------- C4
def _c4_device(nrows):
    from duckdb_mbt_b200 import devgen
    cols = [(ch.T_TIMESTAMP, 0)] * 22 + [(ch.T_DECIMAL, 18)] * 21 + [(ch.T_HUGEINT, 0)] * 21
    db = devgen.fixed_batch(nrows, cols, null_frac=0.30, seed=20260105)
    for c in db.batch.columns:
        if c.type_id == ch.T_DECIMAL:
            c.dec_scale = 3
    return db


def _c4_dst(col):
    return ch.D_I128 if col.type_id in (ch.T_DECIMAL, ch.T_HUGEINT) else ch.D_SAME


def test_c4_all_64_columns_l0_parity():
    _need_gpu()
    n = 10_000_000 + 77
    db = _c4_device(n)
    assert len(db.batch.columns) == 64
    plan = db.plan_fixed([(j, _c4_dst(c)) for j, c in enumerate(db.batch.columns)], bitmap=True)
    db.run_fixed(plan)
    torch.cuda.synchronize()
    hb = db.to_host_batch()
    wins = windows(hb.nchunks, seed=6, k=1)
    oras = [(c0, shard.slice_batch(hb, c0, c1)) for c0, c1 in wins]
    oras = [(c0, sub, oracle.OracleResult(sub)) for c0, sub in oras]
    for o in plan[0]:
        vals, bm = _np(o.values), _np(o.bitmap)
        for c0, sub, ora in oras:
            check_fixed_window(ora, o.col, o.op & 0xFF, o.width, vals, bm, c0 * 2048, sub.nrows)
        valid = db.meta[o.col]["valid"]
        assert int(_np(o.null_count, np.uint64)[0]) == n - int(valid.sum().item())


def test_c4_all_64_columns_host_api(ctx):
    from duckdb_mbt_b200 import arrow_result as ar
    n = 2_000_000 + 5
    db = _c4_device(n)
    hb = db.to_host_batch()
    ora = oracle.OracleResult(hb)
    with ar.ArrowResult.from_chunks(ctx, hb) as res:
        arrays = res.to_arrow()
        assert len(arrays) == 64
        for j, col in enumerate(hb.columns):
            a = arrays[j]
            dst, w = (ch.D_I128, 16) if col.type_id in (ch.T_DECIMAL, ch.T_HUGEINT) else (ch.D_SAME, 8)
            nc = check_fixed_window(ora, j, dst, w, np.frombuffer(a.buffers()[1], dtype=np.uint8),
                                    np.frombuffer(a.buffers()[0], dtype=np.uint8), 0, n)
            assert a.null_count == nc
        assert arrays[0].type == pa.timestamp("us") and arrays[22].type == pa.decimal128(18, 3) and arrays[43].type == pa.decimal128(38, 0)
    ora.close()


# ------------------------------------------------------------------------------------------- C5
C5_ROWS = 50_000_000
C5_OFF = 3


def _c5_arrow(n, off, seed):
    """Arrow int32 id, int64 v, float64 x, bool flag, utf8 s (len U[0,24]); 10 % NULL except id; every array is
    a slice at element offset `off` of a longer one (non-zero bit offsets into bitmaps and bool values)"""
    rng = np.random.default_rng(seed)
    m = n + off

    def bitmap(p_null):
        return pa.py_buffer(np.packbits(rng.random(m) >= p_null, bitorder="little"))

    def nulls(buf):
        return m - int(np.unpackbits(np.frombuffer(buf, dtype=np.uint8), bitorder="little")[:m].sum())

    ids = pa.Array.from_buffers(pa.int32(), m, [None, pa.py_buffer(np.arange(m, dtype=np.int32))])
    bv = bitmap(0.1)
    v = pa.Array.from_buffers(pa.int64(), m, [bv, pa.py_buffer(rng.integers(-2**62, 2**62, m, dtype=np.int64))], null_count=nulls(bv))
    bx = bitmap(0.1)
    x = pa.Array.from_buffers(pa.float64(), m, [bx, pa.py_buffer(rng.standard_normal(m))], null_count=nulls(bx))
    bf = bitmap(0.1)
    flag = pa.Array.from_buffers(pa.bool_(), m, [bf, bitmap(0.5)], null_count=nulls(bf))
    lens = rng.integers(0, 25, m)
    offs = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(lens, out=offs[1:])
    data = rng.integers(0x20, 0x7F, int(offs[-1]) + 16, dtype=np.uint8)
    bs = bitmap(0.1)
    s = pa.Array.from_buffers(pa.string(), m, [bs, pa.py_buffer(offs), pa.py_buffer(data)], null_count=nulls(bs))
    rb = pa.record_batch([ids, v, x, flag, s], names=["id", "v", "x", "flag", "s"])
    return rb.slice(off, n)


class _SlabSink:
    """appender sink that lays the chunks out as the vector slabs the oracle produces"""

    def __init__(self, widths, nrows):
        self.widths = widths
        nch = (nrows + 2047) // 2048
        self.data = [np.zeros(nch * 2048 * w, dtype=np.uint8) for w in widths]
        self.val = [np.zeros(nch * 32, dtype=np.uint64) for _ in widths]
        self.k = 0
        self.rows = 0

    def __call__(self, count, vec_data, vec_validity):
        for c, w in enumerate(self.widths):
            C.memmove(self.data[c].ctypes.data + self.k * 2048 * w, vec_data[c], count * w)
            C.memmove(self.val[c].ctypes.data + self.k * 256, vec_validity[c], 256)
        self.k += 1
        self.rows += count
        return True


def test_c5_reverse_50m_rows_host_api(ctx):
    from duckdb_mbt_b200 import appender as ap
    from test_gpu_appender import C5_TYPES, _oracle_column
    rb = _c5_arrow(C5_ROWS, C5_OFF, 20260106)
    struct = rb.to_struct_array()
    widths = [4, 8, 8, 1, 16]
    sink = _SlabSink(widths, C5_ROWS)
    a = ap.Appender(ctx, C5_TYPES, sink)
    a.append_arrow(struct)
    a.flush()
    assert sink.rows == C5_ROWS and a.flushed_row_count == C5_ROWS
    for c, w in enumerate(widths):
        (exp_out, exp_val, _), w2 = _oracle_column(struct.field(c), C5_TYPES[c])
        assert w2 == w
        nb = C5_ROWS * w
        assert np.array_equal(sink.data[c][:nb], exp_out[:nb]), f"vector payload differs col={c}"
        assert np.array_equal(sink.val[c], exp_val[: sink.val[c].shape[0]]), f"validity masks differ col={c}"
    a.close()


def test_c5_reverse_50m_rows_l0():
    _need_gpu()
    from test_gpu_l0_reverse import _dev, _run_fixed
    n, off = C5_ROWS, C5_OFF
    rng = np.random.default_rng(77)
    raw = rng.integers(0, 2**63, n + off, dtype=np.int64).view(np.uint8)
    bitmap = np.packbits(rng.random(n + off + 8) >= 0.1, bitorder="little")
    exp_out, exp_val, exp_nc = oracle.rev_fixed(np.ascontiguousarray(raw[off * 8:]), bitmap, off, n, 3, 8)
    tv, pv = _dev(raw)
    tb, pb = _dev(bitmap, lead=1)
    got, got_val, got_nc = _run_fixed(pv + off * 8, pb, off, n, 3, 8)
    assert np.array_equal(got, exp_out[: n * 8])
    assert np.array_equal(got_val, exp_val) and got_nc == exp_nc
    # bool bits at a bit offset
    bits = np.packbits(rng.random(n + off + 8) >= 0.5, bitorder="little")
    exp_out, exp_val, exp_nc = oracle.rev_fixed(bits, bitmap, off, n, 5, 1)
    tv, pv = _dev(bits, lead=2)
    got, got_val, got_nc = _run_fixed(pv, pb, off, n, 5, 1)
    assert np.array_equal(got, exp_out[:n])
    assert np.array_equal(got_val, exp_val) and got_nc == exp_nc
