"""GPU parity of the text-rendering path (K7 render_text + the string kernel) through the C ABI:
`duckdb_mb_arrow_get_column_string[_nullable]` on non-VARCHAR columns (the getter the reference's
schema JSON prescribes for DATE / DECIMAL / TIMESTAMP columns, src/duckdb_native.c:2314-2339,
:2456-2514), the `Value::String` typed column form of DECIMAL (src/duckdb_parsing.mbt:120-141) and
the string-form `QueryResult` (src/duckdb.mbt:49-62), bit-exact against the oracle and against the
reference's fixture strings (src/duckdb_fixture_cases.mbt)."""
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


@pytest.mark.parametrize("n,pattern", [(1, "full"), (2049, "full"), (10_000, "ragged"), (40_001, "ragged")])
def test_string_getter_on_fixed_width_columns(ctx, n, pattern):
    batch = _mixed_batch(n, pattern, 900 + n)
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        seen = 0
        for j, c in enumerate(batch.columns):
            if c.name not in RENDERED:
                continue
            for nullable in (False, True):
                got = res.raw_column("string", j, nullable)
                exp = ora.get_column("string", j, nullable)
                assert got == exp, f"col={c.name} nullable={nullable}"
            seen += 1
        assert seen == len(RENDERED)


def test_fixture_strings_through_gpu(ctx):
    # src/duckdb_fixture_cases.mbt: integer extremes :27-32,83-102; BOOLEAN :20-25; DATE :41-46,62-67;
    # TIMESTAMP :55-60; DECIMAL(10,3) :69-74, DECIMAL(9,2) :76-81
    micros = (19877 * 86400 + 12 * 3600 + 34 * 60 + 56) * 1_000_000 + 789_000
    batch = batch_of(
        ("big", ch.T_BIGINT, [9223372036854775807, -9223372036854775808, 0, None]),
        ("small", ch.T_SMALLINT, [32767, -32768, 0, None]),
        ("tiny", ch.T_TINYINT, [127, -128, 0, None]),
        ("int", ch.T_INTEGER, [2147483647, -2147483648, 0, None]),
        ("b", ch.T_BOOLEAN, [1, 0, 1, None]),
        ("d", ch.T_DATE, [19877, -1, 1, None]),
        ("ts", ch.T_TIMESTAMP, [micros, 0, -1, None]),
        ("dec", ch.T_DECIMAL, [123456, 5, -5, None], 10, 3),
        ("dec2", ch.T_DECIMAL, [-99999999, 0, 100, None], 9, 2),
        ("u64", ch.T_UBIGINT, [18446744073709551615, 0, 1, None]),
    )
    expect = [
        ["9223372036854775807", "-9223372036854775808", "0"],
        ["32767", "-32768", "0"],
        ["127", "-128", "0"],
        ["2147483647", "-2147483648", "0"],
        ["true", "false", "true"],
        ["2024-06-03", "1969-12-31", "1970-01-02"],
        ["2024-06-03 12:34:56.789", "1970-01-01 00:00:00", "1969-12-31 23:59:59.999999"],
        ["123.456", "0.005", "-0.005"],
        ["-999999.99", "0.00", "1.00"],
        ["18446744073709551615", "0", "1"],
    ]
    from duckdb_mbt_b200 import typed_result as tr
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        for j, exp in enumerate(expect):
            t = tr.text_column(res, j)
            assert [t.value(i).as_string() for i in range(3)] == exp, j
            assert t.valid.tolist() == [True, True, True, False] and t.value(3).is_null()
            assert int(t.offsets[4]) == int(t.offsets[3])  # NULL cell: zero-length
            # the reference-format blob of the same column equals the oracle's (incl. its NULL-terminator defect)
            for nullable in (False, True):
                assert res.raw_column("string", j, nullable) == ora.get_column("string", j, nullable)


def test_decimal_typed_column_is_value_string(ctx):
    from duckdb_mbt_b200 import typed_result as tr
    rng = np.random.default_rng(11)
    n = 7000
    counts = ch.chunk_counts(n, "ragged", rng)
    valid = rng.random(n) > 0.2
    vals = rng.integers(-10**17, 10**17, n, dtype=np.int64)
    batch = ch.ChunkBatch(counts, [ch.fixed_column("dec", ch.T_DECIMAL, vals, counts, valid=valid, dec_width=18, dec_scale=3, garbage_rng=rng)])
    with _result(ctx, batch) as res:
        tc = tr.typed_column(res, 0)
        assert tc.tag == tr.STRING
        assert np.array_equal(tc.valid, valid)
        for i in (0, 1, 17, n - 1):
            v = tc.value(i)
            if valid[i]:
                assert v.as_string() == oracle.render_decimal64(int(vals[i]), 3)
            else:
                assert v.is_null()


def test_query_result_string_form(ctx):
    from duckdb_mbt_b200.query_result import QueryResult
    batch = batch_of(("id", ch.T_INTEGER, [1, 2, None]),
                     ("name", ch.T_VARCHAR, ["a", None, "ccc"]),
                     ("d", ch.T_DATE, [19877, None, 0]),
                     ("p", ch.T_DECIMAL, [1050, 99, None], 10, 2))
    with _result(ctx, batch) as res:
        q = QueryResult.from_result(res, [c.type_id for c in batch.columns])
        assert q.row_count() == 3 and q.column_count() == 4
        assert q.rows == [["1", "a", "2024-06-03", "10.50"], ["2", "", "", "0.99"], ["", "ccc", "1970-01-01", ""]]
        assert q.nulls == [[False, False, False, False], [False, True, True, False], [True, False, False, True]]
        assert q.cell(0, 3) == "10.50" and q.cell(1, 1) is None and q.cell(9, 0) is None
        t = q.to_typed()
        assert t.get_int(0, 0) == 1 and t.is_null(2, 0)
        assert t.get_date(0, 2) == 19877
        assert t.get_string(0, 3) == "10.50"


def _wide(vals):
    a = np.zeros((len(vals), 16), np.uint8)
    for i, v in enumerate(vals):
        a[i] = np.frombuffer((v & ((1 << 128) - 1)).to_bytes(16, "little"), np.uint8)
    return a


def test_hugeint_time_and_wide_decimal_renderings(ctx):
    # HUGEINT is what SUM() over an integer column returns: src/duckdb_fixture_cases.mbt:34-37 ("6"), :174-177 ("15");
    # TIME '12:34:56.789' :48-51
    from duckdb_mbt_b200 import typed_result as tr
    n = 4
    counts = ch.chunk_counts(n)
    cols = [ch.fixed_column("h", ch.T_HUGEINT, _wide([6, 15, 2**127 - 1, -2**127]), counts),
            ch.fixed_column("u", ch.T_UHUGEINT, _wide([0, 2**128 - 1, 10**19, 10**38]), counts),
            ch.fixed_column("t", ch.T_TIME, np.asarray([45296789000, 0, 86399999999, 1], np.int64), counts),
            ch.fixed_column("tn", ch.T_TIME_NS, np.asarray([45296789000000, 0, 86399999999999, 1000], np.int64), counts),
            ch.fixed_column("d", ch.T_DECIMAL, _wide([5, -5, 10**37, -123456789012345678901234567890]), counts, dec_width=38, dec_scale=3)]
    expect = [["6", "15", "170141183460469231731687303715884105727", "-170141183460469231731687303715884105728"],
              ["0", "340282366920938463463374607431768211455", "10000000000000000000", "100000000000000000000000000000000000000"],
              ["12:34:56.789", "00:00:00", "23:59:59.999999", "00:00:00.000001"],
              ["12:34:56.789", "00:00:00", "23:59:59.999999999", "00:00:00.000001"],
              ["0.005", "-0.005", "10000000000000000000000000000000000.000", "-123456789012345678901234567.890"]]
    batch = ch.ChunkBatch(counts, cols)
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        for j, exp in enumerate(expect):
            t = tr.text_column(res, j)
            assert [t.value(i).as_string() for i in range(n)] == exp, j
            assert [ora.cell_value(j, i).decode() for i in range(n)] == exp, j
        # HUGEINT / DECIMAL(38) / TIME stay Value::String in the reference's typed result (src/duckdb_parsing.mbt:120-141)
        assert tr.typed_column(res, 0).tag == tr.STRING and tr.typed_column(res, 0).value(0).as_string() == "6"
        assert tr.typed_column(res, 2).value(0).as_string() == "12:34:56.789"


def test_uuid_timetz_interval_renderings(ctx):
    # the remaining scalar types of the reference's stream whitelist (src/duckdb_native.c:287-299): device text == oracle text,
    # known answers first, then random cells (NULLs, garbage under NULLs, ragged chunks)
    from duckdb_mbt_b200 import typed_result as tr
    from test_oracle_golden import INTERVAL_CASES, TIMETZ_CASES, UUID_CASES, _interval_rows, _uuid_bytes
    n = len(UUID_CASES)
    one = ch.chunk_counts(n)
    batch = ch.ChunkBatch(one, [ch.fixed_column("u", ch.T_UUID, np.stack([_uuid_bytes(t) for t in UUID_CASES]), one),
                                ch.fixed_column("z", ch.T_TIME_TZ, np.asarray([c[0] for c in TIMETZ_CASES], np.uint64), one)])
    with _result(ctx, batch) as res:
        assert [tr.text_column(res, 0).value(i).as_string() for i in range(n)] == UUID_CASES
        assert [tr.text_column(res, 1).value(i).as_string() for i in range(n)] == [c[1] for c in TIMETZ_CASES]
    m = len(INTERVAL_CASES)
    cm = ch.chunk_counts(m)
    with _result(ctx, ch.ChunkBatch(cm, [ch.fixed_column("i", ch.T_INTERVAL, _interval_rows([c[0] for c in INTERVAL_CASES]), cm)])) as res:
        assert [tr.text_column(res, 0).value(i).as_string() for i in range(m)] == [c[1] for c in INTERVAL_CASES]
    rng = np.random.default_rng(33)
    n = 7000
    counts = ch.chunk_counts(n, "ragged", rng)
    valid = rng.random(n) > 0.2
    iv = np.zeros(n, dtype=np.dtype([("m", "<i4"), ("d", "<i4"), ("us", "<i8")]))
    iv["m"] = rng.integers(-400, 400, n) * (rng.random(n) < 0.7)
    iv["d"] = rng.integers(-5000, 5000, n) * (rng.random(n) < 0.7)
    iv["us"] = rng.integers(-2**46, 2**46, n) * (rng.random(n) < 0.7) * np.where(rng.random(n) < 0.3, 1000, 1)
    iv[:4] = [(-(2**31), -(2**31), -(2**63)), (2**31 - 1, 2**31 - 1, 2**63 - 1), (0, 0, 0), (1, 1, 1)]
    tz = (rng.integers(0, 86400 * 10**6, n).astype(np.uint64) << np.uint64(24)) | rng.integers(0, 2 * 57599 + 1, n).astype(np.uint64)
    tz[rng.random(n) < 0.5] &= ~np.uint64(0xFFFFFF) | np.uint64(57599 - 3600 * 2)
    batch = ch.ChunkBatch(counts, [ch.fixed_column("u", ch.T_UUID, rng.integers(0, 256, (n, 16), dtype=np.uint8), counts, valid=valid, garbage_rng=rng),
                                   ch.fixed_column("z", ch.T_TIME_TZ, tz, counts, valid=valid, garbage_rng=rng),
                                   ch.fixed_column("i", ch.T_INTERVAL, iv.view(np.uint8).reshape(n, 16), counts, valid=valid, garbage_rng=rng)])
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        for j in range(3):
            for nullable in (False, True):
                assert res.raw_column("string", j, nullable) == ora.get_column("string", j, nullable), (j, nullable)


def test_random_hugeint_column_against_the_oracle(ctx):
    rng = np.random.default_rng(21)
    n = 9000
    counts = ch.chunk_counts(n, "ragged", rng)
    valid = rng.random(n) > 0.2
    raw = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    raw[rng.random(n) < 0.3, 8:] = 0          # small positive values too
    raw[rng.random(n) < 0.1, 1:] = 0
    batch = ch.ChunkBatch(counts, [ch.fixed_column("h", ch.T_HUGEINT, raw, counts, valid=valid, garbage_rng=rng),
                                   ch.fixed_column("d", ch.T_DECIMAL, raw.copy(), counts, valid=valid, dec_width=30, dec_scale=7, garbage_rng=rng)])
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        for j in range(2):
            for nullable in (False, True):
                assert res.raw_column("string", j, nullable) == ora.get_column("string", j, nullable), (j, nullable)


def test_double_shortest_round_trip_rendering(ctx):
    # fixtures: "3.5" :20-25, "5.333333333333333" :167-172, nan / inf / -inf :209-214, "3.14159265359" ... :216-221;
    # plus zeros, subnormals, extremes, powers of two (half-width rounding interval below them) and random bit patterns
    import struct
    from duckdb_mbt_b200 import typed_result as tr
    fixed = [3.5, 16.0 / 3.0, float("nan"), float("inf"), float("-inf"), 3.14159265359, 2.71828182846, 1.41421356237,
             0.0, -0.0, 1.0, -1.0, 0.1, 1e15, 1e16, 1e-5, 1e-6, 123456789012345678.0, 5e-324, 1.7976931348623157e308,
             2.2250738585072014e-308, 9007199254740993.0, 1e22, 1e23, 7.120236347223045e-307, 999999.0, 0.3]
    expect = ["3.5", "5.333333333333333", "nan", "inf", "-inf", "3.14159265359", "2.71828182846", "1.41421356237",
              "0.0", "-0.0", "1.0", "-1.0", "0.1", "1000000000000000.0", "1e+16", "0.00001", "1e-06", "1.2345678901234568e+17",
              "5e-324", "1.7976931348623157e+308", "2.2250738585072014e-308", "9007199254740992.0", "1e+22", "1e+23",
              "7.120236347223045e-307", "999999.0", "0.3"]
    rng = np.random.default_rng(33)
    rnd = rng.integers(0, 2**64 - 1, 20000, dtype=np.uint64).view(np.float64)
    pows = np.asarray([2.0 ** k for k in range(-1074, 1024)])
    near = rng.standard_normal(10000) * 10.0 ** rng.integers(-12, 12, 10000)
    ints = rng.integers(-10**9, 10**9, 5000).astype(np.float64)
    vals = np.concatenate([np.asarray(fixed), rnd, pows, near, ints])
    n = vals.shape[0]
    counts = ch.chunk_counts(n)
    f32 = (rng.standard_normal(n) * 1e3).astype(np.float32)
    batch = ch.ChunkBatch(counts, [ch.fixed_column("d", ch.T_DOUBLE, vals, counts), ch.fixed_column("f", ch.T_FLOAT, f32, counts)])
    ora = oracle.OracleResult(batch)
    with _result(ctx, batch) as res:
        t = tr.text_column(res, 0)
        assert [t.value(i).as_string() for i in range(len(fixed))] == expect
        for j in range(2):
            t = tr.text_column(res, j)
            got = [t.data[int(t.offsets[i]):int(t.offsets[i + 1])] for i in range(n)]
            exp = [ora.cell_value(j, i) for i in range(n)]
            bad = [i for i in range(n) if got[i] != exp[i]]
            assert not bad, (j, bad[:5], [(got[i], exp[i]) for i in bad[:5]])
        # every finite rendering reads back as the same double
        t = tr.text_column(res, 0)
        for i in rng.integers(0, n, 3000):
            s_ = t.value(int(i)).as_string()
            if s_ not in ("nan", "inf", "-inf"):
                assert struct.pack("<d", float(s_)) == struct.pack("<d", float(vals[i])), (s_, vals[i])


def test_all_reference_fixture_cases_replayed_on_the_gpu(ctx):
    """Every case of src/duckdb_fixture_cases.mbt (tests/golden/reference_fixture_cases.json) through the device:
    the columnar string form (K7 render + K5) and Connection::query's per-cell symbols must give the fixture's
    rows and null masks -- the assertion of src/duckdb_test.mbt:88."""
    import golden_cases as gc
    from duckdb_mbt_b200.query_result import QueryResult, query_per_cell
    for case in gc.CASES:
        batch = gc.batch_for(case)
        with _result(ctx, batch) as res:
            q = QueryResult.from_result(res, [c.type_id for c in batch.columns])
            assert q.columns == case["columns"], case["name"]
            assert q.rows == case["rows"], case["name"]
            assert q.nulls == case["nulls"], case["name"]
            p = query_per_cell(res)
            assert p.rows == case["rows"] and p.nulls == case["nulls"], case["name"]
