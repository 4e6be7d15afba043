"""The reference's 35 fixture cases (tests/golden/reference_fixture_cases.json, extracted from
src/duckdb_fixture_cases.mbt:4-262 by tests/golden/make_golden.py) as chunk inputs.

There is no SQL engine in the image, so every case names the result vectors DuckDB produces for its SQL (logical
type + physical values; DuckDB's typing rules: integer literals are INTEGER, `3.5` is DECIMAL(2,1), count(*) and
range() are BIGINT, sum(INTEGER) is HUGEINT, avg() is DOUBLE, an untyped NULL column comes back as INTEGER).  The
expected cell strings and null masks are the fixture's own; VARCHAR columns take their payload from them."""
import json
import os

import numpy as np

from duckdb_mbt_b200 import chunks as ch

_HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(_HERE, "golden", "reference_fixture_cases.json"), encoding="utf-8"))["cases"]

V = "varchar"  # payload = the expected strings
_TS = (19877 * 86400 + 12 * 3600 + 34 * 60 + 56) * 1_000_000 + 789_000

# case name -> one entry per column: V | (type_id, values[, dec_width, dec_scale])
INPUTS = {
    "basic select": [(ch.T_INTEGER, [1]), (ch.T_INTEGER, [None]), V],
    "multi row values": [(ch.T_INTEGER, [1, 2, 3]), V],
    "boolean and double": [(ch.T_BOOLEAN, [1]), (ch.T_BOOLEAN, [0]), (ch.T_DECIMAL, [35], 2, 1)],
    "bigint extremes": [(ch.T_BIGINT, [9223372036854775807]), (ch.T_BIGINT, [-9223372036854775808])],
    "simple aggregate": [(ch.T_BIGINT, [3]), (ch.T_HUGEINT, [6])],
    "date literal": [(ch.T_DATE, [19877])],
    "time literal": [(ch.T_TIME, [45296789000])],
    "timestamp literal": [(ch.T_TIMESTAMP, [_TS])],
    "epoch date arithmetic": [(ch.T_DATE, [-1]), (ch.T_DATE, [1])],
    "decimal positive": [(ch.T_DECIMAL, [123456], 10, 3)],
    "decimal negative": [(ch.T_DECIMAL, [-99999999], 9, 2)],
    "smallint extremes": [(ch.T_SMALLINT, [32767]), (ch.T_SMALLINT, [-32768])],
    "tinyint range": [(ch.T_TINYINT, [127]), (ch.T_TINYINT, [-128])],
    "integer extremes": [(ch.T_INTEGER, [2147483647]), (ch.T_INTEGER, [-2147483648])],
    "string escapes": [V, V],
    "string whitespace": [V, V],
    "empty strings": [V, V],
    "single null": [(ch.T_INTEGER, [None])],
    "multiple nulls": [(ch.T_INTEGER, [None])] * 3,
    "mixed nulls": [(ch.T_INTEGER, [1]), (ch.T_INTEGER, [None]), (ch.T_INTEGER, [2]), (ch.T_INTEGER, [None]), (ch.T_INTEGER, [3])],
    "range function": [(ch.T_BIGINT, [0, 1, 2, 3, 4])],
    "range with expression": [(ch.T_BIGINT, [0, 1, 2]), (ch.T_BIGINT, [0, 2, 4])],
    "range with modulo": [(ch.T_BIGINT, [0, 1, 2, 3]), (ch.T_BOOLEAN, [1, 0, 1, 0])],
    "multiple aggregates": [(ch.T_BIGINT, [3]), (ch.T_INTEGER, [1]), (ch.T_INTEGER, [10]), (ch.T_DOUBLE, [16.0 / 3.0])],
    "sum aggregate": [(ch.T_HUGEINT, [15])],
    "boolean logic": [(ch.T_BOOLEAN, [0]), (ch.T_BOOLEAN, [1]), (ch.T_BOOLEAN, [0])],
    "comparison boolean": [(ch.T_BOOLEAN, [0, 0, 1])],
    "cast str to int": [(ch.T_INTEGER, [42]), V],
    "cast null types": [V, (ch.T_INTEGER, [None])],
    "float special values": [(ch.T_DOUBLE, [float("nan")]), (ch.T_DOUBLE, [float("inf")]), (ch.T_DOUBLE, [float("-inf")])],
    "double precision": [(ch.T_DECIMAL, [314159265359], 12, 11), (ch.T_DECIMAL, [271828182846], 12, 11), (ch.T_DECIMAL, [141421356237], 12, 11)],
    "null in values": [(ch.T_INTEGER, [1, None, 3]), (ch.T_INTEGER, [None, 2, 3])],
    "case expression": [(ch.T_INTEGER, [1, 2, 3, 4, 5]), V],
    "string concatenation": [V],
    "coalesce function": [V, V, V],
}


def _wide(vals):
    a = np.zeros((len(vals), 16), np.uint8)
    for i, v in enumerate(vals):
        a[i] = np.frombuffer(((v or 0) & ((1 << 128) - 1)).to_bytes(16, "little"), np.uint8)
    return a


def batch_for(case) -> ch.ChunkBatch:
    """the DataChunk DuckDB returns for the case's SQL"""
    spec = INPUTS[case["name"]]
    n = len(case["rows"])
    assert len(spec) == len(case["columns"]), case["name"]
    counts = ch.chunk_counts(n, "full")
    cols = []
    for j, (name, s) in enumerate(zip(case["columns"], spec)):
        if s == V:
            strs = [None if case["nulls"][i][j] else case["rows"][i][j].encode() for i in range(n)]
            cols.append(ch.string_column(name, strs, counts, ch.T_VARCHAR))
            continue
        type_id, values = s[0], s[1]
        dec_w, dec_s = (s[2], s[3]) if len(s) > 2 else (0, 0)
        assert len(values) == n, case["name"]
        valid = np.asarray([v is not None for v in values], dtype=bool)
        phys = ch.phys_of_type(type_id, dec_w)
        arr = _wide(values) if phys in (ch.P_I128, ch.P_U128) else np.asarray([0 if v is None else v for v in values], dtype=ch.PHYS_NUMPY[phys])
        cols.append(ch.fixed_column(name, type_id, arr, counts, valid=None if valid.all() else valid, dec_width=dec_w, dec_scale=dec_s))
    return ch.ChunkBatch(counts, cols)
