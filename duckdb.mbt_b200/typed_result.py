"""Host-side mirror of the reference's `duckdb_typed_result` surface (src/duckdb_typed_result.mbt,
src/duckdb.mbt:183-201), fed by typed columns converted on the GPU.

The reference builds a `TypedQueryResult` by rendering every cell to text inside libduckdb and
re-parsing it in MoonBit (`Connection::query` src/duckdb_native.mbt:454-501 ->
`QueryResult::to_typed` src/duckdb_typed_result.mbt:8-43 -> `parse_value_with_type`
src/duckdb_parsing.mbt:82-144).  Here the same column-major `Value` storage is filled from
`duckdb_mb_gpu_result_typed_column`: Int32-saturated integers, IEEE doubles, bool bytes, the
reference's day numbers (including its pre-1970 `date_to_days` behaviour), microsecond
timestamps and utf8 strings — no text round trip.

Documented deviations from the text path: doubles are the exact IEEE
value (the reference's `parse_double` is not correctly rounded and drops exponents), and
nan/inf stay `Value.Double` (`Value::String("nan")` in the reference, :100-105).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, List, Optional

import numpy as np

from . import native as nat
from .arrow_result import ArrowResult, DuckDBError

# enum dmb_value_tag / order of `Value` in src/duckdb.mbt:183-193
INT, DOUBLE, BOOL, STRING, DATE, TIMESTAMP, DECIMAL, BLOB, NULL = range(9)
TAG_NAMES = ["Int", "Double", "Bool", "String", "Date", "Timestamp", "Decimal", "Blob", "Null"]


@dataclass(frozen=True)
class Value:
    """`Value` of the reference: a tag and a payload (None for Null)."""
    tag: int
    payload: Any = None

    def is_null(self) -> bool:  # src/duckdb_typed_result.mbt:459
        return self.tag == NULL

    def as_int(self) -> Optional[int]:  # :387
        return self.payload if self.tag == INT else None

    def as_double(self) -> Optional[float]:  # :396
        return self.payload if self.tag == DOUBLE else None

    def as_bool(self) -> Optional[bool]:  # :405
        return self.payload if self.tag == BOOL else None

    def as_string(self) -> Optional[str]:  # :414
        return self.payload if self.tag == STRING else None

    def as_date(self) -> Optional[int]:  # :423
        return self.payload if self.tag == DATE else None

    def as_timestamp(self) -> Optional[int]:  # :432
        return self.payload if self.tag == TIMESTAMP else None

    def __repr__(self) -> str:
        return "Null" if self.tag == NULL else f"{TAG_NAMES[self.tag]}({self.payload!r})"


VALUE_NULL = Value(NULL)


class TypedColumn:
    """One column of typed values, kept columnar (numpy) until a cell or list is asked for."""

    def __init__(self, tag: int, values: Optional[np.ndarray], valid: np.ndarray,
                 offsets: Optional[np.ndarray] = None, data: Optional[bytes] = None):
        self.tag = tag
        self.values = values
        self.valid = valid
        self.offsets = offsets
        self.data = data

    def __len__(self) -> int:
        return int(self.valid.shape[0])

    def _payload(self, row: int):
        if self.tag == STRING:
            a, b = int(self.offsets[row]), int(self.offsets[row + 1])
            return self.data[a:b].decode("utf-8", errors="replace")
        v = self.values[row]
        if self.tag == BOOL:
            return bool(v)
        if self.tag == DOUBLE:
            return float(v)
        return int(v)

    def value(self, row: int) -> Value:
        if not self.valid[row]:
            return VALUE_NULL
        return Value(self.tag, self._payload(row))

    def to_options(self, tag: int) -> List[Optional[Any]]:
        """`get_<T>_column`: Some(v) for cells of that variant, None for NULL and any other variant
        (src/duckdb_typed_result.mbt:215-379)."""
        if tag != self.tag:
            return [None] * len(self)
        return [self._payload(i) if self.valid[i] else None for i in range(len(self))]


class TypedQueryResult:
    """`TypedQueryResult` (src/duckdb.mbt:198-201): column names + column-major values."""

    def __init__(self, columns: List[str], data: List[TypedColumn]):
        self.columns = columns
        self.data = data

    def row_count(self) -> int:  # src/duckdb_typed_result.mbt:51-57
        return 0 if not self.data else len(self.data[0])

    def column_count(self) -> int:  # :61-63
        return len(self.columns)

    def get_value(self, row: int, col: int) -> Optional[Value]:  # :67-78
        if row < 0 or row >= self.row_count() or col < 0 or col >= self.column_count():
            return None
        return self.data[col].value(row)

    def _get(self, row: int, col: int, tag: int):
        v = self.get_value(row, col)
        return v.payload if (v is not None and v.tag == tag) else None

    def get_int(self, row, col):  # :81
        return self._get(row, col, INT)

    def get_double(self, row, col):  # :94
        return self._get(row, col, DOUBLE)

    def get_bool(self, row, col):  # :107
        return self._get(row, col, BOOL)

    def get_string(self, row, col):  # :120
        return self._get(row, col, STRING)

    def get_date(self, row, col):  # :133
        return self._get(row, col, DATE)

    def get_timestamp(self, row, col):  # :146
        return self._get(row, col, TIMESTAMP)

    def is_null(self, row: int, col: int) -> bool:  # :185-197: out of range counts as null
        v = self.get_value(row, col)
        return True if v is None else v.is_null()

    def get_column(self, col: int) -> Optional[List[Value]]:  # :202-211
        if col < 0 or col >= self.column_count():
            return None
        c = self.data[col]
        return [c.value(i) for i in range(len(c))]

    def _options(self, col: int, tag: int):
        if col < 0 or col >= self.column_count():
            return None
        return self.data[col].to_options(tag)

    def get_int_column(self, col):  # :215
        return self._options(col, INT)

    def get_double_column(self, col):  # :239
        return self._options(col, DOUBLE)

    def get_bool_column(self, col):  # :263
        return self._options(col, BOOL)

    def get_string_column(self, col):  # :287
        return self._options(col, STRING)

    def get_date_column(self, col):  # :311
        return self._options(col, DATE)

    def get_timestamp_column(self, col):  # :335
        return self._options(col, TIMESTAMP)


_NUMPY_OF = {(INT, 4): np.dtype("<i4"), (DOUBLE, 8): np.dtype("<f8"), (BOOL, 1): np.dtype(np.uint8),
             (DATE, 4): np.dtype("<i4"), (TIMESTAMP, 8): np.dtype("<i8")}


def text_column(result: ArrowResult, col: int) -> TypedColumn:
    """The string form of one column (duckdb_mb_gpu_result_text_column): every cell's VARCHAR
    rendering as utf8 offsets + data, what `Connection::query` collects per cell."""
    return typed_column(result, col, _fn="duckdb_mb_gpu_result_text_column")


def typed_column(result: ArrowResult, col: int, _fn: str = "duckdb_mb_gpu_result_typed_column") -> TypedColumn:
    """One typed column straight from the GPU (duckdb_mb_gpu_result_typed_column)."""
    tc = nat.TypedColumn()
    if not getattr(result.lib, _fn)(result.handle, col, C.byref(tc)):
        raise DuckDBError(nat.last_error())
    n = int(tc.length)
    valid = np.frombuffer(C.string_at(tc.valid, n), dtype=np.uint8) != 0 if n else np.zeros(0, dtype=bool)
    if tc.tag == STRING:
        offsets = np.frombuffer(C.string_at(tc.offsets, 4 * (n + 1)), dtype="<i4")
        data = C.string_at(tc.data, int(offsets[-1])) if n and offsets[-1] else b""
        return TypedColumn(STRING, None, valid, offsets, data)
    dt = _NUMPY_OF[(tc.tag, tc.width)]
    values = np.frombuffer(C.string_at(tc.values, n * tc.width), dtype=dt) if n else np.zeros(0, dtype=dt)
    return TypedColumn(tc.tag, values, valid)


def to_typed(result: ArrowResult) -> TypedQueryResult:
    """Columnar `QueryResult::to_typed` (src/duckdb_typed_result.mbt:8-43)."""
    names = [f.name for f in result.get_schema().fields]
    return TypedQueryResult(names, [typed_column(result, j) for j in range(result.column_count())])
