"""Host-side mirror of the reference's `QueryResult` (src/duckdb.mbt:49-62,204-240): the string
form of a result — `rows[r][c]` the VARCHAR rendering of every cell, `nulls[r][c]` the null mask —
which `Connection::query` builds with two FFI calls per cell (`duckdb_mb_result_is_null` /
`duckdb_mb_result_value` -> `duckdb_value_varchar`, src/duckdb_native.mbt:477-497,
src/duckdb_native.c:215-238).

Here every column is rendered on the GPU (kernels_render.cu + the string kernel) and fetched
with ONE call per column (`duckdb_mb_gpu_result_text_column`: utf8 offsets + data + validity); the
row-major lists are only assembled when asked for.  `to_typed()` is the columnar `QueryResult::to_typed`
(src/duckdb_typed_result.mbt:8-43) and does not go through the text.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import typed_result as tr
from .arrow_result import ArrowResult, DuckDBError


class QueryResult:
    def __init__(self, columns: List[str], column_types: List[int], col_strings: List[List[str]], col_valid: List[np.ndarray]):
        self.columns = columns            # column names
        self.column_types = column_types  # DUCKDB_TYPE ids (ColumnType, src/duckdb_parsing.mbt:8-52)
        self._col_strings = col_strings
        self._col_valid = col_valid
        self._result: Optional[ArrowResult] = None

    @classmethod
    def from_result(cls, result: ArrowResult, type_ids: List[int]) -> "QueryResult":
        names = [f.name for f in result.get_schema().fields]
        strings, valid = [], []
        for j in range(result.column_count()):
            t = tr.text_column(result, j)  # raises DuckDBError for a type with no device renderer
            off, data = t.offsets, t.data
            strings.append([data[int(off[i]):int(off[i + 1])].decode("utf-8", errors="replace") for i in range(len(t))])
            valid.append(t.valid)
        q = cls(names, list(type_ids), strings, valid)
        q._result = result
        return q

    def row_count(self) -> int:  # src/duckdb.mbt:204-206
        return len(self._col_strings[0]) if self._col_strings else 0

    def column_count(self) -> int:  # :209-211
        return len(self.columns)

    def cell(self, row: int, col: int) -> Optional[str]:  # :214-221: None for NULL or out of range
        if row < 0 or row >= self.row_count() or col < 0 or col >= self.column_count():
            return None
        return self._col_strings[col][row] if self._col_valid[col][row] else None

    @property
    def rows(self) -> List[List[str]]:  # row-major strings; a NULL cell is "" (the reference stores the empty Bytes)
        return [[self._col_strings[c][r] if self._col_valid[c][r] else "" for c in range(self.column_count())]
                for r in range(self.row_count())]

    @property
    def nulls(self) -> List[List[bool]]:
        return [[not bool(self._col_valid[c][r]) for c in range(self.column_count())] for r in range(self.row_count())]

    def to_typed(self) -> tr.TypedQueryResult:
        if self._result is None:
            raise DuckDBError("to_typed needs the live result")
        return tr.to_typed(self._result)


# ---------------------------------------------------------------------------------------------
# The reference's own per-cell paths, call for call (the drop-in C symbols):
#   Connection::query      src/duckdb_native.mbt:454-501  (duckdb_mb_result_is_null / _value per cell)
#   Connection::query_stream / ResultStream::next  :504-582  (duckdb_mb_chunk_is_null / _value per cell)
class DataChunk:
    """`DataChunk` of src/duckdb.mbt:56-62: row-major strings + null mask of one fetched chunk."""

    def __init__(self, rows: List[List[str]], nulls: List[List[bool]]):
        self.rows = rows
        self.nulls = nulls

    def row_count(self) -> int:
        return len(self.rows)


class ResultStream:
    def __init__(self, result: ArrowResult):
        from . import native as nat
        self._nat = nat
        self.lib = result.lib
        self._result = result
        self.handle = self.lib.duckdb_mb_gpu_stream_from_result(result.handle)
        if self.lib.duckdb_mb_is_null_stream(self.handle):
            self.handle = None
            raise DuckDBError(nat.last_error())

    def columns(self) -> List[str]:  # src/duckdb_native.mbt:529-540
        n = self.lib.duckdb_mb_stream_column_count(self.handle)
        return [self._nat.moonbit_bytes(self.lib.duckdb_mb_stream_column_name(self.handle, j)).decode("utf-8", "replace") for j in range(n)]

    def next(self) -> Optional[DataChunk]:  # :543-582: None at the end of the stream, DuckDBError on a fetch error
        chunk = self.lib.duckdb_mb_stream_fetch_chunk(self.handle)
        if self.lib.duckdb_mb_is_null_chunk(chunk):
            err = self._nat.last_error()
            if err:
                raise DuckDBError(err)
            return None
        try:
            nrows = self.lib.duckdb_mb_chunk_row_count(chunk)
            ncols = self.lib.duckdb_mb_chunk_column_count(chunk)
            rows, nulls = [], []
            for r in range(nrows):
                row, nrow = [], []
                for c in range(ncols):
                    if self.lib.duckdb_mb_chunk_is_null(chunk, c, r):
                        row.append("")
                        nrow.append(True)
                    else:
                        row.append(self._nat.moonbit_bytes(self.lib.duckdb_mb_chunk_value(chunk, c, r)).decode("utf-8", "replace"))
                        nrow.append(False)
                rows.append(row)
                nulls.append(nrow)
            return DataChunk(rows, nulls)
        finally:
            self.lib.duckdb_mb_chunk_destroy(chunk)

    def close(self) -> None:
        if self.handle:
            self.lib.duckdb_mb_stream_destroy(self.handle)
            self.handle = None


def query_per_cell(result: ArrowResult) -> QueryResult:
    """`Connection::query`'s cell loop over the drop-in symbols (two calls per cell, like the reference)."""
    from . import native as nat
    L = result.lib
    ncols, nrows = L.duckdb_mb_result_column_count(result.handle), L.duckdb_mb_result_row_count(result.handle)
    names = [nat.moonbit_bytes(L.duckdb_mb_result_column_name(result.handle, j)).decode("utf-8", "replace") for j in range(ncols)]
    types = [L.duckdb_mb_result_column_type(result.handle, j) for j in range(ncols)]
    strings = [[""] * nrows for _ in range(ncols)]
    valid = [np.zeros(nrows, dtype=bool) for _ in range(ncols)]
    for r in range(nrows):
        for c in range(ncols):
            if not L.duckdb_mb_result_is_null(result.handle, c, r):
                valid[c][r] = True
                strings[c][r] = nat.moonbit_bytes(L.duckdb_mb_result_value(result.handle, c, r)).decode("utf-8", "replace")
    q = QueryResult(names, types, strings, valid)
    q._result = result
    return q
