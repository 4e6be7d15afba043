"""CPU-only: the C-ABI library loads and exports every symbol include/duckdb_mb_gpu.h declares;
struct layouts match the ctypes mirrors; the product path fails loudly without a GPU / library."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    import __graft_entry__ as g
    g.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "duckdb_mb_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b((?:dmb|duckdb_mb)_[a-z0-9_]+)\s*\(", text))
    return sorted(n for n in names if not n.startswith("dmb_chunk_sink"))


def test_library_exports_every_declared_symbol():
    _build()
    from duckdb_mbt_b200 import native as nat
    lib = C.CDLL(nat.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 60
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    # the binding's own list and the header agree
    assert sorted(set(nat.EXPORTED_SYMBOLS)) == declared


def test_drop_in_symbols_of_the_reference_are_present():
    # exactly what src/duckdb_arrow_native.mbt:9-104 binds (minus duckdb_mb_query_arrow, which needs libduckdb)
    _build()
    from duckdb_mbt_b200 import native as nat
    lib = C.CDLL(nat.LIB_PATH)
    names = ["duckdb_mb_arrow_column_count", "duckdb_mb_arrow_row_count", "duckdb_mb_arrow_schema", "duckdb_mb_arrow_destroy",
             "duckdb_mb_is_null_arrow_result", "duckdb_mb_bytes_to_double"]
    for kind in ("int32", "int64", "double", "string", "bool"):
        names += [f"duckdb_mb_arrow_get_column_{kind}", f"duckdb_mb_arrow_get_column_{kind}_nullable"]
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layouts():
    from duckdb_mbt_b200 import native as nat
    assert C.sizeof(nat.ArrowArray) == 80 and C.sizeof(nat.ArrowSchema) == 72  # Arrow C Data Interface
    assert C.sizeof(nat.VecDesc) == 16
    assert C.sizeof(nat.FixedJob) == 64
    assert C.sizeof(nat.StringJob) == 112
    assert C.sizeof(nat.RevFixedJob) == 56
    assert C.sizeof(nat.RevStringJob) == 72
    assert C.sizeof(nat.HostColumn) == 80 and C.sizeof(nat.HostList) == 48 and C.sizeof(nat.HostStruct) == 16 and C.sizeof(nat.HostBatch) == 32
    assert C.sizeof(nat.EnumDict) == 24 and C.sizeof(nat.EnumJob) == 72 and C.sizeof(nat.ListJob) == 112
    assert C.sizeof(nat.TypedColumn) == 56


def test_host_side_helpers_without_a_gpu():
    _build()
    from duckdb_mbt_b200 import chunks as ch
    from duckdb_mbt_b200 import native as nat
    L = nat.lib()
    for phys, w in enumerate(ch.PHYS_WIDTH):
        assert L.dmb_phys_width(phys) == w
    assert L.dmb_op_out_width(ch.op(ch.P_I64, ch.D_I128)) == 16
    assert L.dmb_op_out_width(ch.op(ch.P_BOOL, ch.D_BOOL_BITS)) == 0
    assert L.dmb_op_out_width(ch.op(ch.P_I64, ch.D_TS_REF_FROM_NS)) == 8
    assert L.dmb_op_out_width(ch.op(ch.P_STRING, ch.D_SAME)) == -1
    assert L.dmb_op_out_width(ch.op(ch.P_F64, ch.D_I32_SAT)) == -1
    # ticket + error flags + 4 tile status words per chunk + a sum word and a prefix word per group of 32 tiles (two-level look-back)
    assert L.dmb_dev_string_scratch_bytes(10) == (2 + 40 + 2 * 2) * 8
    assert L.dmb_dev_list_scratch_bytes(10) == (2 + 2 * 10 + 2 * 1) * 8
    buf = (C.c_uint8 * 16)()
    C.memmove(buf, C.byref(C.c_double(3.14)), 8)
    C.memmove(C.addressof(buf) + 8, C.byref(C.c_double(-2.5)), 8)
    assert L.duckdb_mb_bytes_to_double(C.addressof(buf), 0) == 3.14   # src/duckdb_native.c:2561-2565
    assert L.duckdb_mb_bytes_to_double(C.addressof(buf), 8) == -2.5
    assert L.duckdb_mb_is_null_arrow_result(None) == 1
    assert L.duckdb_mb_arrow_column_count(None) == 0 and L.duckdb_mb_arrow_row_count(None) == 0
    assert nat.moonbit_bytes(L.duckdb_mb_arrow_schema(None)) == b"[]"
    assert nat.moonbit_bytes(L.duckdb_mb_arrow_get_column_int32(None, 0)) == b""  # :2360-2362


def test_no_cpu_fallback():
    _build()
    import torch
    from duckdb_mbt_b200 import arrow_result as ar
    from duckdb_mbt_b200 import native as nat
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    with pytest.raises(ar.DuckDBError, match="no CUDA device"):
        ar.GpuContext(0)
    assert nat.lib().duckdb_mb_gpu_device_count() == 0


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from duckdb_mbt_b200 import native as nat
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", str(tmp_path / "libduckdb_mb_gpu.so"))
    with pytest.raises(nat.NativeLibraryMissing):
        nat.lib()
