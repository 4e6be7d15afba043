"""duckdb.mbt_b200 — B200-native result/ingest boundary of the MoonBit DuckDB bindings.

Only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of include/duckdb_mb_gpu.h),
`chunks` (DuckDB-shaped chunk batches), `native` (ctypes binding of the C ABI), the host-side
mirrors of the reference's MoonBit API (`arrow_result`, `typed_result`, `appender`) and `shard`
(row-group partitioning across GPUs).  `device` / `devgen` (torch plumbing for the L0 tests and the
bench) are imported on demand.
"""
from . import chunks  # noqa: F401

__all__ = ["chunks", "native", "arrow_result", "typed_result", "appender", "shard"]
