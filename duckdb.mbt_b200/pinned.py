"""Page-locked host memory for chunk batches (duckdb_mb_gpu_host_alloc).

A caller that can place DuckDB vectors / Arrow buffers in page-locked memory sets
DMB_BATCH_PINNED and the stager DMAs them directly (no bounce copy through the staging ring).
bench.py's `e2e` leg and the pinned-path tests build their inputs with these helpers.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from . import chunks as ch
from . import native as nat

_live = {}


def pinned_empty(nbytes: int) -> np.ndarray:
    """uint8[nbytes] backed by cudaHostAlloc'ed memory (freed with pinned_free)."""
    L = nat.lib()
    nbytes = int(nbytes)
    p = L.duckdb_mb_gpu_host_alloc(max(nbytes, 1))
    if not p:
        raise MemoryError(nat.last_error())
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p)
    a = np.frombuffer(buf, dtype=np.uint8, count=nbytes)
    _live[a.ctypes.data if nbytes else p] = p
    return a


def pinned_free(a: np.ndarray) -> None:
    p = _live.pop(a.ctypes.data, None)
    if p:
        nat.lib().duckdb_mb_gpu_host_free(p)


def pinned_copy(a: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    out = pinned_empty(src.shape[0])
    out[:] = src
    return out


def pin_batch(batch: ch.ChunkBatch) -> ch.ChunkBatch:
    """Copy every slab of a host ChunkBatch into page-locked memory.  string_t pointers are
    re-pointed at the pinned copy of the heap (they must stay real addresses)."""
    cols: List[ch.Column] = []
    for col in batch.columns:
        data = pinned_copy(col.data)
        validity = None if col.validity is None else pinned_copy(col.validity).view(np.uint64)
        heap = None
        if col.heap is not None:
            heap = pinned_copy(col.heap)
            delta = np.uint64((int(heap.ctypes.data) - int(col.heap.ctypes.data)) % (1 << 64))
            ent = data.reshape(-1, 16)
            lens = ent[:, 0:4].copy().view(np.uint32).reshape(-1)
            is_ptr = lens > 12
            if is_ptr.any():
                ptrs = ent[is_ptr, 8:16].copy().view(np.uint64).reshape(-1)
                with np.errstate(over="ignore"):
                    ptrs = ptrs + delta
                ent[is_ptr, 8:16] = ptrs.view(np.uint8).reshape(-1, 8)
        cols.append(ch.Column(col.name, col.type_id, col.phys, data, col.data_off, validity, col.val_off,
                              col.dec_width, col.dec_scale, heap))
    return ch.ChunkBatch(batch.counts, cols)


def free_batch(batch: ch.ChunkBatch) -> None:
    for col in batch.columns:
        pinned_free(col.data)
        if col.validity is not None:
            pinned_free(col.validity.view(np.uint8))
        if col.heap is not None:
            pinned_free(col.heap)
