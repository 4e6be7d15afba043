"""Row-group sharding across GPUs: contiguous chunk ranges, no data-path collective.

The chunk list of a result is cut into `world` contiguous ranges of whole chunks; GPU g converts
range g independently (own context, own host link) and returns an independent Arrow record batch
whose utf8 offsets start at 0.  When one logical column is wanted, the only cross-GPU datum is one
byte total per string column per GPU: an exclusive scan of those `world` integers on the host gives
the base each GPU's offsets are shifted by (large_utf8) — no NCCL, NVLink unused.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_chunks(nchunks: int, world: int, rank: int) -> Tuple[int, int]:
    """Chunk range [c0, c1) of `rank`: GPU g gets chunks [g*ceil(C/G), (g+1)*ceil(C/G))."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-nchunks // world) if nchunks > 0 else 0
    c0 = min(rank * per, nchunks)
    c1 = min(c0 + per, nchunks)
    return c0, c1


def shard_rows(counts: np.ndarray, world: int, rank: int) -> Tuple[int, int, int, int]:
    """(c0, c1, row0, row1) for `rank` given per-chunk row counts."""
    c0, c1 = shard_chunks(int(counts.shape[0]), world, rank)
    ro = np.zeros(counts.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, dtype=np.int64, out=ro[1:])
    return c0, c1, int(ro[c0]), int(ro[c1])


def string_bases(per_gpu_totals: Sequence[int]) -> List[int]:
    """Host exclusive scan of the per-GPU utf8 byte totals of one string column."""
    out, acc = [], 0
    for t in per_gpu_totals:
        out.append(acc)
        acc += int(t)
    return out


def concat_utf8(parts: Sequence[Tuple[np.ndarray, bytes]]) -> Tuple[np.ndarray, bytes]:
    """Stitch per-GPU (int32/int64 offsets[n_g+1], data) into one large_utf8 column."""
    bases = string_bases([len(d) for _, d in parts])
    offs = [np.asarray(o[:-1], dtype=np.int64) + b for (o, _), b in zip(parts, bases)]
    total = bases[-1] + len(parts[-1][1]) if parts else 0
    offsets = np.concatenate(offs + [np.asarray([total], dtype=np.int64)]) if parts else np.zeros(1, dtype=np.int64)
    return offsets, b"".join(d for _, d in parts)


def slice_batch(batch, c0: int, c1: int):
    """Sub-batch of chunks [c0, c1) sharing the parent's slabs (host side)."""
    from . import chunks as ch

    cols = []
    for col in batch.columns:
        cols.append(ch.Column(col.name, col.type_id, col.phys, col.data, col.data_off[c0:c1], col.validity,
                              col.val_off[c0:c1], col.dec_width, col.dec_scale, col.heap))
    return ch.ChunkBatch(batch.counts[c0:c1].copy(), cols)
