"""L0 harness: a chunk batch resident in HBM + launches of the device API (dmb_dev_*).

torch is plumbing only: device allocations (uint8 tensors), the current stream, H<->D copies for
tests.  All conversion work is done by the hand-written kernels in csrc/ through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import chunks as ch
from . import native as nat


def _dev(a: np.ndarray, device) -> torch.Tensor:
    """numpy (any dtype) -> uint8 device tensor (256-B aligned by the caching allocator)."""
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a.view(np.uint8).reshape(-1))
    return t.to(device, non_blocking=False)


def _zeros(nbytes: int, device) -> torch.Tensor:
    return torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _empty(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def vec_descs(col: ch.Column) -> np.ndarray:
    d = np.zeros(col.data_off.shape[0], dtype=[("data_off", np.uint64), ("val_off", np.int64)])
    d["data_off"] = col.data_off
    d["val_off"] = col.val_off
    return d


@dataclass
class FixedOut:
    col: int
    op: int
    width: int  # bytes per value; 0 = bit-packed
    values: Optional[torch.Tensor]
    bitmap: Optional[torch.Tensor]
    valid_bytes: Optional[torch.Tensor]
    null_count: torch.Tensor


@dataclass
class StringOut:
    col: int
    mode: int
    offsets: torch.Tensor
    data: torch.Tensor
    total: torch.Tensor
    scratch: torch.Tensor
    job: nat.StringJob


class DeviceBatch:
    """DuckDB-shaped chunk batch uploaded to one GPU (column slabs, validity slabs, descriptors)."""

    def __init__(self, batch: ch.ChunkBatch, device: str = "cuda:0"):
        self.lib = nat.lib()
        self.device = torch.device(device)
        self.batch = batch
        self.nrows = batch.nrows
        self.nchunks = batch.nchunks
        self.counts = _dev(batch.counts.astype(np.uint32), self.device)
        self.row_off = _dev(batch.row_off, self.device)
        self.data: List[torch.Tensor] = []
        self.validity: List[Optional[torch.Tensor]] = []
        self.vecs: List[torch.Tensor] = []
        self.heap: List[Optional[torch.Tensor]] = []
        for col in batch.columns:
            self.data.append(_dev(col.data, self.device))
            self.validity.append(None if col.validity is None else _dev(col.validity, self.device))
            self.vecs.append(_dev(vec_descs(col), self.device))
            if col.heap is not None:
                padded = np.zeros(col.heap.shape[0] + 16, dtype=np.uint8)
                padded[: col.heap.shape[0]] = col.heap
                self.heap.append(_dev(padded, self.device))
            else:
                self.heap.append(None)

    # ---------------------------------------------------------------- fixed width (K1-K4)
    def plan_fixed(self, specs: Sequence[Tuple[int, int]], bitmap: bool = True, valid_bytes: bool = False):
        """specs: (column index, dst kind or OP_VALIDITY_ONLY).  Allocates outputs and builds the job
        arrays sorted by op (one launch per distinct conversion)."""
        outs: List[FixedOut] = []
        n = self.nrows
        for col_idx, dst in specs:
            col = self.batch.columns[col_idx]
            if dst == ch.OP_VALIDITY_ONLY:
                op, width, values = ch.OP_VALIDITY_ONLY, 0, None
            else:
                op = ch.op(col.phys, dst)
                width = self.lib.dmb_op_out_width(op)
                if width < 0:
                    raise ValueError(f"unsupported conversion phys={col.phys} dst={dst}")
                values = _empty((n + 7) // 8 + 8 if width == 0 else n * width, self.device)
            outs.append(FixedOut(col_idx, op, width, values,
                                 _empty((n + 63) // 64 * 8, self.device) if bitmap else None,
                                 _empty(n, self.device) if valid_bytes else None,
                                 torch.zeros(8, dtype=torch.uint8, device=self.device)))
        outs.sort(key=lambda o: o.op)
        jobs = (nat.FixedJob * len(outs))()
        for j, o in enumerate(outs):
            v = self.validity[o.col]
            jobs[j] = nat.FixedJob(self.data[o.col].data_ptr(), v.data_ptr() if v is not None else None,
                                   self.vecs[o.col].data_ptr(),
                                   o.values.data_ptr() if o.values is not None else None,
                                   o.bitmap.data_ptr() if o.bitmap is not None else None,
                                   o.valid_bytes.data_ptr() if o.valid_bytes is not None else None,
                                   o.null_count.data_ptr(), o.op,
                                   self.batch.columns[o.col].dec_scale if (o.op & 0xFF) >= ch.D_DEC_I64 and o.op != ch.OP_VALIDITY_ONLY else 0)
        jobs_np = np.frombuffer(bytes(jobs), dtype=np.uint8)
        jobs_dev = _dev(jobs_np, self.device)
        return outs, jobs, jobs_dev

    def run_fixed(self, plan, stream: Optional[int] = None) -> None:
        outs, jobs, jobs_dev = plan
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.dmb_dev_fixed_batch(jobs_dev.data_ptr(), C.cast(jobs, C.c_void_p), len(outs),
                                          self.counts.data_ptr(), self.row_off.data_ptr(), self.nchunks, self.nrows,
                                          C.c_void_p(stream))
        nat.check(rc, "dmb_dev_fixed_batch")

    # ---------------------------------------------------------------- strings (K5)
    def plan_string(self, col_idx: int, mode: int = 0, data_capacity: Optional[int] = None) -> StringOut:
        col = self.batch.columns[col_idx]
        assert col.phys == ch.P_STRING
        n = self.nrows
        if data_capacity is None:
            # upper bound without looking at lengths: inline bytes + heap bytes (+ terminators)
            data_capacity = 12 * n + (0 if col.heap is None else col.heap.shape[0]) + (n if mode == 2 else 0)
        off_w = 8 if mode == 1 else 4
        offsets = _empty((n + 1) * off_w, self.device)
        data = _empty(data_capacity + 32, self.device)
        total = torch.zeros(8, dtype=torch.uint8, device=self.device)
        scratch = _empty(self.lib.dmb_dev_string_scratch_bytes(self.nchunks), self.device)
        v = self.validity[col_idx]
        heap = self.heap[col_idx]
        # chunks.py pads every heap with 16 bytes: a heap of <= 16 bytes holds no string at all, and a
        # column without a heap takes the whole-vector inline kernel
        heap_len = 0 if (col.heap is None or col.heap.shape[0] <= 16) else col.heap.shape[0]
        job = nat.StringJob(self.data[col_idx].data_ptr(), v.data_ptr() if v is not None else None,
                            self.vecs[col_idx].data_ptr(), heap.data_ptr() if heap is not None else None,
                            col.heap_base, heap_len,
                            offsets.data_ptr(), data.data_ptr(), None, None, None, total.data_ptr(), mode, 0)
        return StringOut(col_idx, mode, offsets, data, total, scratch, job)

    def run_string(self, so: StringOut, stream: Optional[int] = None) -> None:
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.dmb_dev_string_batch(C.byref(so.job), self.counts.data_ptr(), self.row_off.data_ptr(),
                                           self.nchunks, self.nrows, so.scratch.data_ptr(), C.c_void_p(stream))
        nat.check(rc, "dmb_dev_string_batch")

    def string_error(self, so: StringOut) -> int:
        stream = torch.cuda.current_stream(self.device).cuda_stream
        return self.lib.dmb_dev_string_error(so.scratch.data_ptr(), C.c_void_p(stream))


def to_numpy(t: Optional[torch.Tensor], dtype=np.uint8, count: Optional[int] = None) -> Optional[np.ndarray]:
    if t is None:
        return None
    a = t.cpu().numpy()
    if count is not None:
        a = a[: count * np.dtype(dtype).itemsize]
    else:
        a = a[: a.shape[0] // np.dtype(dtype).itemsize * np.dtype(dtype).itemsize]
    return a.view(dtype)


# ---- algorithmic bytes per SURVEY.md §8(d): required DuckDB-side bytes read + Arrow-side bytes written
def algorithmic_bytes_fixed(n: int, w_in: int, w_out_bits: int) -> int:
    """w_out_bits: output bits per row (bool = 1)."""
    return n * w_in + (n * w_out_bits + 7) // 8 + 2 * ((n + 7) // 8)


def algorithmic_bytes_string(n: int, total_len: int, ptr_len: int) -> int:
    """in: 16n + sum(len>12) + n/8 ; out: 4(n+1) + sum(len) + n/8"""
    return 16 * n + ptr_len + (n + 7) // 8 + 4 * (n + 1) + total_len + (n + 7) // 8
