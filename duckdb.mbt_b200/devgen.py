"""Synthetic DuckDB-shaped chunk batches generated directly in HBM (bench scale).

Same layouts as chunks.py (2048-row vector slots, uint64[32] validity per vector, 16-byte
duckdb_string_t with inline / prefix+pointer forms into a contiguous heap), but built with torch
ops on the device so that the 60M-100M row configs of BASELINE.json do not need minutes of numpy.
torch is plumbing here (RNG + allocations); the string_t entries are assembled by the library's
own helper kernel (dmb_dev_make_string_t).  Deterministic per seed.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import chunks as ch
from . import native as nat
from .device import DeviceBatch

VS = ch.VECTOR_SIZE


class GeneratedBatch(DeviceBatch):
    """DeviceBatch whose slabs were created on the device (no host ChunkBatch behind it)."""

    def __init__(self, nrows: int, device="cuda:0"):
        self.lib = nat.lib()
        self.device = torch.device(device)
        self.nrows = int(nrows)
        self.nchunks = (self.nrows + VS - 1) // VS
        counts = torch.full((self.nchunks,), VS, dtype=torch.int32, device=self.device)
        if self.nrows % VS:
            counts[-1] = self.nrows % VS
        self.counts = counts.view(torch.uint8)
        row_off = torch.arange(self.nchunks + 1, dtype=torch.int64, device=self.device) * VS
        row_off[-1] = self.nrows
        self.row_off = row_off.view(torch.uint8)
        self.data: List[torch.Tensor] = []
        self.validity: List[Optional[torch.Tensor]] = []
        self.vecs: List[torch.Tensor] = []
        self.heap: List[Optional[torch.Tensor]] = []
        self.batch = SimpleNamespace(columns=[])
        self.meta = []

    @property
    def capacity(self) -> int:
        return self.nchunks * VS

    def _vecs(self, width: int, has_validity: bool) -> torch.Tensor:
        k = torch.arange(self.nchunks, dtype=torch.int64, device=self.device)
        d = torch.empty((self.nchunks, 2), dtype=torch.int64, device=self.device)
        d[:, 0] = k * (VS * width)
        d[:, 1] = k * ch.VALIDITY_WORDS if has_validity else -1
        return d.view(torch.uint8).reshape(-1)

    def _validity(self, gen: torch.Generator, null_frac: float):
        """-> (uint8 tensor of packed masks (capacity/8 bytes), bool valid[nrows])"""
        cap = self.capacity
        valid = torch.rand(cap, generator=gen, device=self.device) >= null_frac
        valid[self.nrows:] = False
        w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=self.device)
        packed = (valid.view(-1, 8).to(torch.int32) * w).sum(dim=1).to(torch.uint8)
        return packed, valid[: self.nrows]

    def add_fixed(self, type_id: int, dec_width: int, gen: torch.Generator, null_frac: float, name: str = "c",
                  lo: Optional[int] = None, hi: Optional[int] = None, dec_scale: int = 0):
        """lo/hi: uniform integer payload in [lo, hi) (valid rows and NULL slots alike); default: random bits."""
        phys = ch.phys_of_type(type_id, dec_width)
        width = ch.PHYS_WIDTH[phys]
        nbytes = self.capacity * width
        if phys == ch.P_BOOL:
            data = torch.randint(0, 2, (nbytes,), generator=gen, device=self.device, dtype=torch.uint8)
        elif lo is not None and width in (4, 8):
            dt = torch.int32 if width == 4 else torch.int64
            data = torch.randint(lo, hi, (self.capacity,), generator=gen, device=self.device, dtype=dt).view(torch.uint8)
        else:
            # random payload everywhere, including under NULLs (garbage that must be zeroed on output)
            data = torch.randint(-2**63, 2**63 - 1, ((nbytes + 7) // 8,), generator=gen, device=self.device,
                                 dtype=torch.int64).view(torch.uint8)[:nbytes]
        packed, valid = (None, None)
        if null_frac > 0:
            packed, valid = self._validity(gen, null_frac)
        self.data.append(data)
        self.validity.append(packed)
        self.vecs.append(self._vecs(width, packed is not None))
        self.heap.append(None)
        col = SimpleNamespace(name=name, type_id=type_id, phys=phys, dec_width=dec_width, dec_scale=dec_scale,
                              heap=None, heap_base=0, width=width)
        self.batch.columns.append(col)
        self.meta.append({"valid": valid})
        return len(self.data) - 1

    def add_string(self, gen: torch.Generator, null_frac: float, min_len: int, max_len: int,
                   host_base: int = 0x7F0000000000, name: str = "s", len_choices: Optional[Sequence[int]] = None,
                   host_heap_alloc=None):
        """len_choices: lengths drawn uniformly from this list (dictionary-like columns) instead of
        U[min_len, max_len].  host_heap_alloc(nbytes) -> uint8 numpy array: the string_t pointers then
        are real addresses into that (page-locked) host array, which to_host_batch fills."""
        n = self.nrows
        if len_choices is not None:
            table = torch.tensor(list(len_choices), dtype=torch.int64, device=self.device)
            lens = table[torch.randint(0, len(len_choices), (n,), generator=gen, device=self.device)]
        else:
            lens = torch.randint(min_len, max_len + 1, (n,), generator=gen, device=self.device, dtype=torch.int64)
        packed, valid = (None, None)
        if null_frac > 0:
            packed, valid = self._validity(gen, null_frac)
        is_ptr = lens > 12
        heap_lens = torch.where(is_ptr, lens, torch.zeros_like(lens))
        heap_off = torch.cumsum(heap_lens, 0) - heap_lens
        heap_total = int(heap_lens.sum().item())
        heap = torch.randint(0x20, 0x7F, (heap_total + 64,), generator=gen, device=self.device, dtype=torch.uint8)
        host_heap = None
        if host_heap_alloc is not None:
            host_heap = host_heap_alloc(heap_total + 64)
            host_base = int(host_heap.ctypes.data)
        # inline rows take their bytes from an arbitrary heap position (they own no heap storage)
        idx = torch.arange(n, dtype=torch.int64, device=self.device)
        inline_src = (idx * 13) % max(heap_total, 1)
        src_off = torch.where(is_ptr, heap_off, inline_src)
        entries = torch.zeros(self.capacity * 16, dtype=torch.uint8, device=self.device)
        lens32 = lens.to(torch.int32)
        rc = self.lib.dmb_dev_make_string_t(lens32.data_ptr(), src_off.data_ptr(), heap.data_ptr(), host_base,
                                            entries.data_ptr(), n,
                                            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        nat.check(rc, "dmb_dev_make_string_t")
        torch.cuda.synchronize(self.device)
        self.data.append(entries)
        self.validity.append(packed)
        self.vecs.append(self._vecs(16, packed is not None))
        self.heap.append(heap)
        heap_ns = SimpleNamespace(shape=(heap_total,))
        col = SimpleNamespace(name=name, type_id=ch.T_VARCHAR, phys=ch.P_STRING, dec_width=0, dec_scale=0,
                              heap=heap_ns, heap_base=host_base, width=16)
        self.batch.columns.append(col)
        live = lens if valid is None else torch.where(valid, lens, torch.zeros_like(lens))
        live_ptr = torch.where(is_ptr, live, torch.zeros_like(live))
        self.meta.append({"valid": valid, "total_len": int(live.sum().item()), "ptr_len": int(live_ptr.sum().item()),
                          "heap_total": heap_total, "host_heap": host_heap})
        return len(self.data) - 1

    # ---- host copy (bench e2e leg, L1 tests): same bytes in (page-locked) host slabs
    def to_host_batch(self, alloc=None) -> ch.ChunkBatch:
        """alloc(nbytes) -> uint8 numpy array (e.g. pinned.pinned_empty); default: pageable numpy."""
        alloc = alloc or (lambda nb: np.empty(max(int(nb), 1), dtype=np.uint8)[: int(nb)])

        def to_host(t: torch.Tensor) -> np.ndarray:
            out = alloc(t.numel())
            torch.from_numpy(out).copy_(t.view(torch.uint8).reshape(-1))
            return out

        counts = self.counts.view(torch.int32).cpu().numpy().astype(np.uint32)
        cols = []
        for j, c in enumerate(self.batch.columns):
            data = to_host(self.data[j])
            validity = None if self.validity[j] is None else to_host(self.validity[j]).view(np.uint64)
            vecs = self.vecs[j].view(torch.int64).reshape(-1, 2).cpu().numpy()
            heap = None
            if c.phys == ch.P_STRING:
                host_heap = self.meta[j].get("host_heap")
                if host_heap is None:
                    raise ValueError("string column was generated without host_heap_alloc: its pointers are not host addresses")
                nb = self.heap[j].numel()
                torch.from_numpy(host_heap[:nb]).copy_(self.heap[j])
                # a vector whose strings are all inlined owns no string heap (DuckDB allocates none): register none
                heap = host_heap[: self.meta[j]["heap_total"] + 16] if self.meta[j]["heap_total"] > 0 else None
            cols.append(ch.Column(c.name, c.type_id, c.phys, data, vecs[:, 0].astype(np.uint64).copy(), validity,
                                  vecs[:, 1].copy(), c.dec_width, c.dec_scale, heap,
                                  inline_only=c.phys == ch.P_STRING and heap is None))
        torch.cuda.synchronize(self.device)
        return ch.ChunkBatch(counts, cols)

    def window_to_host(self, c0: int, c1: int) -> ch.ChunkBatch:
        """Chunks [c0, c1) as a host ChunkBatch (pageable numpy): what a parity check hands to the CPU oracle.  Pointer
        strings are re-pointed at a host copy of the window's heap span, so the oracle dereferences real addresses."""
        k = c1 - c0
        counts = self.counts.view(torch.int32)[c0:c1].cpu().numpy().astype(np.uint32)
        cols = []
        for j, c in enumerate(self.batch.columns):
            w = c.width
            data = self.data[j][c0 * VS * w: c1 * VS * w].cpu().numpy().copy()
            validity, val_off = None, np.full(k, -1, dtype=np.int64)
            if self.validity[j] is not None:
                validity = self.validity[j][c0 * 256: c1 * 256].cpu().numpy().copy().view(np.uint64)
                val_off = np.arange(k, dtype=np.int64) * ch.VALIDITY_WORDS
            heap = None
            if c.phys == ch.P_STRING:
                ent = data.reshape(-1, 16)
                lens = ent[:, 0:4].copy().view(np.uint32).reshape(-1)
                ptrs = ent[:, 8:16].copy().view(np.uint64).reshape(-1)
                isp = lens > 12
                if isp.any():
                    base = np.uint64(c.heap_base)
                    lo = int(ptrs[isp].min() - base)
                    hi = int((ptrs[isp] + lens[isp].astype(np.uint64)).max() - base)
                    heap = np.zeros(hi - lo + 16, dtype=np.uint8)
                    heap[: hi - lo] = self.heap[j][lo:hi].cpu().numpy()
                    new = ptrs[isp] - (base + np.uint64(lo)) + np.uint64(heap.ctypes.data)
                    ent[isp, 8:16] = new.view(np.uint8).reshape(-1, 8)
            cols.append(ch.Column(c.name, c.type_id, c.phys, data, np.arange(k, dtype=np.uint64) * np.uint64(VS * w), validity, val_off,
                                  c.dec_width, c.dec_scale, heap))
        return ch.ChunkBatch(counts, cols)

    # ---- algorithmic bytes
    def alg_bytes_fixed(self, plan) -> int:
        total = 0
        n = self.nrows
        for o in plan[0]:
            col = self.batch.columns[o.col]
            if o.op == ch.OP_VALIDITY_ONLY:
                total += 2 * ((n + 7) // 8)
                continue
            out_bits = 1 if o.width == 0 else 8 * o.width
            total += n * col.width + (n * out_bits + 7) // 8 + 2 * ((n + 7) // 8)
        return total

    @property
    def total_len(self) -> int:
        return sum(m.get("total_len", 0) for m in self.meta)

    @property
    def alg_bytes_string(self) -> int:
        total = 0
        n = self.nrows
        for m in self.meta:
            if "total_len" in m:
                total += 16 * n + m["ptr_len"] + 4 * (n + 1) + m["total_len"]
        return total


def fixed_batch(nrows: int, cols: Sequence[Tuple[int, int]], null_frac: float, seed: int, device="cuda:0") -> GeneratedBatch:
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    b = GeneratedBatch(nrows, device)
    for j, (type_id, dec_width) in enumerate(cols):
        b.add_fixed(type_id, dec_width, gen, null_frac, name=f"c{j}")
    return b


def string_batch(nrows: int, seed: int, null_frac: float, max_len: int, min_len: int = 0, device="cuda:0") -> GeneratedBatch:
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    b = GeneratedBatch(nrows, device)
    b.add_string(gen, null_frac, min_len, max_len)
    return b
