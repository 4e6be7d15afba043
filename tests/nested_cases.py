"""DuckDB-shaped nested columns for the tests: STRUCT, LIST<VARCHAR>, MAP = LIST<STRUCT<key, value>>,
LIST<LIST<x>>, built vector by vector like libduckdb lays them out -- a LIST vector of duckdb_list_entry {offset, length}
per chunk indexing that chunk's own child vector (duckdb_list_vector_get_child / _get_size), a STRUCT vector as a validity
mask + one vector per field, VARCHAR children as duckdb_string_t with real pointers into per-vector heaps -- together
with the Python value every row stands for (what pyarrow must read back)."""
import numpy as np

from duckdb_mbt_b200 import chunks as ch


def string_entries(strings):
    """python bytes / None -> (duckdb_string_t entries uint8[n,16], heap uint8, valid bool[n]); pointers are real addresses"""
    n = len(strings)
    lens = np.asarray([0 if s is None else len(s) for s in strings], dtype=np.int64)
    heap_off = np.zeros(n, dtype=np.int64)
    pos = 0
    for i in range(n):
        if lens[i] > 12:
            heap_off[i] = pos
            pos += int(lens[i])
    heap = np.zeros(pos + 16, dtype=np.uint8)
    ent = np.zeros((n, 16), dtype=np.uint8)
    base = int(heap.ctypes.data)
    for i, s in enumerate(strings):
        if s is None:  # NULL element: unspecified payload, never followed
            ent[i, 0:4] = np.frombuffer(np.uint32(9).tobytes(), dtype=np.uint8)
            ent[i, 4:13] = np.frombuffer(b"nullslot!", dtype=np.uint8)
            continue
        L = len(s)
        ent[i, 0:4] = np.frombuffer(np.uint32(L).tobytes(), dtype=np.uint8)
        if L <= 12:
            ent[i, 4:4 + L] = np.frombuffer(s, dtype=np.uint8)
        else:
            heap[heap_off[i]: heap_off[i] + L] = np.frombuffer(s, dtype=np.uint8)
            ent[i, 4:8] = np.frombuffer(s[:4], dtype=np.uint8)
            ent[i, 8:16] = np.frombuffer(np.uint64(base + int(heap_off[i])).tobytes(), dtype=np.uint8)
    return ent, heap, np.asarray([s is not None for s in strings], dtype=bool)


def _mask_words(valid):
    words = np.zeros((len(valid) + 63) // 64 + 1, dtype=np.uint64)
    bits = np.packbits(np.asarray(valid, dtype=np.uint8), bitorder="little")
    words.view(np.uint8)[: bits.shape[0]] = bits
    return words


class ChildBuilder:
    """collects one child vector per chunk (of any size) into a Column: data_off / val_off per chunk"""

    def __init__(self, name, type_id, width, phys, dec_width=0, dec_scale=0):
        self.name, self.type_id, self.width, self.phys = name, type_id, width, phys
        self.dec_width, self.dec_scale = dec_width, dec_scale
        self.data, self.masks, self.data_off, self.val_off, self.keep = [], [], [], [], []
        self.pos = self.wpos = 0

    def add(self, raw_bytes, valid, all_valid_as_null_pointer=False):
        raw = np.ascontiguousarray(raw_bytes).view(np.uint8).reshape(-1)
        pad = (-raw.shape[0]) % 16
        self.data_off.append(self.pos)
        self.data.append(np.concatenate([raw, np.zeros(pad + 16, dtype=np.uint8)]))
        self.pos += raw.shape[0] + pad + 16
        if valid is None or (all_valid_as_null_pointer and bool(np.all(valid))):
            self.val_off.append(-1)
        else:
            w = _mask_words(valid)
            self.val_off.append(self.wpos)
            self.masks.append(w)
            self.wpos += w.shape[0]

    def column(self):
        data = np.concatenate(self.data) if self.data else np.zeros(16, dtype=np.uint8)
        validity = np.concatenate(self.masks) if self.masks else None
        col = ch.Column(self.name, self.type_id, self.phys, data, np.asarray(self.data_off, dtype=np.uint64), validity,
                        np.asarray(self.val_off, dtype=np.int64), self.dec_width, self.dec_scale)
        col._keep = self.keep
        return col


def _rebase_string_pointers(ent, heap):
    """entries built against `heap` stay valid as long as `heap` lives: nothing to do, but keep both together"""
    return ent


def _list_entries(rng, cnt, valid, lens, layout):
    """-> (starts uint64[cnt], child vector size): where each row's elements sit in the chunk's child vector"""
    if layout == "contiguous":
        eff = np.where(valid, lens, 0)
        starts = (np.cumsum(eff) - eff).astype(np.uint64)
        return starts, int(eff.sum())
    order = rng.permutation(cnt)
    starts = np.zeros(cnt, dtype=np.uint64)
    pos = int(rng.integers(0, 3))
    for i in order:
        starts[i] = pos
        pos += int(lens[i]) + int(rng.integers(0, 2))
    return starts, pos + 1


def make_list_of(kind, n, pattern, seed, layout="contiguous", null_frac=0.2, max_len=5):
    """kind: "varchar" (LIST<VARCHAR>), "map" (MAP<VARCHAR, INTEGER>), "struct" (LIST<STRUCT<a INTEGER, s VARCHAR>>),
    "list" (LIST<LIST<INTEGER>>), "list_varchar" (LIST<LIST<VARCHAR>>).  -> (Column, expected python rows)"""
    rng = np.random.default_rng(seed)
    counts = ch.chunk_counts(n, pattern, rng)
    nch = counts.shape[0]
    valid_rows = rng.random(n) >= null_frac if null_frac else np.ones(n, bool)
    lens_rows = rng.integers(0, max_len + 1, n)
    entries = np.zeros((nch, ch.VECTOR_SIZE, 2), dtype=np.uint64)
    expected = []
    sizes = []
    keep = []

    def rand_str():
        if rng.random() < 0.15:
            return None
        L = int(rng.integers(0, 30))
        return bytes(rng.integers(0x61, 0x7B, L, dtype=np.uint8))

    if kind == "varchar":
        child = ChildBuilder("item", ch.T_VARCHAR, 16, ch.P_STRING)
    elif kind in ("map", "struct"):
        f0 = ChildBuilder("key" if kind == "map" else "a", ch.T_VARCHAR if kind == "map" else ch.T_INTEGER, 16 if kind == "map" else 4,
                          ch.P_STRING if kind == "map" else ch.P_I32)
        f1 = ChildBuilder("value" if kind == "map" else "s", ch.T_INTEGER if kind == "map" else ch.T_VARCHAR, 4 if kind == "map" else 16,
                          ch.P_I32 if kind == "map" else ch.P_STRING)
        svalid = ChildBuilder("entries", ch.T_STRUCT, 1, ch.P_U8)
    else:
        inner_entries = ChildBuilder("item", ch.T_LIST, 16, ch.P_U128)
        grand = ChildBuilder("item", ch.T_VARCHAR if kind == "list_varchar" else ch.T_INTEGER, 16 if kind == "list_varchar" else 4,
                             ch.P_STRING if kind == "list_varchar" else ch.P_I32)
        grand_sizes = []
    row = 0
    for k in range(nch):
        cnt = int(counts[k])
        valid = valid_rows[row: row + cnt]
        lens = lens_rows[row: row + cnt]
        starts, size = _list_entries(rng, cnt, valid, lens, layout)
        e = entries[k]
        e[:cnt, 0], e[:cnt, 1] = starts, lens
        if (~valid).any():  # the entry of a NULL row is unspecified
            e[:cnt][~valid] = rng.integers(1 << 40, 1 << 50, (int((~valid).sum()), 2), dtype=np.uint64)
        sizes.append(size)
        if kind == "varchar":
            vals = [rand_str() for _ in range(size)]
            ent, heap, cv = string_entries(vals)
            keep.append(heap)
            child.add(ent, cv, all_valid_as_null_pointer=(k % 2 == 1))
            pyvals = [None if v is None else v.decode() for v in vals]
        elif kind == "map":
            keys = [bytes(rng.integers(0x41, 0x5B, int(rng.integers(1, 20)), dtype=np.uint8)) for _ in range(size)]
            ent, heap, _ = string_entries(keys)
            keep.append(heap)
            f0.add(ent, None)
            v = rng.integers(-1000, 1000, size).astype(np.int32)
            vv = rng.random(size) >= 0.2
            f1.add(v, vv)
            svalid.add(np.zeros(size, dtype=np.uint8), None)
            pyvals = [(keys[i].decode(), int(v[i]) if vv[i] else None) for i in range(size)]
        elif kind == "struct":
            a = rng.integers(-10**6, 10**6, size).astype(np.int32)
            av = rng.random(size) >= 0.2
            f0.add(a, av)
            ss = [rand_str() for _ in range(size)]
            ent, heap, sv = string_entries(ss)
            keep.append(heap)
            f1.add(ent, sv)
            stv = rng.random(size) >= 0.1
            svalid.add(np.zeros(size, dtype=np.uint8), stv, all_valid_as_null_pointer=True)
            pyvals = [None if not stv[i] else {"a": int(a[i]) if av[i] else None, "s": None if ss[i] is None else ss[i].decode()} for i in range(size)]
        else:
            ivalid = rng.random(size) >= 0.15
            ilens = rng.integers(0, 4, size)
            istarts, gsize = _list_entries(rng, size, ivalid, ilens, layout)
            ie = np.zeros((size, 2), dtype=np.uint64)
            ie[:, 0], ie[:, 1] = istarts, ilens
            if (~ivalid).any():
                ie[~ivalid] = rng.integers(1 << 40, 1 << 50, (int((~ivalid).sum()), 2), dtype=np.uint64)
            inner_entries.add(ie, ivalid, all_valid_as_null_pointer=True)
            grand_sizes.append(gsize)
            if kind == "list_varchar":
                gvals = [rand_str() for _ in range(gsize)]
                gent, gheap, gv = string_entries(gvals)
                keep.append(gheap)
                grand.add(gent, gv)
                gpy = [None if x is None else x.decode() for x in gvals]
            else:
                g = rng.integers(-2**31, 2**31 - 1, gsize).astype(np.int32)
                gv = rng.random(gsize) >= 0.1
                grand.add(g, gv)
                gpy = [int(g[i]) if gv[i] else None for i in range(gsize)]
            pyvals = [None if not ivalid[i] else gpy[int(istarts[i]): int(istarts[i]) + int(ilens[i])] for i in range(size)]
        for i in range(cnt):
            if not valid[i]:
                expected.append(None)
            else:
                s0, ln = int(starts[i]), int(lens[i])
                expected.append(pyvals[s0: s0 + ln])
        row += cnt
    vslab, val_off = ch.make_validity(valid_rows if null_frac else None, counts, True)
    col = ch.Column("m" if kind == "map" else "l", ch.T_MAP if kind == "map" else ch.T_LIST, ch.P_U128, entries.reshape(-1).view(np.uint8),
                    np.arange(nch, dtype=np.uint64) * np.uint64(ch.VECTOR_SIZE * 16), vslab, val_off)
    col.list_child_sizes = np.asarray(sizes, dtype=np.uint64)
    if kind == "varchar":
        col.list_child_col = child.column()
    elif kind in ("map", "struct"):
        sc = svalid.column()
        sc.struct_fields = [f0.column(), f1.column()]
        col.list_child_col = sc
    else:
        ic = inner_entries.column()
        ic.list_child_col = grand.column()
        ic.list_child_sizes = np.asarray(grand_sizes, dtype=np.uint64)
        col.list_child_col = ic
    col._keep = keep
    return counts, col, expected


def make_struct(n, pattern, seed):
    """STRUCT<i INTEGER, s VARCHAR, d DECIMAL(9,2), inner STRUCT<b BOOLEAN>> with NULL structs -> (counts, Column, expected dicts)"""
    import decimal
    rng = np.random.default_rng(seed)
    counts = ch.chunk_counts(n, pattern, rng)
    sv = rng.random(n) >= 0.15
    iv, i = rng.random(n) >= 0.2, rng.integers(-2**31, 2**31 - 1, n).astype(np.int32)
    lens = rng.integers(0, 40, n)
    strv = rng.random(n) >= 0.2
    scol = ch.string_column_bulk("s", lens, strv, counts, rng)
    svals = ch.string_values(scol, counts)
    dv, d = rng.random(n) >= 0.1, rng.integers(-10**8, 10**8, n).astype(np.int32)
    bv, b = rng.random(n) >= 0.3, rng.integers(0, 2, n).astype(np.uint8)
    inner = ch.Column("inner", ch.T_STRUCT, ch.P_U8, np.zeros(16, np.uint8), np.zeros(counts.shape[0], np.uint64), None,
                      np.full(counts.shape[0], -1, np.int64))
    inner.struct_fields = [ch.fixed_column("b", ch.T_BOOLEAN, b, counts, valid=bv)]
    vslab, val_off = ch.make_validity(sv, counts, True)
    col = ch.Column("st", ch.T_STRUCT, ch.P_U8, np.zeros(16, np.uint8), np.zeros(counts.shape[0], np.uint64), vslab, val_off)
    col.struct_fields = [ch.fixed_column("i", ch.T_INTEGER, i, counts, valid=iv, garbage_rng=rng), scol,
                         ch.fixed_column("d", ch.T_DECIMAL, d, counts, valid=dv, dec_width=9, dec_scale=2), inner]
    expected = []
    for r in range(n):
        if not sv[r]:
            expected.append(None)
            continue
        expected.append({"i": int(i[r]) if iv[r] else None, "s": None if svals[r] is None else svals[r].decode("utf-8", "replace"),
                         "d": decimal.Decimal(int(d[r])).scaleb(-2) if dv[r] else None, "inner": {"b": bool(b[r]) if bv[r] else None}})
    return counts, col, expected
