"""Device-resident throughput of the other BASELINE.json configs (C1, C3, C4, C5) — bench.py covers C2.
Inputs are generated in HBM, kernels timed with CUDA events on the launching stream, L2 flushed between
iterations when the working set is small.  Prints one JSON line per config with algorithmic GB/s and the
fraction of the measured HBM peak (SURVEY.md §8d byte accounting).

    python profiles/bench_configs.py [--configs c1,c3,c4,c5] [--scale 1.0]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from duckdb_mbt_b200 import chunks as ch  # noqa: E402
from duckdb_mbt_b200 import devgen  # noqa: E402
from duckdb_mbt_b200 import native as nat  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def timeit(fn, iters=5, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


QUIET = False   # bench.py runs clist / cenum inside its own clock-sampled region and reads LINES instead of stdout
LINES = []


def report(name, rows, alg_bytes, ms, extra=None):
    pk, how = peak()
    line = {"config": name, "rows": rows, "ms": ms, "rows_per_s": rows / ms * 1e3, "alg_GB": alg_bytes / 1e9,
            "gb_per_s": alg_bytes / 1e6 / ms, f"frac_of_{how}_hbm_peak": alg_bytes / 1e6 / ms / pk}
    if extra:
        line.update(extra)
    LINES.append(line)
    if not QUIET:
        print(json.dumps(line), flush=True)


def c1(scale, flush):
    n = 1_000_000  # the reference's own CPU-runnable case: launch-latency bound, reported for completeness
    db = devgen.GeneratedBatch(n)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(20260102)
    db.add_fixed(ch.T_INTEGER, 0, gen, 0.0, "i", lo=0, hi=n)
    db.add_fixed(ch.T_DOUBLE, 0, gen, 0.0, "d")
    db.add_fixed(ch.T_INTEGER, 0, gen, 1.0, "case")  # every row NULL
    plan = db.plan_fixed([(0, ch.D_SAME), (1, ch.D_SAME), (2, ch.D_SAME)], bitmap=True)
    ms = timeit(lambda: db.run_fixed(plan), flush=flush)
    report("C1 range(1000000) 3 cols", n, db.alg_bytes_fixed(plan), ms, {"note": "32.75 MB per pass: launch-latency bound"})


def c3(scale, flush):
    n = int(100_000_000 * scale)
    db = devgen.string_batch(n, seed=20260104, null_frac=0.10, max_len=64)
    # 100M rows carry ~2.9 GB of utf8 data: int32 offsets would overflow, so the large_utf8 mode is timed
    mode = 1 if db.total_len > 2**31 - 1 else 0
    so = db.plan_string(0, mode, data_capacity=db.total_len)
    val = db.plan_fixed([(0, ch.OP_VALIDITY_ONLY)], bitmap=True)

    def run():
        db.run_fixed(val)
        db.run_string(so)
    ms = timeit(run)
    assert db.string_error(so) == 0
    m = db.meta[0]
    alg = 16 * n + m["ptr_len"] + (8 if mode else 4) * (n + 1) + m["total_len"] + 2 * ((n + 7) // 8)
    report("C3 VARCHAR-heavy len U[0,64] 10% NULL", n, alg, ms, {"offsets": "int64" if mode else "int32", "utf8_bytes": m["total_len"]})


def c4(scale, flush):
    n = int(10_000_000 * scale)
    cols = [(ch.T_TIMESTAMP, 0)] * 22 + [(ch.T_DECIMAL, 18)] * 21 + [(ch.T_HUGEINT, 0)] * 21
    db = devgen.fixed_batch(n, cols, null_frac=0.30, seed=20260105)
    specs = [(j, ch.D_I128 if t == ch.T_DECIMAL else (ch.D_I128 if t == ch.T_HUGEINT else ch.D_SAME)) for j, (t, _) in enumerate(cols)]
    plan = db.plan_fixed(specs, bitmap=True)
    ms = timeit(lambda: db.run_fixed(plan))
    report("C4 wide 64 cols (22 TIMESTAMP, 21 DECIMAL(18,3)->decimal128, 21 HUGEINT), 30% NULL", n, db.alg_bytes_fixed(plan), ms,
           {"launches": len({o.op for o in plan[0]})})


def c5(scale, flush):
    """Reverse path: Arrow int32 id, int64 v, float64 x, bool flag, utf8 s (len U[0,24]); 10% NULL except id;
    slices start at a non-zero bit offset.  Processed in 50M-row device-resident batches."""
    L = nat.lib()
    n = int(50_000_000 * scale)
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(20260106)
    off = 3  # array offset (bits) of every sliced column
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nch = (n + 2047) // 2048

    def bitmap(frac):
        bits = torch.rand(n + off + 64, generator=gen, device=dev) >= frac
        w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=dev)
        return (bits[: (n + off + 64) // 8 * 8].view(-1, 8).to(torch.int32) * w).sum(dim=1).to(torch.uint8)

    ids = torch.arange(n + off, dtype=torch.int32, device=dev)
    v = torch.randint(-2**62, 2**62, (n + off,), generator=gen, device=dev, dtype=torch.int64)
    x = torch.rand(n + off, generator=gen, device=dev, dtype=torch.float64)
    flag = bitmap(0.5)
    vm, xm, fm, sm_ = bitmap(0.1), bitmap(0.1), bitmap(0.1), bitmap(0.1)
    lens = torch.randint(0, 25, (n + off,), generator=gen, device=dev, dtype=torch.int64)
    offs64 = torch.zeros(n + off + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=offs64[1:])
    total = int(offs64[-1].item())
    offs = offs64.to(torch.int32)
    data = torch.randint(0x20, 0x7F, (total + 64,), generator=gen, device=dev, dtype=torch.uint8)
    outs = {w: torch.empty(nch * 2048 * w + 64, dtype=torch.uint8, device=dev) for w in (1, 4, 8)}
    out_x = torch.empty(nch * 2048 * 8 + 64, dtype=torch.uint8, device=dev)
    out_s = torch.empty(nch * 2048 * 16 + 64, dtype=torch.uint8, device=dev)
    vals = [torch.empty(nch * 32 * 8 + 64, dtype=torch.uint8, device=dev) for _ in range(5)]
    jobs = (nat.RevFixedJob * 4)()
    jobs[0] = nat.RevFixedJob(ids.data_ptr() + 4 * off, None, off, outs[4].data_ptr(), vals[0].data_ptr(), None, 2, 0)       # COPY4
    jobs[1] = nat.RevFixedJob(v.data_ptr() + 8 * off, vm.data_ptr(), off, outs[8].data_ptr(), vals[1].data_ptr(), None, 3, 0)  # COPY8
    jobs[2] = nat.RevFixedJob(x.data_ptr() + 8 * off, xm.data_ptr(), off, out_x.data_ptr(), vals[2].data_ptr(), None, 3, 0)
    jobs[3] = nat.RevFixedJob(flag.data_ptr(), fm.data_ptr(), off, outs[1].data_ptr(), vals[3].data_ptr(), None, 5, 0)        # BITS_TO_BOOL
    jobs_dev = torch.from_numpy(np.frombuffer(bytes(jobs), dtype=np.uint8).copy()).to(dev)
    sjob = nat.RevStringJob(offs.data_ptr() + 4 * off, data.data_ptr(), sm_.data_ptr(), off, 0x7F0000000000, out_s.data_ptr(),
                            vals[4].data_ptr(), None, 0, 0)

    def run():
        nat.check(L.dmb_dev_rev_fixed_batch(jobs_dev.data_ptr(), C.cast(jobs, C.c_void_p), 4, n, stream), "rev_fixed")
        nat.check(L.dmb_dev_rev_string_batch(C.byref(sjob), n, stream), "rev_string")
    ms = timeit(run)
    ms_fixed = timeit(lambda: nat.check(L.dmb_dev_rev_fixed_batch(jobs_dev.data_ptr(), C.cast(jobs, C.c_void_p), 4, n, stream), "rev_fixed"))
    ms_string = timeit(lambda: nat.check(L.dmb_dev_rev_string_batch(C.byref(sjob), n, stream), "rev_string"))
    live = int((offs64[n + off] - offs64[off]).item())
    # SURVEY.md §8d reverse accounting: Arrow buffers read once, vectors + masks written; string bytes are read
    # for the prefix / inline fill only (pointer strings refer to the Arrow data buffer in place)
    alg_s = 4 * (n + 1) + live + 16 * n + 2 * (n // 8)
    alg = n * (4 + 4) + n * (8 + 8) * 2 + (n // 8 + n) + 4 * (n // 8) * 2 + n // 8 + alg_s
    report("C5 appender reverse: Arrow (int32,int64,float64,bool,utf8 U[0,24]) -> DataChunk vectors, 10% NULL, bit offset 3",
           n, alg, ms, {"batch": "50M-row device-resident batch (a 500M-row table is 10 such batches per GPU)",
                        "rev_fixed_kernel": {"ms": ms_fixed, "gb_per_s": (alg - alg_s) / 1e6 / ms_fixed},
                        "rev_string_kernel": {"ms": ms_string, "gb_per_s": alg_s / 1e6 / ms_string}})


def clist(scale, flush):
    """LIST<INTEGER> column: len U[0,6], 10 % NULL rows, 10 % NULL elements, entries contiguous in row
    order (what a scan produces) -> Arrow list<int32>: offsets + gathered child + child bitmap (kernels_list.cu)."""
    L = nat.lib()
    n = int(20_000_000 * scale)
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(20260107)
    nch = (n + 2047) // 2048
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    valid = torch.rand(nch * 2048, generator=gen, device=dev) >= 0.1
    lens = torch.randint(0, 7, (nch * 2048,), generator=gen, device=dev, dtype=torch.int64)
    lens[n:] = 0
    eff = torch.where(valid, lens, torch.zeros_like(lens))
    pref = torch.cumsum(eff, 0) - eff
    total = int(eff.sum().item())
    child_base = pref.view(nch, 2048)[:, 0].contiguous()                      # element index of each chunk's child vector
    ent = torch.stack([pref - child_base.repeat_interleave(2048), lens], dim=1).contiguous()   # {offset, length} per row
    w = torch.tensor([1 << i for i in range(8)], dtype=torch.int32, device=dev)
    vmask = (valid.view(-1, 8).to(torch.int32) * w).sum(dim=1).to(torch.uint8).contiguous()   # 32 uint64 words per chunk
    child = torch.randint(-2**31, 2**31 - 1, (total + 64,), generator=gen, device=dev, dtype=torch.int32)
    cvalid = torch.rand((total + 64 + 7) // 8 * 8, generator=gen, device=dev) >= 0.1
    cmask = (cvalid.view(-1, 8).to(torch.int32) * w).sum(dim=1).to(torch.uint8)
    # per-chunk child masks start at a word boundary in DuckDB; here chunk k's mask is staged at its own word offset
    # by re-packing: simplest faithful staging for a device-side generator is all-valid masks per chunk except a shared one
    # -> give every chunk the NULL mask pointer except via one staged mask per chunk computed from the global bits
    counts = torch.full((nch,), 2048, dtype=torch.int32, device=dev)
    counts[-1] = n - (nch - 1) * 2048
    row_off = torch.arange(nch + 1, dtype=torch.int64, device=dev) * 2048
    row_off[-1] = n
    vecs = torch.stack([torch.arange(nch, dtype=torch.int64, device=dev) * (2048 * 16), torch.arange(nch, dtype=torch.int64, device=dev) * 32], dim=1).contiguous()
    # child masks: chunk k's bits = global bits [child_base[k], child_base[k+1]) shifted to bit 0 of its own words (host-side
    # staging does exactly this copy); built here with a gather over bit indices
    sizes = torch.diff(torch.cat([child_base, torch.tensor([total], device=dev)]))
    words = (sizes + 63) // 64 + 1
    val_off = torch.cumsum(words, 0) - words
    nbits = int(words.sum().item()) * 64
    bit_chunk = torch.repeat_interleave(torch.arange(nch, device=dev), words * 64)
    local = torch.arange(nbits, device=dev) - (val_off * 64)[bit_chunk]
    src = (child_base[bit_chunk] + local).clamp_(max=cvalid.numel() - 1)
    bits = cvalid[src] & (local < sizes[bit_chunk])
    cmask_staged = (bits.view(-1, 8).to(torch.int32) * w).sum(dim=1).to(torch.uint8).contiguous()
    del bit_chunk, local, src, bits
    out_off = torch.empty(4 * (n + 1) + 64, dtype=torch.uint8, device=dev)
    out_child = torch.empty(4 * total + 64, dtype=torch.uint8, device=dev)
    out_bm = torch.empty(((total + 63) // 64 + 1) * 8, dtype=torch.uint8, device=dev)
    ctr = torch.zeros(2, dtype=torch.int64, device=dev)
    scratch = torch.zeros(L.dmb_dev_list_scratch_bytes(nch) + 64, dtype=torch.uint8, device=dev)
    job = nat.ListJob(ent.data_ptr(), vmask.data_ptr(), vecs.data_ptr(), child_base.data_ptr(), child.data_ptr(), cmask_staged.data_ptr(),
                      val_off.data_ptr(), out_off.data_ptr(), out_child.data_ptr(), out_bm.data_ptr(), ctr.data_ptr(), ctr.data_ptr() + 8, 4, 0)

    def run():
        nat.check(L.dmb_dev_list_batch(C.byref(job), counts.data_ptr(), row_off.data_ptr(), nch, n, total, scratch.data_ptr(), stream), "list")
    ms = timeit(run)
    assert int(ctr[0].item()) == total and int(scratch[:8].view(torch.int64)[0].item()) == 0
    alg = 16 * n + n // 8 + 4 * total + total // 8 + 4 * (n + 1) + 4 * total + total // 8
    report("LIST<INTEGER> len U[0,6], 10% NULL rows / elements, contiguous entries -> Arrow list<int32>", n, alg, ms,
           {"child_elements": total, "launches": 1})


def cenum(scale, flush):
    """ENUM column: uint8 indices over the 7 l_shipmode labels, no NULLs -> utf8 offsets + data: the fused
    enum_pack_kernel (dmb_dev_enum_utf8), checked bit for bit against the two-step path (enum_to_string_t_kernel: indices ->
    string_t into the dictionary, + the heap-less string kernel), which is timed beside it."""
    L = nat.lib()
    n = int(60_000_000 * scale)
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(20260108)
    nch = (n + 2047) // 2048
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    labels = [b"REG AIR", b"AIR", b"RAIL", b"SHIP", b"TRUCK", b"MAIL", b"FOB"]
    d_offs, d_data = ch.enum_dict_arrays(labels)
    t_offs = torch.from_numpy(d_offs.view(np.uint8).copy()).to(dev)
    t_data = torch.from_numpy(np.concatenate([d_data, np.zeros(64, np.uint8)])).to(dev)
    idx = torch.randint(0, len(labels), (nch * 2048,), generator=gen, device=dev, dtype=torch.uint8)
    counts = torch.full((nch,), 2048, dtype=torch.int32, device=dev)
    counts[-1] = n - (nch - 1) * 2048
    row_off = torch.arange(nch + 1, dtype=torch.int64, device=dev) * 2048
    row_off[-1] = n
    vecs_in = torch.stack([torch.arange(nch, dtype=torch.int64, device=dev) * 2048, torch.full((nch,), -1, dtype=torch.int64, device=dev)], dim=1).contiguous()
    vecs_str = torch.stack([torch.arange(nch, dtype=torch.int64, device=dev) * (2048 * 16), torch.full((nch,), -1, dtype=torch.int64, device=dev)], dim=1).contiguous()
    str_t = torch.empty(nch * 2048 * 16 + 64, dtype=torch.uint8, device=dev)
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    ejob = nat.EnumJob(idx.data_ptr(), None, vecs_in.data_ptr(), str_t.data_ptr(), t_offs.data_ptr(), t_data.data_ptr(), 1 << 41,
                       bad.data_ptr(), len(labels), ch.P_U8)
    total_len = int(torch.tensor([len(x) for x in labels], device=dev)[idx[:n].long()].sum().item())
    offsets = torch.empty(4 * (n + 1) + 64, dtype=torch.uint8, device=dev)
    data = torch.empty(total_len + 64, dtype=torch.uint8, device=dev)
    total = torch.zeros(8, dtype=torch.uint8, device=dev)
    scratch = torch.empty(L.dmb_dev_string_scratch_bytes(nch), dtype=torch.uint8, device=dev)
    sjob = nat.StringJob(str_t.data_ptr(), None, vecs_str.data_ptr(), None, 1 << 41, 0, offsets.data_ptr(), data.data_ptr(), None, None, None,
                         total.data_ptr(), 0, 0)

    def lookup():
        nat.check(L.dmb_dev_enum_to_string_t(C.byref(ejob), counts.data_ptr(), nch, stream), "enum")

    def pack():
        total.zero_()
        nat.check(L.dmb_dev_string_batch(C.byref(sjob), counts.data_ptr(), row_off.data_ptr(), nch, n, scratch.data_ptr(), stream), "string")

    def run():
        lookup()
        pack()
    ms2 = timeit(run)
    ms_lookup, ms_pack = timeit(lookup), timeit(pack)
    assert int(bad.item()) == 0 and int(total.view(torch.int64)[0].item()) == total_len
    two_step = (offsets[:4 * (n + 1)].clone(), data[:total_len].clone())
    offsets.zero_()
    data.zero_()

    def fused():
        total.zero_()
        nat.check(L.dmb_dev_enum_utf8(C.byref(ejob), C.byref(sjob), counts.data_ptr(), row_off.data_ptr(), nch, n, scratch.data_ptr(), stream), "enum_utf8")
    ms = timeit(fused)
    assert int(bad.item()) == 0 and int(total.view(torch.int64)[0].item()) == total_len
    assert torch.equal(offsets[:4 * (n + 1)], two_step[0]) and torch.equal(data[:total_len], two_step[1])  # bit-identical to the two-step form
    alg = n * 1 + 4 * (n + 1) + total_len  # indices in, offsets + label bytes out (no validity: all valid)
    report("ENUM(7 labels) uint8 indices, no NULLs -> utf8, one launch (dmb_dev_enum_utf8: enum_pack_kernel, the label table in shared memory)", n, alg, ms,
           {"two_step_ms": ms2, "enum_to_string_t_kernel_ms": ms_lookup, "heap_less_string_kernel_ms": ms_pack,
            "note": "two_step = lookup kernel + heap-less string kernel through a 16-byte string_t per row (32 B/row of traffic that is not algorithmic)"})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c3,c4,c5")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
    for name in args.configs.split(","):
        {"c1": c1, "c3": c3, "c4": c4, "c5": c5, "list": clist, "enum": cenum}[name](args.scale, flush)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
