"""L0 parity of the reverse kernels (Arrow buffers -> DataChunk vectors) on DEVICE pointers that are only
element-aligned: an Arrow `offset` moves the slice start off the 16-byte grid and bitmaps may start at any byte.
Checked bit for bit against the oracle's restatement (oracle.c ora_rev_fixed / ora_rev_string; reference row path
src/duckdb_native.c:1100-1235, chunk door :2029-2132)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from duckdb_mbt_b200 import native as nat  # noqa: E402

COPY_OP = {1: 0, 2: 1, 4: 2, 8: 3, 16: 4}


def _dev(a: np.ndarray, lead: int = 0, pad: int = 64):
    """device copy of `a` placed `lead` bytes after a 256-byte aligned address, padded at the end"""
    t = torch.zeros(lead + a.nbytes + pad, dtype=torch.uint8, device="cuda:0")
    assert t.data_ptr() % 256 == 0
    t[lead: lead + a.nbytes] = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to("cuda:0")
    return t, t.data_ptr() + lead


def _run_fixed(vals_ptr, bm_ptr, bit_off, n, op, w_out):
    L = nat.lib()
    nch = (n + 2047) // 2048
    out = torch.full((nch * 2048 * w_out + 64,), 0xAB, dtype=torch.uint8, device="cuda:0")
    val = torch.full((nch * 32,), -1, dtype=torch.int64, device="cuda:0")
    nc = torch.zeros(1, dtype=torch.int64, device="cuda:0")
    jobs = (nat.RevFixedJob * 1)()
    jobs[0] = nat.RevFixedJob(vals_ptr, bm_ptr, bit_off, out.data_ptr(), val.data_ptr(), nc.data_ptr(), op, 0)
    jd = torch.from_numpy(np.frombuffer(bytes(jobs), dtype=np.uint8).copy()).to("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(L.dmb_dev_rev_fixed_batch(jd.data_ptr(), C.cast(jobs, C.c_void_p), 1, n, st), "rev_fixed")
    torch.cuda.synchronize()
    return out.cpu().numpy()[: n * w_out], val.cpu().numpy().view(np.uint64), int(nc.item())


@pytest.mark.parametrize("w", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("n,elem_off,bm_lead,with_bm", [
    (1, 1, 1, True), (15, 3, 2, True), (2048, 0, 0, True), (2049, 1, 3, True), (70_001, 3, 1, True),
    (70_001, 5, 0, False), (300_000, 7, 2, True), (1_100_003, 2, 3, True)])
def test_rev_copy_element_aligned_slices(w, n, elem_off, bm_lead, with_bm):
    rng = np.random.default_rng(1000 * w + n % 997 + elem_off)
    raw = rng.integers(0, 256, (n + elem_off) * w, dtype=np.uint8)
    bitmap = rng.integers(0, 256, (n + elem_off + 7) // 8 + 1, dtype=np.uint8) if with_bm else None
    exp_out, exp_val, exp_nc = oracle.rev_fixed(np.ascontiguousarray(raw[elem_off * w:]), bitmap, elem_off, n, COPY_OP[w], w)
    tv, pv = _dev(raw)
    pb = None
    if with_bm:
        tb, pb = _dev(bitmap, lead=bm_lead)
    got, got_val, got_nc = _run_fixed(pv + elem_off * w, pb, elem_off, n, COPY_OP[w], w)
    assert got.tobytes() == exp_out.tobytes()[: n * w]
    assert np.array_equal(got_val, exp_val)
    assert got_nc == exp_nc


@pytest.mark.parametrize("op,w_out", [(6, 8), (7, 4), (8, 2)])
@pytest.mark.parametrize("n,elem_off", [(9, 1), (5000, 0), (150_001, 3)])
def test_rev_decimal128_narrowing(op, w_out, n, elem_off):
    rng = np.random.default_rng(op * 31 + n)
    lo = rng.integers(-2**(8 * w_out - 1), 2**(8 * w_out - 1), n + elem_off, dtype=np.int64)
    raw = np.zeros((n + elem_off, 2), dtype=np.int64)
    raw[:, 0] = lo
    raw[:, 1] = lo >> 63
    raw = raw.view(np.uint8).reshape(-1)
    bitmap = rng.integers(0, 256, (n + elem_off + 7) // 8 + 1, dtype=np.uint8)
    exp_out, exp_val, exp_nc = oracle.rev_fixed(np.ascontiguousarray(raw[elem_off * 16:]), bitmap, elem_off, n, op, w_out)
    tv, pv = _dev(raw)
    tb, pb = _dev(bitmap, lead=1)
    got, got_val, got_nc = _run_fixed(pv + elem_off * 16, pb, elem_off, n, op, w_out)
    assert got.tobytes() == exp_out.tobytes()[: n * w_out]
    assert np.array_equal(got_val, exp_val) and got_nc == exp_nc


@pytest.mark.parametrize("n,bit_off,lead", [(1, 0, 0), (13, 5, 1), (2048, 7, 2), (33_333, 3, 3), (1_000_001, 6, 1)])
def test_rev_bool_bits(n, bit_off, lead):
    rng = np.random.default_rng(n + bit_off)
    vals = rng.integers(0, 256, (n + bit_off + 7) // 8 + 1, dtype=np.uint8)
    bitmap = rng.integers(0, 256, (n + bit_off + 7) // 8 + 1, dtype=np.uint8)
    exp_out, exp_val, exp_nc = oracle.rev_fixed(vals, bitmap, bit_off, n, 5, 1)
    tv, pv = _dev(vals, lead=lead)
    tb, pb = _dev(bitmap, lead=(lead + 1) % 4)
    got, got_val, got_nc = _run_fixed(pv, pb, bit_off, n, 5, 1)
    assert got.tobytes() == exp_out.tobytes()[:n]
    assert np.array_equal(got_val, exp_val) and got_nc == exp_nc


@pytest.mark.parametrize("large", [False, True])
@pytest.mark.parametrize("n,elem_off,bm_lead,with_bm,max_len", [
    (1, 0, 0, True, 5), (31, 1, 1, True, 24), (2048, 3, 2, True, 24), (2049, 0, 3, False, 40),
    (100_003, 3, 1, True, 24), (600_000, 5, 2, True, 14)])
def test_rev_string_slices(large, n, elem_off, bm_lead, with_bm, max_len):
    rng = np.random.default_rng(n + elem_off + (7 if large else 0))
    lens = rng.integers(0, max_len + 1, n + elem_off)
    offs = np.zeros(n + elem_off + 1, dtype=np.int64 if large else np.int32)
    np.cumsum(lens, out=offs[1:])
    data = rng.integers(0x20, 0x7F, int(offs[-1]) + 1, dtype=np.uint8)
    bitmap = rng.integers(0, 256, (n + elem_off + 7) // 8 + 1, dtype=np.uint8) if with_bm else None
    base = 0x7F00_0000_0000
    exp_out, exp_val, exp_nc = oracle.rev_string(np.ascontiguousarray(offs[elem_off:]), data, base, bitmap, elem_off, n)
    L = nat.lib()
    to, po = _dev(offs)
    td, pd = _dev(data, lead=1)
    pb = None
    if with_bm:
        tb, pb = _dev(bitmap, lead=bm_lead)
    nch = (n + 2047) // 2048
    out = torch.full((nch * 2048 * 16,), 0xAB, dtype=torch.uint8, device="cuda:0")
    val = torch.full((nch * 32,), -1, dtype=torch.int64, device="cuda:0")
    nc = torch.zeros(1, dtype=torch.int64, device="cuda:0")
    job = nat.RevStringJob(po + elem_off * offs.itemsize, pd, pb, elem_off, base, out.data_ptr(), val.data_ptr(),
                           nc.data_ptr(), 1 if large else 0, 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(L.dmb_dev_rev_string_batch(C.byref(job), n, st), "rev_string")
    torch.cuda.synchronize()
    assert out.cpu().numpy()[: n * 16].tobytes() == exp_out.tobytes()[: n * 16]
    assert np.array_equal(val.cpu().numpy().view(np.uint64), exp_val)
    assert int(nc.item()) == exp_nc
