"""Kernel-level probe (device API, inputs resident in HBM): times the fixed-width and string
kernels on synthetic chunk batches generated on the device, prints algorithmic GB/s and the
fraction of the measured HBM peak.  Development tool; bench.py is the judged entry point."""
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
from duckdb_mbt_b200 import native as nat  # noqa: E402
from duckdb_mbt_b200 import devgen  # noqa: E402


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def time_ms(fn, iters=10, warmup=3, flush=None):
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
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20_000_000)
    ap.add_argument("--which", default="copy8,widen,huge,int32,bool,string")
    args = ap.parse_args()
    peak, how = peak_gbs()
    n = args.rows
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for which in args.which.split(","):
        if which == "string":
            db = devgen.string_batch(n, seed=4, null_frac=0.1, max_len=64)
            so = db.plan_string(0, 0, data_capacity=db.total_len)
            med, best = time_ms(lambda: db.run_string(so), flush=flush)
            assert db.string_error(so) == 0
            alg = db.alg_bytes_string
        else:
            spec = {"copy8": (ch.T_TIMESTAMP, 0, ch.D_SAME), "widen": (ch.T_DECIMAL, 18, ch.D_I128),
                    "huge": (ch.T_HUGEINT, 0, ch.D_SAME), "int32": (ch.T_INTEGER, 0, ch.D_SAME),
                    "bool": (ch.T_BOOLEAN, 0, ch.D_BOOL_BITS)}[which]
            db = devgen.fixed_batch(n, [(spec[0], spec[1])] * 4, null_frac=0.3, seed=5)
            plan = db.plan_fixed([(c, spec[2]) for c in range(4)], bitmap=True)
            med, best = time_ms(lambda: db.run_fixed(plan), flush=flush)
            alg = db.alg_bytes_fixed(plan)
        res[which] = {"ms_median": med, "ms_best": best, "alg_GB": alg / 1e9, "GBs": alg / 1e6 / med,
                      "frac_of_%s_peak" % how: alg / 1e6 / med / peak}
        print(which, json.dumps(res[which]), flush=True)
        del db
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
