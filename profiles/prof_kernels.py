"""Kernel-level driver for ncu (device API, inputs resident in HBM): a few launches of one kernel
family on a bench-shaped column.  `python profiles/prof_kernels.py --which string|fixed|bool|rev`."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from duckdb_mbt_b200 import chunks as ch  # noqa: E402
from duckdb_mbt_b200 import devgen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="string")
    ap.add_argument("--rows", type=int, default=16_000_000)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    n = args.rows
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    if args.which == "string":  # l_comment shape: len U[10,43], no NULLs
        db = devgen.GeneratedBatch(n)
        db.add_string(gen, 0.0, 10, 43)
        so = db.plan_string(0, 0, data_capacity=db.total_len)
        fn = lambda: db.run_string(so)  # noqa: E731
    elif args.which == "string_c3":  # C3 shape: len U[0,64], 10% NULL
        db = devgen.GeneratedBatch(n)
        db.add_string(gen, 0.1, 0, 64)
        so = db.plan_string(0, 0, data_capacity=db.total_len)
        fn = lambda: db.run_string(so)  # noqa: E731
    elif args.which == "string_mixed":  # l_shipinstruct shape: lens {17,11,4,16}: half inline, half pointer
        db = devgen.GeneratedBatch(n)
        db.add_string(gen, 0.0, 0, 0, len_choices=[17, 11, 4, 16])
        so = db.plan_string(0, 0, data_capacity=db.total_len)
        fn = lambda: db.run_string(so)  # noqa: E731
    elif args.which == "string_run":  # every row a pointer string, heap in row order: each tile is one run
        db = devgen.GeneratedBatch(n)
        db.add_string(gen, 0.0, 13, 43)
        so = db.plan_string(0, 0, data_capacity=db.total_len)
        fn = lambda: db.run_string(so)  # noqa: E731
    elif args.which == "string_mode":  # l_shipmode shape: 3..7 bytes inline
        db = devgen.GeneratedBatch(n)
        db.add_string(gen, 0.0, 0, 0, len_choices=[7, 3, 4, 4, 5, 4, 3])
        so = db.plan_string(0, 0, data_capacity=db.total_len)
        fn = lambda: db.run_string(so)  # noqa: E731
    elif args.which == "string_short":  # l_returnflag shape: 1 byte inline
        db = devgen.GeneratedBatch(n)
        db.add_string(gen, 0.0, 1, 1)
        so = db.plan_string(0, 0, data_capacity=db.total_len)
        fn = lambda: db.run_string(so)  # noqa: E731
    else:
        spec = {"fixed": (ch.T_INTEGER, 0, ch.D_SAME), "widen": (ch.T_DECIMAL, 15, ch.D_I128),
                "bool": (ch.T_BOOLEAN, 0, ch.D_BOOL_BITS), "copy8": (ch.T_TIMESTAMP, 0, ch.D_SAME)}[args.which]
        db = devgen.fixed_batch(n, [(spec[0], spec[1])] * 4, null_frac=0.0 if args.which in ("fixed", "widen") else 0.3, seed=5)
        plan = db.plan_fixed([(c, spec[2]) for c in range(4)], bitmap=True)
        fn = lambda: db.run_fixed(plan)  # noqa: E731
    for _ in range(args.iters):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    print(args.which, "ms", s.elapsed_time(e))


if __name__ == "__main__":
    main()
