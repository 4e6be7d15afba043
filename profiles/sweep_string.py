"""Time the string kernels on bench-shaped columns for one library build (DMB_LIB_PATH selects a tuning variant).
usage: python profiles/sweep_string.py [--rows N] [--shapes comment,mixed,c3,run,mode,short] [--iters K]
prints one JSON line per shape: min / median ms of K launches (CUDA events), algorithmic GB/s and fraction of the HBM peak."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from duckdb_mbt_b200 import devgen  # noqa: E402

SHAPES = {
    "comment": dict(null=0.0, lo=10, hi=43),
    "mixed": dict(null=0.0, lo=0, hi=0, len_choices=[17, 11, 4, 16]),
    "c3": dict(null=0.1, lo=0, hi=64),
    "run": dict(null=0.0, lo=13, hi=43),
    "mode": dict(null=0.0, lo=0, hi=0, len_choices=[7, 3, 4, 4, 5, 4, 3]),
    "short": dict(null=0.0, lo=1, hi=1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=60_000_000)
    ap.add_argument("--shapes", default="comment,mixed,c3")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--peak", type=float, default=6454.9)
    args = ap.parse_args()
    for name in args.shapes.split(","):
        sp = SHAPES[name]
        gen = torch.Generator(device="cuda")
        gen.manual_seed(1)
        db = devgen.GeneratedBatch(args.rows)
        kw = {"len_choices": sp["len_choices"]} if "len_choices" in sp else {}
        db.add_string(gen, sp["null"], sp["lo"], sp["hi"], **kw)
        large = db.total_len >= (1 << 31)
        so = db.plan_string(0, 1 if large else 0, data_capacity=db.total_len)
        for _ in range(3):
            db.run_string(so)
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            db.run_string(so)
            e.record()
            torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        raw = [round(x, 3) for x in ms]
        ms.sort()
        alg = db.alg_bytes_string + (4 * (args.rows + 1) if large else 0)
        out = {"lib": os.path.basename(os.environ.get("DMB_LIB_PATH", "default")), "shape": name, "rows": args.rows,
               "ms_min": round(ms[0], 4), "ms_med": round(ms[len(ms) // 2], 4)}
        if os.environ.get("SWEEP_RAW"):
            out["ms_all"] = raw
        if alg:
            out["gb_per_s"] = round(alg / ms[len(ms) // 2] / 1e6, 1)
            out["frac"] = round(alg / ms[len(ms) // 2] / 1e6 / args.peak, 4)
        print(json.dumps(out), flush=True)
        del db, so
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
