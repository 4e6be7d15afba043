"""Per-phase timeline of the string kernel's CTAs (variant library built with -DDMB_STR_TRACE).
DMB_LIB_PATH=duckdb.mbt_b200/csrc/variants/lib_trace.so python profiles/trace_string.py [--which string|string_short]"""
import argparse, ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from duckdb_mbt_b200 import devgen, native as nat  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--which", default="string"); ap.add_argument("--rows", type=int, default=16_000_000)
a = ap.parse_args()
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
db = devgen.GeneratedBatch(a.rows)
if a.which == "string": db.add_string(gen, 0.0, 10, 43)
else: db.add_string(gen, 0.0, 1, 1)
so = db.plan_string(0, 0, data_capacity=db.total_len)
for _ in range(3): db.run_string(so)
torch.cuda.synchronize()
L = C.CDLL(nat.LIB_PATH)
n = 40000 * 8
buf = np.zeros(n, dtype=np.uint64)
L.dmb_dev_string_trace(C.c_void_p(buf.ctypes.data), C.c_int64(n))
t = buf.reshape(-1, 8)[:31000].astype(np.int64)
names = ["load+rowinfo+scan", "sync_or", "offsets written+lookback", "map build", "fast pass", "slow pass"]
print("kernel span us", (t[:, 6].max() - t[:, 0].min()) / 1e3, "tiles", t.shape[0])
if a.which != "string":
    names = ["load+rowinfo+scan", "sync_or", "lookback", "", "", ""]
for k in range(6):
    d = t[:, k + 1] - t[:, k]
    if (t[:, k + 1] == 0).any() or (t[:, k] == 0).any():
        d = d[(t[:, k + 1] != 0) & (t[:, k] != 0)]
    if d.size: print(f"phase {k} {names[k]:28s} mean {d.mean()/1e3:7.2f} us  p50 {np.median(d)/1e3:7.2f}  p95 {np.percentile(d,95)/1e3:7.2f}")
life = t[:, 6] - t[:, 0]
life = life[(t[:, 6] != 0)]
if life.size: print("CTA lifetime mean us", life.mean() / 1e3, "p95", np.percentile(life, 95) / 1e3)
starts = np.sort(t[:, 0]); print("tile start rate per us", t.shape[0] / ((starts[-1] - starts[0]) / 1e3))
