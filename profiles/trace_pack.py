"""Per-role timeline of string_pack_kernel (variant library built with -DDMB_STR_TRACE):
  bash profiles/build_variant.sh trace -DDMB_STR_TRACE
  DMB_LIB_PATH=duckdb.mbt_b200/csrc/variants/lib_trace.so python profiles/trace_pack.py [--which string|string_mixed|string_short|string_c3]
Events per CTA and iteration j (worker thread 0 / lane 0 of L and P), globaltimer ns:
  W 0 top  1 S arrived  2 front done  3 at the base barrier  4 span arrived  5 pack done
  L 8 aggregate of tile j+1 seen  9 look-back done  10 pack seen  11 stores issued  12 stage drained
  P 14 tile claimed, loads issued  13 aggregate published"""
import argparse, ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from duckdb_mbt_b200 import devgen, native as nat  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--which", default="string"); ap.add_argument("--rows", type=int, default=16_000_000)
a = ap.parse_args()
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
db = devgen.GeneratedBatch(a.rows)
shape = {"string": (0.0, 10, 43, None), "string_mixed": (0.0, 0, 0, [17, 11, 4, 16]), "string_short": (0.0, 1, 1, None), "string_c3": (0.1, 0, 64, None)}[a.which]
if shape[3]: db.add_string(gen, shape[0], 0, 0, len_choices=shape[3])
else: db.add_string(gen, shape[0], shape[1], shape[2])
so = db.plan_string(0, 0, data_capacity=db.total_len)
for _ in range(3): db.run_string(so)
torch.cuda.synchronize()
L = C.CDLL(nat.LIB_PATH)
n = 32 * 64 * 16
buf = np.zeros(n, dtype=np.uint64)
L.dmb_dev_string_trace(C.c_void_p(buf.ctypes.data), C.c_int64(n))
t = buf.reshape(32, 64, 16).astype(np.int64)[:, 8:40, :]  # steady state iterations


def span(a_, b_, name, next_iter=False):
    x, y = t[:, :, a_], t[:, :, b_]
    if next_iter: x, y = x[:, :-1], y[:, 1:]
    m = (x != 0) & (y != 0)
    d = (y - x)[m]
    if d.size: print(f"{name:52s} mean {d.mean()/1e3:6.2f} us  p50 {np.median(d)/1e3:6.2f}  p90 {np.percentile(d, 90)/1e3:6.2f}")


span(0, 0, "iteration (W top -> next W top)", True)
span(0, 1, "W wait S (string_t of tile j)")
span(1, 2, "W front: lengths + scan (incl. worker barrier)")
span(2, 3, "W wait base(j-1)")
span(3, 4, "W offsets + wait stage free + wait heap span")
span(4, 5, "W pack")
span(14, 6, "P ticket + metadata round trips")
span(6, 7, "P string_t bulk load latency")
span(7, 13, "A sum lengths + publish aggregate")
span(13, 8, "aggregate published -> L starts look-back")
span(8, 15, "L look-back")
span(15, 3, "look-back(j) done -> W needs base(j) (slack)", True)
span(10, 11, "T issue stores")
span(11, 12, "T drain stage")
span(5, 10, "W(thread 0) pack done -> T sees all packed")

st = t[:, :, 9]
st = st[st != 0]
if st.size:
    print("look-back rounds  mean %.2f  p90 %d   |  re-polls mean %.2f  p90 %d  max %d" % ((st // 1000).mean(), np.percentile(st // 1000, 90), (st % 1000).mean(), np.percentile(st % 1000, 90), (st % 1000).max()))
