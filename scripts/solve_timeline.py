"""Development aid: where the sync-free triangular solve spends its time, per DAG level (nngp_solve_timeline).
python scripts/solve_timeline.py [--m 10]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nngp_b200 as nb  # noqa: E402
from nngp_b200 import _lib as L  # noqa: E402
import bench  # noqa: E402

m = int(sys.argv[sys.argv.index("--m") + 1]) if "--m" in sys.argv else 10
n = 1_000_000
_, locs, nn, col, lm, _ = bench.build_problem(n, m, seed=1, reordering="maxmin")
ctx = nb.NNGPContext(locs, nn, col, lm)
assert ctx.factor_build([1.0, 0.05, 0.0]) == 0
ctx.factor_commit()
ctx.field_init(0.0, 0.0, np.random.default_rng(1).standard_normal(n))
ctx.time_op("sptrsv", reps=3)
cap = 8192
for rep in range(2):
    out, lev, nc, st = np.zeros(cap), np.zeros(cap, dtype=np.int32), C.c_int(cap), C.c_int(0)
    L.load().nngp_solve_timeline(L.ci(ctx._id), L.dptr(out), L.iptr(lev), C.byref(nc), C.byref(st))
    L.check(st)
k = nc.value
out, lev = out[:k], lev[:k]
print(f"n={n} m={m}: {k} chunks, {ctx.n_levels} levels, last chunk done at {out.max() / 1e3:.1f} us")
# time at which the last chunk STARTING in each level finished, and the level's width in chunks
lv_end = {}
for c in range(k):
    lv_end[lev[c]] = max(lv_end.get(lev[c], 0.0), out[c])
levels = sorted(lv_end)
prev = 0.0
print("level  chunks  done_us  delta_us")
for i, l in enumerate(levels):
    cnt = int((lev == l).sum())
    if i % 8 == 0 or cnt > 8:
        print(f"{l:5d} {cnt:7d} {lv_end[l] / 1e3:8.1f} {(lv_end[l] - prev) / 1e3:8.2f}")
    prev = lv_end[l]
ctx.close()
