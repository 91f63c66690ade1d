"""Prints the per-stage timeline of CTA 0 of the persistent sweep kernel (development aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nngp_b200 as nb
from nngp_b200.context import debug_timeline

n, m = 1_000_000, 10
rng = np.random.default_rng(1)
locs = rng.random((n, 2))
nn = nb.find_ordered_nn(locs, m)
col = nb.greedy_coloring(nn)
ctx = nb.NNGPContext(locs, nn, col, np.arange(1, n + 1, dtype=np.int32))
ctx.factor_build([1.0, 0.05, 0.0]); ctx.factor_commit()
ctx.field_init(0.0, 0.0, rng.standard_normal(n))
ctx.obs_set(ctx.field_get() + np.sqrt(0.1) * rng.standard_normal(n))
for variant in (0, 5, 4):
    ctx.set_option("sweep_variant", variant)
    ctx.set_option("debug_timeline", 1)
    ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), 2, seed=1)
    ms, _ = ctx.time_op("gibbs_sweep", reps=3)
    t = debug_timeline()
    print(f"variant {variant}: {len(t)} stamps, sweep {ms.mean()*1e3:.1f} us")
    names = {0: "start", 1: "gathered", 2: "summed", 3: "arrived", 4: "staged"}
    t0 = t[0, 0]
    last = t0
    durs = {}
    for ts, st in t:
        durs.setdefault(int(st), []).append(ts - last)
        last = ts
    for st in sorted(durs):
        d = np.array(durs[st])
        print(f"   -> {names[st]:9s}: mean {d.mean():8.0f} ns  median {np.median(d):8.0f}  max {d.max():8.0f}  (n={d.size})")
    print("   first 30 stamps (us since start):", [(round((ts - t0) / 1e3, 1), int(st)) for ts, st in t[:30]])
