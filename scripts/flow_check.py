"""Dataflow sweep (variants 24-27) against the PDL chain (variant 22): timing and bit-for-bit equality of long runs.

    python scripts/flow_check.py [n] [m] [order]
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nngp_b200 as nb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
order = sys.argv[3] if len(sys.argv) > 3 else "random"
variants = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [22, 24, 28, 29, 30, 31, 32, 33]
rng = np.random.default_rng(1)
locs = rng.random((n, 2))
if order == "maxmin":
    locs = locs[nb.order_maxmin(locs) - 1]
nn = nb.find_ordered_nn(locs, m)
col = nb.greedy_coloring(nn)
t0 = time.perf_counter()
ctx = nb.NNGPContext(locs, nn, col, np.arange(1, n + 1, dtype=np.int32))
print(f"n={n} m={m} order={order} colours={ctx.n_colors} ctx_create {time.perf_counter() - t0:.2f} s")
ctx.factor_build([1.0, 0.05, 0.0]); ctx.factor_commit()
ctx.field_init(0.0, 0.0, rng.standard_normal(n))
ctx.obs_set(ctx.field_get() + np.sqrt(0.1) * rng.standard_normal(n))
f0 = ctx.field_get()
ns = 10
zz = rng.standard_normal(ns * n)
outs = {}
for sv in variants:
    ctx.set_option("sweep_variant", sv)
    ctx.field_set(f0)
    ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), ns, z=zz)
    outs[sv] = ctx.field_get()
    ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), 1, seed=1)
    ctx.time_op("gibbs_sweep", reps=5)
    ms, nl = ctx.time_op("gibbs_sweep", reps=50)
    ms2, _ = ctx.time_op("sweep_loglik", reps=50)
    print(f"  variant {sv}: sweep mean {ms.mean()*1e3:7.1f} us  min {ms.min()*1e3:7.1f} us  launches {nl} | sweep+loglik {ms2.mean()*1e3:7.1f} us"
          f" | {ns} sweeps max|diff vs {variants[0]}| = {np.max(np.abs(outs[sv] - outs[variants[0]])):.3e}", flush=True)
ctx.close()
