"""Runs a few sweeps of one kernel variant on config-3-like inputs (driver for ncu captures).

    python scripts/sweep_once.py [n] [variant] [sweeps]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nngp_b200 as nb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = 10
rng = np.random.default_rng(1)
locs = rng.random((n, 2))
nn = nb.find_ordered_nn(locs, m)
col = nb.greedy_coloring(nn)
ctx = nb.NNGPContext(locs, nn, col, np.arange(1, n + 1, dtype=np.int32))
ctx.factor_build([1.0, 0.05, 0.0]); ctx.factor_commit()
ctx.field_init(0.0, 0.0, rng.standard_normal(n))
ctx.obs_set(ctx.field_get() + np.sqrt(0.1) * rng.standard_normal(n))
ctx.set_option("sweep_variant", variant)
ctx.set_option("use_graph", 0)
ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), sweeps, seed=1)
print("colours", ctx.n_colors, "done")
ctx.close()
