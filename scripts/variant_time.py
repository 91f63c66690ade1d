"""Times sweep variants back to back on config-3-like inputs: python scripts/variant_time.py n m v1,v2,... [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nngp_b200 as nb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
variants = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "22").split(",")]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 200
rng = np.random.default_rng(1)
locs = rng.random((n, 2))
nn = nb.find_ordered_nn(locs, m)
col = nb.greedy_coloring(nn)
ctx = nb.NNGPContext(locs, nn, col, np.arange(1, n + 1, dtype=np.int32))
ctx.factor_build([1.0, 0.05, 0.0]); ctx.factor_commit()
ctx.field_init(0.0, 0.0, rng.standard_normal(n))
ctx.obs_set(ctx.field_get() + np.sqrt(0.1) * rng.standard_normal(n))
ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), 1, seed=1)
print(f"n={n} m={m} colours={ctx.n_colors}", flush=True)
for v in variants:
    ctx.set_option("sweep_variant", v)
    ctx.time_op("gibbs_sweep", reps=20)
    ms, nl = ctx.time_op("gibbs_sweep", reps=reps)
    print(f"  variant {v:2d}: mean {1e3 * ms.mean():7.1f} us  min {1e3 * ms.min():7.1f} us  launches {nl}", flush=True)
ctx.close()
