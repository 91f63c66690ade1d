"""Times factor-build variants and checks them against each other: python scripts/factor_time.py n m v1,v2,.. [covfun]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nngp_b200 as nb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
variants = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]
covfun = sys.argv[4] if len(sys.argv) > 4 else "exponential_isotropic"
cp = [1.0, 0.05, 0.0] if covfun.startswith("exp") else [1.0, 0.05, 0.75, 0.0]
rng = np.random.default_rng(1)
locs = rng.random((n, 2))
nn = nb.find_ordered_nn(locs, m)
col = nb.greedy_coloring(nn)
ctx = nb.NNGPContext(locs, nn, col, np.arange(1, n + 1, dtype=np.int32), covfun)
ref = None
print(f"n={n} m={m} {covfun}", flush=True)
for v in variants:
    ctx.set_option("factor_variant", v)
    ctx.factor_build(cp)
    L = ctx.factor_get()
    ctx.factor_commit()
    ctx.time_op("factor_build", reps=5)
    ms, nl = ctx.time_op("factor_build", reps=30)
    err = 0.0 if ref is None else float(np.max(np.linalg.norm(L - ref, axis=1) / np.linalg.norm(ref, axis=1)))
    if ref is None:
        ref = L
    print(f"  factor variant {v}: mean {1e3 * ms.mean():7.1f} us  min {1e3 * ms.min():7.1f} us  launches {nl}  max row rel diff vs first {err:.2e}", flush=True)
ctx.close()
