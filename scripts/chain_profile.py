"""Per-phase host wall-clock of nngp_chain_run at config 3 (NNGP_CHAIN_PROFILE=1 makes the library synchronise at phase
boundaries and print the breakdown to stderr).  Usage: python scripts/chain_profile.py [n] [m] [n_iter]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nngp_b200 as nb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rng = np.random.default_rng(1)
locs = rng.random((n, 2))
nn = nb.find_ordered_nn(locs, m)
col = nb.greedy_coloring(nn)
ctx = nb.NNGPContext(locs, nn, col, np.arange(1, n + 1, dtype=np.int32), "exponential_isotropic", device=0)
ctx.factor_build([1.0, 0.05, 0.0])
ctx.factor_commit()
ctx.field_init(0.0, 0.0, rng.standard_normal(n))
w = ctx.field_get()
y = w + np.sqrt(0.1) * rng.standard_normal(n)
ctx.obs_set(y)
p = {"shape": [np.log(0.05)], "beta_0": 0.0, "log_scale": 0.0, "log_noise_variance": float(np.log(0.1))}
var_y = float(np.var(y, ddof=1))
for prof in (False, True):
    if prof:
        os.environ["NNGP_CHAIN_PROFILE"] = "1"
    ctx.chain_run(p, 3, var_y, thin=0.0, n_chromatic=10, iter_start=5000, chain_index=1, keep_field=False)
    t0 = time.perf_counter()
    _, _, _, acc = ctx.chain_run(p, n_iter, var_y, thin=0.0, n_chromatic=10, iter_start=5000, chain_index=1, keep_field=False)
    dt = time.perf_counter() - t0
    print(f"profile={prof}: {1e6 * dt / n_iter:.1f} us / iteration, accepts (anc, suf) = {acc.sum(axis=0)}", flush=True)
ctx.close()
