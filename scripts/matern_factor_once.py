"""Development aid: a few Matern factor builds at m = 20 (BASELINE config 4's shape) for ncu / timing.
python scripts/matern_factor_once.py [--n 500000] [--m 20] [--range 0.02]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nngp_b200 as nb  # noqa: E402
import bench  # noqa: E402


def arg(name, default, cast):
    return cast(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


n, m, rng_ = arg("--n", 500_000, int), arg("--m", 20, int), arg("--range", 0.02, float)
_, locs, nn, col, lm, _ = bench.build_problem(n, m, seed=1, reordering="maxmin")
for covfun, cp in (("matern_isotropic", [1.0, rng_, 0.75, 0.0]), ("exponential_isotropic", [1.0, rng_, 0.0])):
    with nb.NNGPContext(locs, nn, col, lm, covfun) as ctx:
        assert ctx.factor_build(cp) == 0
        ctx.time_op("factor_build", reps=2)
        ms, nl = ctx.time_op("factor_build", reps=10)
        print(f"n={n} m={m} {covfun} range {rng_}: factor build median {np.median(ms) * 1e3:.1f} us, min {ms.min() * 1e3:.1f} us, launches {nl}", flush=True)
