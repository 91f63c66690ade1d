"""Replays the reference's vignette session on the CPU oracle (R's own random stream) and prints what the vignette prints.
Used once to establish how far the printed values are reproduced; the pinned subset is tests/test_vignette_pin.py."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from oracle import reference_driver as R


def toy():
    O.set_seed(1)
    locs = np.column_stack([500.0 * O.runif(2000), np.ones(2000)])
    locs[0, 1] = 1.01
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    field = np.sqrt(10.0) * (np.linalg.cholesky(np.exp(-D / 5.0)) @ O.rnorm(2000))
    X = np.column_stack([locs[:, 0], O.rnorm(2000)])
    beta = np.array([0.01, O.rnorm(1)[0]])
    beta_0 = O.rnorm(1)[0]
    noise = np.sqrt(5.0) * O.rnorm(2000)
    return locs, field + noise + X @ beta + beta_0, X


if __name__ == "__main__":
    form = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    which = sys.argv[2] if len(sys.argv) > 2 else "both"
    locs, y, X = toy()
    t0 = time.time()
    if which in ("both", "locs"):
        lst = R.initialize(locs, y, X_locs=X, m=5, seed=1)
        p = lst["states"]["chain_1"]["params"]
        print("init", p["beta_0"], p["beta"], p["log_scale"], p["shape"], p["log_noise_variance"], p["field"][:3])
        R.run(lst, n_cycles=5, n_iterations_update=200, n_chromatic=5, burn_in=.5, field_thinning=.01, Gelman_Rubin_Brooks_stop=(1.0, 1.0), sweep_form=form, verbose=True)
        print("run 2", time.time() - t0)
        R.run(lst, n_cycles=1000, n_iterations_update=100, burn_in=.5, field_thinning=.2, Gelman_Rubin_Brooks_stop=(1.0, 1.05), sweep_form=form, verbose=True)
        print("run 3", time.time() - t0)
        R.run(lst, n_cycles=10, n_iterations_update=100, burn_in=.5, field_thinning=.2, Gelman_Rubin_Brooks_stop=(1.0, 1.0), sweep_form=form, verbose=True)
        e = R.estimate(lst, .5)
        np.set_printoptions(precision=9, linewidth=160)
        print("GpGp_covparams\n", e["GpGp_covparams"]); print("fixed_effects\n", e["fixed_effects"]); print("field\n", e["field"][:6])
        print("done", time.time() - t0)
    if which in ("both", "obs"):
        lst = R.initialize(locs, y, X_obs=X, m=5, seed=1)
        R.run(lst, n_cycles=5, n_iterations_update=200, burn_in=.5, field_thinning=.01, Gelman_Rubin_Brooks_stop=(1.0, 1.0), sweep_form=form, verbose=True)
        print("done", time.time() - t0)
