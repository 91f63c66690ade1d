"""The reference's vignette (Vignette.rmd) through the GPU library, printing what the vignette prints next to the printed values
(tests/golden/vignette_golden.json).  Same calls as tests/test_gpu_vignette.py; needs a CUDA device.

    python scripts/vignette_gpu.py            # first example (location-level regressors): 41 cycles + estimates
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nngp_b200 as nb  # noqa: E402


def main():
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "vignette_golden.json")))
    rs = nb.RStream(1)                                                               # Vignette.rmd:26-47
    locs = np.column_stack([500.0 * rs.runif(2000), np.ones(2000)])
    locs[0, 1] = 1.01
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    field = np.sqrt(10.0) * (np.linalg.cholesky(np.exp(-D / 5.0)) @ rs.rnorm(2000))
    X = np.column_stack([locs[:, 0], rs.rnorm(2000)])
    beta = np.array([0.01, rs.rnorm(1)[0]])
    beta_0 = rs.rnorm(1)[0]
    y = field + np.sqrt(5.0) * rs.rnorm(2000) + X @ beta + beta_0
    t0 = time.time()
    lst = nb.mcmc_nngp_initialize(locs, y, X_locs=X, m=5, stationary_covfun="exponential_isotropic", seed=1, rng="R")
    p = lst["states"]["chain_1"]["params"]
    print("chain_1 initial field[1:6]  ", np.round(p["field"][:6], 8))
    print("printed (Vignette.md:507)   ", np.array(golden["init_chain_1"]["field_100"][:6]))
    runs = [dict(n_cycles=5, n_iterations_update=200, n_chromatic=5, field_thinning=.01, Gelman_Rubin_Brooks_stop=(1.0, 1.0)),
            dict(n_cycles=1000, n_iterations_update=100, field_thinning=.2, Gelman_Rubin_Brooks_stop=(1.0, 1.05)),
            dict(n_cycles=10, n_iterations_update=100, field_thinning=.2, Gelman_Rubin_Brooks_stop=(1.0, 1.0))]
    for kw in runs:
        lst = nb.mcmc_nngp_run(lst, n_cores=3, burn_in=.5, rng="R", verbose=False, **kw)
    for k, d in enumerate(lst["diagnostics"]["Gelman_Rubin_Brooks"]):
        print(f"block {k + 1:2d}  gpu    ", np.array2string(d["R_hat"], precision=6, floatmode="fixed"))
        print(f"          printed", np.array2string(np.array(golden["R_hat_blocks"][k]["R_hat"]), precision=6, floatmode="fixed"))
    est = nb.mcmc_nngp_estimate(lst, burn_in=.5)
    print("GpGp_covparams (gpu)\n", est["covariance_params"]["GpGp_covparams"]["summary"])
    print("printed (Vignette.md:1000-1002)\n", np.array(golden["estimate"]["GpGp_covparams"]))
    print(f"{int(lst['records']['chain_1']['iterations'][-1, 0])} iterations of 3 chains in {time.time() - t0:.1f} s")
    nb.release_contexts(lst)


if __name__ == "__main__":
    main()
