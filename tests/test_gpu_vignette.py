"""The reference's vignette replayed THROUGH THE PRODUCT -- the Python mirror of the reference's entry points over the C ABI,
everything numerical on the GPU -- and compared with the numbers the vignette prints (tests/golden/vignette_golden.json,
extracted from /root/reference/Vignette.md).  The calls are the vignette's own (Vignette.rmd:74-78, Vignette.md:642-644,
:687-689, :879-881, :995, :1129-1141) with rng = "R": R's random stream, consumed in the reference's order.

tests/test_vignette_pin.py does the same with the CPU oracle; here the GPU library is checked against the reference's
printed output directly, not via the oracle."""
import numpy as np
import pytest

import nngp_b200 as nb

pytestmark = pytest.mark.gpu

# Printed values carry 6 decimals.  The device arithmetic differs from R's in the last bits (summation order), and a
# Gelman-Rubin-Brooks value is a ratio of variances of 100-2300 samples, so compare to 1e-5 relative: a single differing
# accept / reject decision anywhere earlier in the session moves these values in the second or third digit.
RTOL = 1e-5


def vignette_toy():
    """Vignette.rmd:26-47 with R's stream (product side: nb.RStream)"""
    rs = nb.RStream(1)
    locs = np.column_stack([500.0 * rs.runif(2000), np.ones(2000)])
    locs[0, 1] = 1.01
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    field = np.sqrt(10.0) * (np.linalg.cholesky(np.exp(-D / 5.0)) @ rs.rnorm(2000))
    X = np.column_stack([locs[:, 0], rs.rnorm(2000)])
    beta = np.array([0.01, rs.rnorm(1)[0]])
    beta_0 = rs.rnorm(1)[0]
    noise = np.sqrt(5.0) * rs.rnorm(2000)
    return locs, field + noise + X @ beta + beta_0, X


@pytest.fixture(scope="module")
def toy():
    return vignette_toy()


def rhat(lst):
    return np.array([d["R_hat"] for d in lst["diagnostics"]["Gelman_Rubin_Brooks"]])


def blocks(golden, a, b):
    return np.array([x["R_hat"] for x in golden["R_hat_blocks"][a:b]])


def check_initial_state(lst, golden):
    va = lst["vecchia_approx"]
    assert np.array_equal(va["hctam_scol_1"][:100], golden["hctam_scol_1_100"])                        # Vignette.md:406-419
    assert np.array_equal(va["locs_match"][:100], golden["locs_match_100"])                            # :322-328
    g, p = golden["init_chain_1"], lst["states"]["chain_1"]["params"]
    assert abs(p["beta_0"] - g["beta_0"]) < 6e-6 and np.allclose(p["beta"], g["beta"], rtol=0, atol=6e-10)   # :476, :483
    assert abs(p["log_scale"] - g["log_scale"]) < 6e-7 and abs(p["shape"][0] - g["shape"][0]) < 6e-7          # :489, :495
    assert abs(p["log_noise_variance"] - g["log_noise_variance"]) < 6e-7                               # :501
    # factor build + triangular solve on the device against GpGp::vecchia_Linv + Matrix::solve, printed with 8 decimals
    assert np.allclose(p["field"][:100], g["field_100"], rtol=0, atol=2e-8)                            # :507-523


def test_initialize_matches_the_printed_initial_state(toy, golden):
    locs, y, X = toy
    lst = nb.mcmc_nngp_initialize(locs, y, X_locs=X, m=5, stationary_covfun="exponential_isotropic", seed=1, rng="R")
    check_initial_state(lst, golden)
    nb.release_contexts(lst)


def test_first_run_matches_the_printed_diagnostics(toy, golden):
    """Vignette.md:642-684: 5 cycles of 200 iterations, 3 chains, n_chromatic = 5, location-level regressors (interweaving)"""
    locs, y, X = toy
    lst = nb.mcmc_nngp_initialize(locs, y, X_locs=X, m=5, stationary_covfun="exponential_isotropic", seed=1, rng="R")
    lst = nb.mcmc_nngp_run(lst, n_cores=3, n_cycles=5, n_iterations_update=200, ancillary=True, n_chromatic=5, burn_in=.5,
                           field_thinning=.01, Gelman_Rubin_Brooks_stop=(1.0, 1.0), rng="R", verbose=False)
    assert np.allclose(rhat(lst), blocks(golden, 0, 5), rtol=RTOL, atol=0)
    nb.release_contexts(lst)


def test_observation_level_regressors_match_the_printed_diagnostics(toy, golden):
    """Vignette.md:1129-1178: the same regressors passed as X_obs; 5 cycles of 200 iterations"""
    locs, y, X = toy
    lst = nb.mcmc_nngp_initialize(locs, y, X_obs=X, m=5, stationary_covfun="exponential_isotropic", seed=1, rng="R")
    check_initial_state(lst, golden)
    lst = nb.mcmc_nngp_run(lst, n_cores=3, n_cycles=5, n_iterations_update=200, burn_in=.5, field_thinning=.01,
                           Gelman_Rubin_Brooks_stop=(1.0, 1.0), rng="R", verbose=False)
    assert np.allclose(rhat(lst), blocks(golden, 41, 46), rtol=RTOL, atol=0)
    nb.release_contexts(lst)


def test_whole_session_matches_every_printed_diagnostic_and_estimate(toy, golden):
    """Vignette.md:642-1028: 1000 + 2600 + 1000 iterations of 3 chains in 41 cycles (the second run stops by itself after its
    26th cycle), then mcmc_nngp_estimate: covariance parameters, fixed effects, head of the field summaries"""
    locs, y, X = toy
    lst = nb.mcmc_nngp_initialize(locs, y, X_locs=X, m=5, stationary_covfun="exponential_isotropic", seed=1, rng="R")
    lst = nb.mcmc_nngp_run(lst, n_cores=3, n_cycles=5, n_iterations_update=200, ancillary=True, n_chromatic=5, burn_in=.5,
                           field_thinning=.01, Gelman_Rubin_Brooks_stop=(1.0, 1.0), rng="R", verbose=False)
    lst = nb.mcmc_nngp_run(lst, n_cores=3, n_cycles=1000, n_iterations_update=100, burn_in=.5, field_thinning=.2,
                           Gelman_Rubin_Brooks_stop=(1.0, 1.05), rng="R", verbose=False)
    assert len(lst["diagnostics"]["Gelman_Rubin_Brooks"]) == 31
    lst = nb.mcmc_nngp_run(lst, n_cores=3, n_cycles=10, n_iterations_update=100, burn_in=.5, field_thinning=.2,
                           Gelman_Rubin_Brooks_stop=(1.0, 1.0), rng="R", verbose=False)
    assert int(lst["records"]["chain_1"]["iterations"][-1, 0]) == 4600
    assert np.allclose(rhat(lst), blocks(golden, 0, 41), rtol=RTOL, atol=0)
    est = nb.mcmc_nngp_estimate(lst, burn_in=.5)
    g = golden["estimate"]
    assert est["covariance_params"]["GpGp_covparams"]["names"] == ["scale", "noise_variance", "range"]
    assert np.allclose(est["covariance_params"]["GpGp_covparams"]["summary"], g["GpGp_covparams"], rtol=RTOL, atol=0)   # :1000-1002
    assert np.allclose(est["fixed_effects"]["summary"], g["fixed_effects"], rtol=RTOL, atol=0)                          # :1009-1011
    assert np.allclose(est["field"][:6], g["field_head"], rtol=RTOL, atol=0)                                            # :1022-1027
    nb.release_contexts(lst)
