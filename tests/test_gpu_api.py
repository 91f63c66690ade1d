"""End-to-end runs of the reference's entry points (Python mirror over the C ABI) with distributional acceptance bands."""
import numpy as np
import pytest

import nngp_b200 as nb
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def vignette_toy():
    """Vignette.rmd:26-47 regenerated with R's own random stream (oracle/r_rng.c): same locs, field, X, beta, noise."""
    O.set_seed(1)
    locs = np.column_stack([500.0 * O.runif(2000), np.ones(2000)])
    locs[0, 1] = 1.01
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    Lc = np.linalg.cholesky(np.exp(-D / 5.0))                       # t(chol(GpGp::exponential_isotropic(c(1,5,0), locs)))
    field = np.sqrt(10.0) * (Lc @ O.rnorm(2000))
    X = np.column_stack([locs[:, 0], O.rnorm(2000)])
    beta = np.array([0.01, O.rnorm(1)[0]])
    beta_0 = O.rnorm(1)[0]
    noise = np.sqrt(5.0) * O.rnorm(2000)
    y = field + noise + X @ beta + beta_0
    return locs, y, X, beta, beta_0, field


def test_config1a_vignette_toy_posterior_bands(golden):
    """config 1a: n = 2000 quasi-1-D, m = 5, exponential, 2 location regressors, 3 chains.  Acceptance: posterior means
    inside the 95 % intervals printed by the reference's vignette (Vignette.md:999-1011); truth 10 / 5 / 5."""
    locs, y, X, beta, beta_0, _ = vignette_toy()
    lst = nb.mcmc_nngp_initialize(locs, y, X_locs=X, m=5, stationary_covfun="exponential_isotropic", n_chains=3, seed=1)
    assert sorted(lst.keys()) == sorted(["locs", "X", "observed_field", "observed_locs", "space_time_model", "vecchia_approx", "states",
                                         "records", "diagnostics", "t_begin", "seed"])
    va = lst["vecchia_approx"]
    assert va["NNarray"].shape == (2000, 6) and va["coloring"].min() == 1
    assert np.array_equal(lst["locs"][va["locs_match"] - 1], lst["observed_locs"])
    lst = nb.mcmc_nngp_run(lst, n_cycles=8, n_iterations_update=250, Gelman_Rubin_Brooks_stop=(1.0, 1.0), verbose=False)
    it = int(lst["records"]["chain_1"]["iterations"][-1, 0])
    assert it == 2000 and lst["records"]["chain_1"]["params"]["field"].shape == (2000, 2000)
    est = nb.mcmc_nngp_estimate(lst, burn_in=.5)
    g = dict(zip(est["covariance_params"]["GpGp_covparams"]["names"], est["covariance_params"]["GpGp_covparams"]["summary"][:, 0]))
    bands = golden["posterior_bands"]
    assert bands["scale"][1] < g["scale"] < bands["scale"][2]
    assert bands["noise_variance"][1] < g["noise_variance"] < bands["noise_variance"][2]
    assert bands["range"][1] < g["range"] < bands["range"][2]
    fx = dict(zip(est["fixed_effects"]["names"], est["fixed_effects"]["summary"][:, 0]))
    assert abs(fx["V2"] - beta[1]) < 0.25                    # white-noise regressor: vignette -1.598 +- 0.055
    assert abs(fx["V1"] - 0.01) < 0.02                       # slope (vignette: 0.0032 +- 0.0036)
    nb.release_contexts(lst)


def test_config1b_no_regressor_chain_and_prediction():
    """config 1b: 2-D, n = 5000, m = 10, one chain through nngp_chain_run; then mcmc_nngp_predict_field."""
    rng = np.random.default_rng(3)
    n = 5000
    locs = rng.random((n, 2)) * 100.0
    nn = nb.find_ordered_nn(locs, 10)
    with nb.NNGPContext(locs, nn, nb.greedy_coloring(nn), np.arange(1, n + 1)) as ctx:
        ctx.factor_build([1.0, 5.0, 0.0])
        ctx.field_init(0.0, np.log(4.0), rng.standard_normal(n))
        w = ctx.field_get()
    y = 2.0 + w + np.sqrt(0.5) * rng.standard_normal(n)
    lst = nb.mcmc_nngp_initialize(locs, y, m=10, reordering="none", n_chains=1, seed=2)
    lst = nb.mcmc_nngp_run(lst, n_cycles=5, n_iterations_update=300, field_thinning=.1, Gelman_Rubin_Brooks_stop=(1.0, 1.0), verbose=False)
    est = nb.mcmc_nngp_estimate(lst, burn_in=.5)
    g = dict(zip(est["covariance_params"]["GpGp_covparams"]["names"], est["covariance_params"]["GpGp_covparams"]["summary"][:, 0]))
    assert 2.0 < g["scale"] < 8.0 and 0.3 < g["noise_variance"] < 0.8 and 2.5 < g["range"] < 10.0
    # the posterior mean of the latent field tracks the simulated one
    fmean = est["field"][:, 0]
    assert np.corrcoef(fmean, w)[0, 1] > 0.9
    # conditional simulation at new sites next to observed ones
    new = locs[:300] + 0.05
    pred = nb.mcmc_nngp_predict_field(lst, new, burn_in=.5, m=10)
    assert pred["predicted_field_summary"].shape == (300, 5)
    assert np.corrcoef(pred["predicted_field_summary"][:, 0], w[:300])[0, 1] > 0.85
    nb.release_contexts(lst)
