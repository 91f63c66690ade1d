"""Pins the CPU oracle against the numbers the reference itself prints (the rendered vignette, /root/reference/Vignette.md,
extracted into tests/golden/vignette_golden.json by tests/golden/make_vignette_golden.py).

The vignette is a complete session of the reference on seeded toy data: mcmc_nngp_initialize(seed = 1), 4600 iterations of
three chains over 41 cycles of mcmc_nngp_run, mcmc_nngp_estimate, and a second model with the regressors passed per
observation.  Because the reference seeds everything (initialize.R:17 set.seed(seed); update_Gaussian.R:36
set.seed(iter_start + i) inside every forked chain) the session is deterministic, and oracle/reference_driver.py replays it on
R's random stream.  What these tests establish, to the last printed digit:

  * GpGp::order_maxmin / find_ordered_nn as restated in oracle/gpgp_order.c (ordering, locs_match, NNarray);
  * GpGp::vecchia_Linv + Matrix::solve (the initial field, 100 values printed with 8 decimals);
  * the whole of mcmc_nngp_update_Gaussian (log-likelihood ratios, ancillary / sufficient steps, the regression block with
    and without interweaving, the chromatic sweep in both of the oracle's forms, the noise-variance steps): any accept /
    reject decision that differed anywhere in 13 800 chain iterations would change every later Gelman-Rubin-Brooks value.
"""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import reference_driver as R

NA = O.NA_INT


def vignette_toy():
    """Vignette.rmd:26-47"""
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


@pytest.fixture(scope="module")
def toy():
    return vignette_toy()


def check_init(lst, golden):
    """everything the vignette prints about the freshly initialised list (Vignette.md:133-523)"""
    va = lst["vecchia_approx"]
    assert np.array_equal(va["hctam_scol_1"][:100], golden["hctam_scol_1_100"])                        # Vignette.md:406-419
    assert np.array_equal(va["locs_match"][:100], golden["locs_match_100"])                            # :322-328 (whole permutation)
    assert np.allclose(lst["locs"][:6], golden["locs_head"], rtol=0, atol=5e-6)                        # :148-154
    want = np.array([[NA if v is None else v for v in r] for r in golden["NNarray_head"]], dtype=np.int32)
    assert np.array_equal(va["NNarray"][:6], want)                                                     # :221-227
    assert np.array_equal(va["sparse_chol_column_idx"][:100], np.arange(1, 101))                       # :247-252
    assert np.array_equal(va["sparse_chol_row_idx"][:100], np.arange(1, 101))                          # :259-264
    if lst["X"]["X"] is not None:
        assert np.allclose(lst["X"]["X"][:6], golden["X_head"], rtol=0, atol=5e-6)                     # :180-186
    g = golden["init_chain_1"]
    p = lst["states"]["chain_1"]["params"]
    assert abs(p["beta_0"] - g["beta_0"]) < 6e-6                                                       # :476  "2.88572"
    assert np.allclose(p["beta"], g["beta"], rtol=0, atol=6e-10)                                       # :483
    assert abs(p["log_scale"] - g["log_scale"]) < 6e-7                                                 # :489
    assert abs(p["shape"][0] - g["shape"][0]) < 6e-7                                                   # :495
    assert abs(p["log_noise_variance"] - g["log_noise_variance"]) < 6e-7                               # :501
    assert np.allclose(p["field"][:100], g["field_100"], rtol=0, atol=6e-9)                            # :507-523, 8 decimals


def test_initialize_reproduces_the_vignette(toy, golden):
    locs, y, X = toy
    lst = R.initialize(locs, y, X_locs=X, m=5, seed=1)                                                 # Vignette.rmd:74-78
    check_init(lst, golden)
    # the printed 30 x 30 block of the moral graph, from the regenerated (not the printed) ordering
    adj_p, adj_i = lst["vecchia_approx"]["MRF_adjacency"]
    A = np.zeros((30, 30), dtype=int)
    for s in range(30):
        for t in adj_i[adj_p[s]:adj_p[s + 1]]:
            if t < 30:
                A[t, s] = 1
    assert np.array_equal(A, np.array(golden["MRF_adjacency_30"]))                                     # :275-306


def rhat_matches(got, blocks):
    got = np.array(got)
    want = np.array([b["R_hat"] for b in blocks])
    assert got.shape == want.shape, (got.shape, want.shape)
    # every block is printed with 6 decimals
    assert np.max(np.abs(got - want)) < 6e-7, np.argwhere(np.abs(got - want) >= 6e-7)[:3]


def test_whole_session_reproduces_every_printed_diagnostic_and_estimate(toy, golden):
    """Vignette.md:642-1028: 5 x 200 + 26 x 100 + 10 x 100 iterations of 3 chains, 41 Gelman-Rubin-Brooks blocks (the second
    run has to stop by itself after its 26th cycle), then mcmc_nngp_estimate.  The first run uses the sweep exactly as the
    reference writes it (one full product per colour), the others the residual-maintained form that the GPU kernel uses."""
    locs, y, X = toy
    B = golden["R_hat_blocks"]
    lst = R.initialize(locs, y, X_locs=X, m=5, seed=1)
    R.run(lst, n_cycles=5, n_iterations_update=200, n_chromatic=5, burn_in=.5, field_thinning=.01,
          Gelman_Rubin_Brooks_stop=(1.0, 1.0), sweep_form=0)                                           # :642-644
    rhat_matches(lst["diagnostics"], B[:5])
    R.run(lst, n_cycles=1000, n_iterations_update=100, burn_in=.5, field_thinning=.2, Gelman_Rubin_Brooks_stop=(1.0, 1.05),
          sweep_form=1)                                                                                # :687-689
    assert len(lst["diagnostics"]) == 31                                                               # stopped after cycle 26
    rhat_matches(lst["diagnostics"], B[:31])
    R.run(lst, n_cycles=10, n_iterations_update=100, burn_in=.5, field_thinning=.2, Gelman_Rubin_Brooks_stop=(1.0, 1.0),
          sweep_form=1)                                                                                # :879-881
    rhat_matches(lst["diagnostics"], B[:41])
    assert lst["records"]["chain_1"]["iterations"][-1] == 4600
    e = R.estimate(lst, burn_in=.5)                                                                    # :995
    g = golden["estimate"]
    assert e["covparams_names"] == ["scale", "noise_variance", "range"]
    assert np.allclose(e["GpGp_covparams"], g["GpGp_covparams"], rtol=6e-7, atol=0)                    # :1000-1002
    assert np.allclose(e["fixed_effects"], g["fixed_effects"], rtol=6e-7, atol=0)                      # :1009-1011
    assert np.allclose(e["field"][:6], g["field_head"], rtol=6e-7, atol=0)                             # :1022-1027


def test_observation_level_regressors_session(toy, golden):
    """Vignette.md:1129-1178: the same regressors passed as X_obs (no interweaving, and the beta_0-only update of
    update_Gaussian.R:219-224 runs as well because length(X$locs) == 0): 5 x 200 iterations, 5 printed blocks."""
    locs, y, X = toy
    lst = R.initialize(locs, y, X_obs=X, m=5, seed=1)
    check_init(lst, golden)                                                                            # same seed, same draws
    R.run(lst, n_cycles=5, n_iterations_update=200, burn_in=.5, field_thinning=.01, Gelman_Rubin_Brooks_stop=(1.0, 1.0), sweep_form=0)
    rhat_matches(lst["diagnostics"], golden["R_hat_blocks"][41:46])
