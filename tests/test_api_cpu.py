"""Host-side pieces of the API mirror that need no GPU."""
import numpy as np

import nngp_b200 as nb
from nngp_b200 import api


def test_get_summary_matches_r_definitions():
    """estimate.R:1-6: mean, quantile type 7 at .025/.5/.975, sd with n-1"""
    x = np.array([[1.0, 10.0], [2.0, 20.0], [3.0, 30.0], [4.0, 50.0]])
    s = nb.get_summary(x)
    assert np.allclose(s[0], [2.5, 1.075, 2.5, 3.925, np.std(x[:, 0], ddof=1)])   # R: quantile(1:4, .025) = 1.075
    assert s.shape == (2, 5)


def test_shape_to_covparms_transforms():
    cp = api.shape_to_covparms([np.log(0.3), 0.0], ["log_range", "qlogis_smoothness"], lambda v: .5 + .5 * api.plogis(v))
    assert np.allclose(cp, [1.0, 0.3, 0.75, 0.0])
    cp = api.shape_to_covparms([0.0], ["qlogis_smoothness"], lambda v: 1.5 * api.plogis(v))      # predict.R:37
    assert np.allclose(cp, [1.0, 0.75, 0.0])


def test_model_matrix_expands_factors_with_treatment_contrasts():
    import pandas as pd
    df = pd.DataFrame({"a": [1.0, 2.0, 3.0], "f": pd.Categorical(["x", "y", "z"])})
    M, names = api._model_matrix(df)
    assert names == ["a", "fy", "fz"]
    assert np.array_equal(M, [[1, 0, 0], [2, 1, 0], [3, 0, 1]])


def test_gelman_rubin_brooks_on_identical_and_separated_chains():
    rng = np.random.default_rng(0)

    def chain(shift):
        p = {"beta_0": rng.standard_normal((400, 1)) + shift, "log_scale": rng.standard_normal((400, 1)),
             "log_noise_variance": rng.standard_normal((400, 1)), "shape": rng.standard_normal((400, 1)), "field": np.zeros((1, 3))}
        return {"params": p}

    good = nb.Gelman_Rubin_Brooks({"chain_1": chain(0), "chain_2": chain(0), "chain_3": chain(0)})
    bad = nb.Gelman_Rubin_Brooks({"chain_1": chain(0), "chain_2": chain(5), "chain_3": chain(-5)})
    assert np.all(good["R_hat"][1:] < 1.05)
    assert bad["R_hat"][1] > 2.0 and bad["R_hat"][0] > 2.0
    e = nb.ESS({"chain_1": chain(0), "chain_2": chain(0)})
    assert e.shape == (3, 4) and np.all(e[:2] > 100)


def test_host_mirror_reproduces_the_vignette_when_the_chains_come_from_the_oracle(monkeypatch, golden):
    """The host-side mirror alone -- api._initialize_host(rng = "R"), mcmc_nngp_run's bookkeeping, Gelman_Rubin_Brooks,
    mcmc_nngp_estimate -- with the device steps (initial field draw, the chains of a cycle) supplied by the CPU oracle through
    the same return schema: the first run of the vignette (Vignette.md:642-684) and the estimates computed from it must come
    out as the ORACLE's own driver computes them (oracle/reference_driver.py, itself pinned to the printed values).  The
    device steps themselves are compared with the printed values in tests/test_gpu_vignette.py."""
    from oracle import oracle as O
    from oracle import reference_driver as R

    rs = nb.RStream(1)
    locs = np.column_stack([500.0 * rs.runif(2000), np.ones(2000)])
    locs[0, 1] = 1.01
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    f0 = np.sqrt(10.0) * (np.linalg.cholesky(np.exp(-D / 5.0)) @ rs.rnorm(2000))
    X = np.column_stack([locs[:, 0], rs.rnorm(2000)])
    y = f0 + X @ np.array([0.01, rs.rnorm(1)[0]]) + rs.rnorm(1)[0]
    y = y + np.sqrt(5.0) * rs.rnorm(2000)
    lst, pending = api._initialize_host(locs, y, None, X, 5, "maxmin", "exponential_isotropic", "Gaussian", 3, 1, None, "R")
    va = lst["vecchia_approx"]
    for name, (cp, z) in pending.items():
        p = lst["states"][name]["params"]
        Linv = O.vecchia_Linv(cp, "exponential_isotropic", lst["locs"], va["NNarray"])
        p["field"] = p["beta_0"] + np.sqrt(np.exp(p["log_scale"])) * O.sparse_chol_solve(Linv, va["NNarray"], z)

    def oracle_update(locs, X, observed_field, space_time_model, vecchia_approx, states, n_iterations_update, n_cores=None,
                      field_thinning=1, ancillary=True, n_chromatic=10, iterations=None, n_gpus=None, rng="philox"):
        assert rng == "R"
        out = []
        for i, (name, st) in enumerate(states.items()):
            p, tk = st["params"], st["transition_kernels"]
            cp = dict(shape=p["shape"], beta_0=p["beta_0"], log_scale=p["log_scale"], log_noise_variance=p["log_noise_variance"],
                      logvar_sufficient=tk["covariance_params_sufficient"]["logvar"], logvar_ancillary=tk["covariance_params_ancillary"]["logvar"])
            reg = dict(X=X["X"], xlocs=np.array([c + 1 for c in X["locs"]], dtype=np.int32), first_obs=vecchia_approx["hctam_scol_1"],
                       solve_1XT1X=X["solve_1XT1X"], chol_solve_1XT1X=X["chol_solve_1XT1X"], beta=p["beta"])
            po, f, rec, frec, _, brec = O.update_gaussian_chain(locs, vecchia_approx["NNarray"], vecchia_approx["coloring"], vecchia_approx["locs_match"],
                                                                vecchia_approx["obs_per_loc"], observed_field, "exponential_isotropic", cp, p["field"],
                                                                n_iterations_update, field_thinning, n_chromatic, int(iterations[-1, 0]), i + 1, 1, regressors=reg)
            state = {"transition_kernels": {"covariance_params_sufficient": {"logvar": po["logvar_sufficient"]},
                                            "covariance_params_ancillary": {"logvar": po["logvar_ancillary"]}, "log_noise_variance": {"logvar": -1.0}},
                     "params": {"shape": po["shape"], "beta_0": po["beta_0"], "log_scale": po["log_scale"], "log_noise_variance": po["log_noise_variance"],
                                "field": f, "beta": np.array(po["beta"]).copy()}}
            out.append({"state": state, "records": api._records_dict(rec, space_time_model["covfun"]["shape_params"], frec, beta=brec, beta_names=X["names"])})
        return out

    monkeypatch.setattr(api, "mcmc_nngp_update_Gaussian", oracle_update)
    lst = api.mcmc_nngp_run(lst, n_cores=3, n_cycles=5, n_iterations_update=200, n_chromatic=5, burn_in=.5, field_thinning=.01,
                            Gelman_Rubin_Brooks_stop=(1.0, 1.0), rng="R", verbose=False)
    got = np.array([d["R_hat"] for d in lst["diagnostics"]["Gelman_Rubin_Brooks"]])
    want = np.array([b["R_hat"] for b in golden["R_hat_blocks"][:5]])
    assert np.max(np.abs(got - want)) < 6e-7                                                    # the printed values themselves
    # estimates after these 1000 iterations: the vignette prints them only after 4600, so compare with the oracle's driver
    ref = R.initialize(locs, y, X_locs=X, m=5, seed=1)
    R.run(ref, n_cycles=5, n_iterations_update=200, n_chromatic=5, burn_in=.5, field_thinning=.01, Gelman_Rubin_Brooks_stop=(1.0, 1.0), sweep_form=1)
    e, er = api.mcmc_nngp_estimate(lst, burn_in=.5), R.estimate(ref, burn_in=.5)
    assert np.allclose(e["covariance_params"]["GpGp_covparams"]["summary"], er["GpGp_covparams"], rtol=1e-12)
    assert np.allclose(e["fixed_effects"]["summary"], er["fixed_effects"], rtol=1e-12)
    assert np.allclose(e["field"], er["field"], rtol=1e-12, atol=1e-13)
