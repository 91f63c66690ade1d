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
