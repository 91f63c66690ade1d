"""Seeded synthetic NNGP problems shared by the parity tests, smoke() and bench.py."""
from __future__ import annotations

import numpy as np

import nngp_b200 as nb


def make_problem(n, m, d=2, seed=0, n_extra_obs=0, drop_obs_frac=0.0, locs=None):
    """U(0,1)^d sites in the order generated (= reordering "random"), exact ordered NN, first-fit colouring, y = w + noise."""
    rng = np.random.default_rng(seed)
    if locs is None:
        locs = rng.random((n, d))
    nn = nb.find_ordered_nn(locs, m)
    coloring = nb.greedy_coloring(nn)
    sites = np.arange(1, n + 1, dtype=np.int32)
    if drop_obs_frac > 0:
        sites = sites[rng.random(n) >= drop_obs_frac]
    if n_extra_obs > 0:
        sites = np.concatenate([sites, rng.integers(1, n + 1, n_extra_obs).astype(np.int32)])
    locs_match = sites
    n_obs = locs_match.size
    obs_per_loc = np.bincount(locs_match - 1, minlength=n).astype(np.float64)
    y = rng.standard_normal(n_obs)
    field = 0.3 + rng.standard_normal(n)
    return dict(rng=rng, locs=locs, NNarray=nn, coloring=coloring, locs_match=locs_match, n_obs=n_obs,
                obs_per_loc=obs_per_loc, y=y, field=field, n=n, m=m, d=d)


def make_regression_problem(n, m, seed=0, n_extra_obs=0, p_locs=2, p_obs=1, range_=0.1):
    """A Gaussian NNGP model WITH regressors as mcmc_nngp_initialize prepares it (initialize.R:116-137): X$X = centred
    cbind(X_locs[locs_match], X_obs) without intercept, X$locs = 1..p_locs, solve_1XT1X / chol_solve_1XT1X (upper factor),
    hctam_scol_1 = first observation of every site; y = beta_0 + X beta + w[locs_match] + noise with w drawn from the prior
    by the host utilities (no oracle, no GPU)."""
    P = make_problem(n, m, seed=seed, n_extra_obs=n_extra_obs)
    rng, lm = P["rng"], P["locs_match"] - 1
    n_obs = P["n_obs"]
    cols = []
    if p_locs:
        Xl = np.column_stack([P["locs"][:, 0]] + [rng.standard_normal(n) for _ in range(p_locs - 1)])
        cols.append(Xl[lm])
    if p_obs:
        cols.append(rng.standard_normal((n_obs, p_obs)))
    X = np.column_stack(cols)
    X = X - X.mean(axis=0)                                                       # initialize.R:132
    one = np.column_stack([np.ones(n_obs), X])
    S = np.linalg.inv(one.T @ one)                                               # :135
    first = np.full(n, -1, dtype=np.int64)
    for o in range(n_obs - 1, -1, -1):
        first[lm[o]] = o
    beta_true = rng.standard_normal(X.shape[1])
    # a smooth field with roughly the right range (exact prior draws are not needed for these tests)
    w = np.sin(P["locs"][:, 0] / range_) * np.cos(P["locs"][:, 1] / range_) + 0.3 * rng.standard_normal(n)
    y = 0.7 + X @ beta_true + w[lm] + np.sqrt(0.1) * rng.standard_normal(n_obs)
    P.update(X=X, xlocs=np.arange(1, p_locs + 1, dtype=np.int32), first_obs=(first + 1).astype(np.int32), solve_1XT1X=S,
             chol_solve_1XT1X=np.linalg.cholesky(S).T, beta_true=beta_true, y=y, w=w)
    return P
