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
