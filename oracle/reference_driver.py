"""oracle/reference_driver.py -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).

The reference's entry points restated over the C oracle, on R's own random stream, so that a whole session of the
reference -- mcmc_nngp_initialize, mcmc_nngp_run cycle after cycle, mcmc_nngp_estimate -- can be replayed here and compared
with the numbers the reference's rendered vignette prints (tests/test_vignette_pin.py).  Every draw is taken in the
order R takes it:

  initialize  Scripts/mcmc_nngp_initialize.R:17 set.seed(seed); :29 GpGp::order_maxmin (rnorm(n*d), sample(n));
              :93 GpGp::find_ordered_nn (rnorm(n*d)); :154-161 sample(., 1) (+ rnorm(1) for Matern) per chain;
              :189 rnorm(p+1); :193-194 rbeta, rbeta; :208 rnorm(n_locs)  -- per chain, in that order
  run         Scripts/mcmc_nngp_run.R:8-48; each chain of a cycle starts from set.seed(iter_start + i)
              (Scripts/mcmc_nngp_update_Gaussian.R:36) -- deterministic although the chains are forked
  diagnostics Scripts/mcmc_nngp_diagnose.R:1-24;  summaries Scripts/mcmc_nngp_estimate.R:1-96

Only tests/ may import this module.  Python dicts stand for R lists; index arrays stay 1-based.
"""
from __future__ import annotations

import numpy as np

from . import oracle as O

SHAPE_PARAMS = {                                                                     # initialize.R:62-69
    "exponential_isotropic": lambda d: ["log_range"],
    "exponential_sphere": lambda d: ["log_range"],
    "exponential_scaledim": lambda d: [f"log_range_{k + 1}" for k in range(d)],
    "exponential_spacetime": lambda d: ["log_range_1", "log_range_2"],
    "matern_isotropic": lambda d: ["log_range", "qlogis_smoothness"],
    "matern_sphere": lambda d: ["log_range", "qlogis_smoothness"],
    "matern_scaledim": lambda d: [f"log_range_{k + 1}" for k in range(d)] + ["qlogis_smoothness"],
    "matern_spacetime": lambda d: ["log_range_1", "log_range_2", "qlogis_smoothness"],
}


def _max_dist(P):
    return float(np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)).max())


def _sample_one(x):
    """sample(x, 1) = x[sample.int(length(x), 1)]"""
    return x[int(O.sample_int(len(x), 1)[0]) - 1]


def initialize(observed_locs, observed_field, X_obs=None, X_locs=None, m=10, reordering="maxmin",
               stationary_covfun="exponential_isotropic", n_chains=3, seed=1):
    """mcmc_nngp_initialize (Gaussian response; numeric regressors; not the *_sphere orderings, whose lon/lat branch of
    GpGp::order_maxmin no reference output pins)."""
    observed_locs = np.asarray(observed_locs, dtype=np.float64)
    observed_field = np.asarray(observed_field, dtype=np.float64).ravel()
    O.set_seed(seed)                                                                                   # :17
    _, first = np.unique(observed_locs, axis=0, return_index=True)
    locs = observed_locs[np.sort(first)]                                                               # :28
    if reordering == "maxmin":
        assert "sphere" not in stationary_covfun
        order = O.order_maxmin_gpgp(locs)                                                              # :29
    elif reordering == "random":
        order = O.sample_perm(locs.shape[0])                                                           # :30
    else:
        raise ValueError(reordering)
    locs = locs[order - 1]                                                                             # :34
    n, d = locs.shape
    shape_params = SHAPE_PARAMS[stationary_covfun](d)
    key = {tuple(r): i + 1 for i, r in enumerate(map(tuple, locs))}
    locs_match = np.array([key[tuple(r)] for r in observed_locs], dtype=np.int32)                      # :85
    n_obs = observed_field.size
    hctam_scol_1 = np.zeros(n, dtype=np.int32)
    for o in range(n_obs - 1, -1, -1):
        hctam_scol_1[locs_match[o] - 1] = o + 1                                                        # :88-89
    obs_per_loc = np.bincount(locs_match - 1, minlength=n).astype(np.float64)                          # :91
    NN = O.find_ordered_nn_gpgp(locs, m)                                                               # :93
    non_na = NN != O.NA_INT
    col_idx = NN.T[non_na.T]                                                                           # :99
    row_idx = np.tile(np.arange(1, n + 1, dtype=np.int32)[:, None], (1, m + 1)).T[non_na.T]            # :101
    adj_p, adj_i = O.moral_graph(NN)                                                                   # :103
    coloring = O.naive_greedy_coloring(adj_p, adj_i)                                                   # :110
    va = dict(n_locs=n, n_obs=n_obs, locs_match=locs_match, hctam_scol_1=hctam_scol_1, obs_per_loc=obs_per_loc, NNarray=NN,
              NNarray_non_NA=non_na, sparse_chol_column_idx=col_idx, sparse_chol_row_idx=row_idx, coloring=coloring,
              MRF_adjacency=(adj_p, adj_i))
    X = dict(X=None, locs=[])
    parts = [np.asarray(a, dtype=np.float64).reshape(n_obs, -1) for a in (X_locs, X_obs) if a is not None]   # :118-121
    if parts:
        XX = np.column_stack(parts)
        X["locs"] = list(range(1, (0 if X_locs is None else parts[0].shape[1]) + 1))                   # :129 (1-based)
        X["X_mean"] = XX.mean(axis=0)                                                                  # :131
        XX = XX - X["X_mean"]                                                                          # :132
        X["X"] = XX
        one = np.column_stack([np.ones(n_obs), XX])
        X["solve_1XT1X"] = np.linalg.inv(one.T @ one)                                                  # :135
        X["chol_solve_1XT1X"] = np.linalg.cholesky(X["solve_1XT1X"]).T                                 # :136 (upper factor)
    head = locs[:100]
    shapes = []
    for _ in range(n_chains):                                                                          # :152-162
        sh = []
        for name in shape_params:
            if name.startswith("log_range"):
                if "scaledim" in stationary_covfun:
                    P = head[:, [int(name.split("_")[-1]) - 1]]
                elif "spacetime" in stationary_covfun:
                    P = head[:, :-1] if name.endswith("_1") else head[:, [-1]]
                else:
                    P = head
                sh.append(_sample_one(np.log(_max_dist(P)) - np.log(np.arange(20.0, 201.0))))
            else:
                sh.append(float(O.rnorm(1)[0]))
        shapes.append(np.array(sh))
    design = np.ones((n_obs, 1)) if X["X"] is None else np.column_stack([np.ones(n_obs), X["X"]])      # :173-174 lm()
    G = np.linalg.inv(design.T @ design)
    coef = G @ (design.T @ observed_field)
    resid = observed_field - design @ coef
    vcov = G * (resid @ resid) / (n_obs - design.shape[1])
    var_resid = resid.var(ddof=1)
    states = {}
    for i in range(n_chains):                                                                          # :181-209
        perturb = np.linalg.cholesky(vcov) @ O.rnorm(coef.size)                                        # :189
        params = dict(shape=shapes[i], beta_0=float(coef[0] + perturb[0]))
        if X["X"] is not None:
            params["beta"] = coef[1:] + perturb[1:]                                                    # :191
        params["log_scale"] = float(np.log(O.rbeta(10, 10) * var_resid))                               # :193
        params["log_noise_variance"] = float(np.log(O.rbeta(10, 10) * var_resid))                      # :194
        cp = [1.0]
        for v, name in zip(shapes[i], shape_params):                                                   # :196-200
            cp.append(float(np.exp(v)) if name.startswith("log") else float(.4 + .7 / (1.0 + np.exp(-v))))
        cp.append(0.0)
        Linv = O.vecchia_Linv(cp, stationary_covfun, locs, NN)                                         # :201
        z = O.rnorm(n)
        params["field"] = params["beta_0"] + np.sqrt(np.exp(params["log_scale"])) * O.sparse_chol_solve(Linv, NN, z)   # :208
        states[f"chain_{i + 1}"] = dict(params=params, transition_kernels=dict(sufficient=-2.0, ancillary=-2.0))   # :185-187
    records = {k: dict(iterations=[0], saved_field=[], params={}) for k in states}
    return dict(locs=locs, X=X, observed_field=observed_field, observed_locs=observed_locs, order=order,
                space_time_model=dict(stationary_covfun=stationary_covfun, shape_params=shape_params), vecchia_approx=va,
                states=states, records=records, diagnostics=[], seed=seed)


def gelman_rubin_brooks(records, burn_in=.5):
    """diagnose.R:1-24.  Columns in the order of names(records$chain_1$params) without "field": beta_0, beta, log_scale,
    log_noise_variance, shape (update_Gaussian.R:42-56); rows seq(burn_in*n, n) (1-based, both ends)."""
    chains = []
    for r in records.values():
        p = r["params"]
        cols = [p["beta_0"]] + ([p["beta"]] if "beta" in p else []) + [p["log_scale"], p["log_noise_variance"], p["shape"]]
        chains.append(np.column_stack(cols))
    n = chains[0].shape[0]
    a = int(burn_in * n)
    samples = [c[a - 1:n] for c in chains]
    mm = len(samples)
    W = sum(np.cov(s, rowvar=False) for s in samples) / mm
    means = np.array([s.mean(axis=0) for s in samples])
    B = np.cov(means, rowvar=False)
    sv = np.linalg.svd(np.linalg.solve(W, B), compute_uv=False)[0]
    mpsrf = (n - 1) / n + (mm + 1) / mm * sv
    ind = ((mm + 1) / mm) * ((n - 1) / n) * (np.diag(B) / np.diag(W)) + (n + 1) / n
    return np.concatenate([[mpsrf], ind])


def run(lst, n_cycles=1, n_iterations_update=200, burn_in=.5, field_thinning=1.0, n_chromatic=10,
        Gelman_Rubin_Brooks_stop=(1.1, 1.1), sweep_form=0, verbose=False):
    """mcmc_nngp_run (run.R:8-48) with the chains of mcmc_nngp_update_Gaussian advanced one after another (each is seeded by
    iter_start + i, so the order does not matter)."""
    va, X = lst["vecchia_approx"], lst["X"]
    covfun = lst["space_time_model"]["stationary_covfun"]
    n_it = int(n_iterations_update)
    its = np.arange(1, n_it + 1)
    saved = its[np.round(its * field_thinning) == its * field_thinning]                                # run.R:26
    cycle = 1
    while cycle <= n_cycles:
        iter_start = lst["records"]["chain_1"]["iterations"][-1]
        for i, name in enumerate(lst["states"]):
            st = lst["states"][name]
            p = st["params"]
            cp = dict(shape=p["shape"], beta_0=p["beta_0"], log_scale=p["log_scale"], log_noise_variance=p["log_noise_variance"],
                      logvar_sufficient=st["transition_kernels"]["sufficient"], logvar_ancillary=st["transition_kernels"]["ancillary"])
            reg = None
            if X["X"] is not None:
                reg = dict(X=X["X"], xlocs=np.array(X["locs"], dtype=np.int32), first_obs=va["hctam_scol_1"],
                           solve_1XT1X=X["solve_1XT1X"], chol_solve_1XT1X=X["chol_solve_1XT1X"], beta=p["beta"])
            res = O.update_gaussian_chain(lst["locs"], va["NNarray"], va["coloring"], va["locs_match"], va["obs_per_loc"],
                                          lst["observed_field"], covfun, cp, p["field"], n_it, field_thinning, n_chromatic,
                                          iter_start, i + 1, sweep_form, regressors=reg)
            po, f, rec, frec = res[0], res[1], res[2], res[3]
            new = dict(shape=np.array(po["shape"]), beta_0=po["beta_0"], log_scale=po["log_scale"],
                       log_noise_variance=po["log_noise_variance"], field=f)
            out = dict(beta_0=rec[:, [0]], log_scale=rec[:, [1]], log_noise_variance=rec[:, [2]], shape=rec[:, 3:], field=frec)
            if reg is not None:
                new["beta"] = np.array(po["beta"]).copy()
                out["beta"] = res[5]
            st["params"] = new                                                                         # run.R:24
            st["transition_kernels"] = dict(sufficient=po["logvar_sufficient"], ancillary=po["logvar_ancillary"])
            r = lst["records"][name]
            r["saved_field"] = list(r["saved_field"]) + [iter_start + int(s) for s in saved]           # run.R:27
            r["iterations"] = list(r["iterations"]) + [iter_start + n_it]                              # run.R:28
            for k, v in out.items():                                                                   # run.R:29-32
                r["params"][k] = v if k not in r["params"] else np.vstack([r["params"][k], v])
        rhat = gelman_rubin_brooks(lst["records"], burn_in)                                            # run.R:38
        lst["diagnostics"].append(rhat)
        if verbose:
            print("cycle =", cycle, np.array2string(rhat, precision=6, floatmode="fixed"), flush=True)
        if rhat[0] < Gelman_Rubin_Brooks_stop[0] or np.all(rhat[1:] < Gelman_Rubin_Brooks_stop[1]):     # run.R:42-46
            break
        cycle += 1
    return lst


def get_summary(samples):
    """estimate.R:1-6: mean, quantile() type 7 at .025 / .5 / .975, sd (n-1)"""
    q = np.quantile(samples, [0.025, 0.5, 0.975], axis=0, method="linear")
    return np.column_stack([samples.mean(axis=0), q[0], q[1], q[2], samples.std(axis=0, ddof=1)])


def estimate(lst, burn_in=.5):
    """mcmc_nngp_estimate (estimate.R:9-96): GpGp_covparams, fixed_effects, field summaries"""
    recs = lst["records"]
    it = recs["chain_1"]["iterations"][-1]
    a = int(burn_in * it)
    cov = np.vstack([np.column_stack([r["params"]["log_scale"], r["params"]["log_noise_variance"], r["params"]["shape"]])[a - 1:it]
                     for r in recs.values()])                                                          # :22-32
    names = ["log_scale", "log_noise_variance"] + list(lst["space_time_model"]["shape_params"])
    g = cov.copy()
    for k, nm in enumerate(names):                                                                     # :37-38
        g[:, k] = np.exp(g[:, k]) if nm.startswith("log_") else 1.5 / (1.0 + np.exp(-g[:, k]))
    out = dict(sampled_covparams=get_summary(cov), GpGp_covparams=get_summary(g),
               covparams_names=[nm[4:] if nm.startswith("log_") else nm[7:] for nm in names])
    fx = []
    for r in recs.values():                                                                            # :72-78
        p = r["params"]
        o = np.column_stack([p["beta_0"]] + ([p["beta"]] if "beta" in p else []))[a - 1:it].copy()
        if o.shape[1] > 1:
            o[:, 0] = o[:, 0] - o[:, 1:] @ lst["X"]["X_mean"]
        fx.append(o)
    out["fixed_effects"] = get_summary(np.vstack(fx))
    sf = np.array(recs["chain_1"]["saved_field"])
    keep = sf > it * burn_in                                                                           # :90
    fl = [r["params"]["field"][keep] - r["params"]["beta_0"][sf[keep] - 1] for r in recs.values()]     # :91
    out["field"] = get_summary(np.vstack(fl))
    return out
