"""Host-side mirror of the reference's R entry points over the C ABI (the R originals are cited per function; paths relative
to /root/reference).  Same names, same argument meaning, same list-of-state schema (Python dicts instead of R lists, numpy
arrays instead of R matrices, 1-based index arrays kept 1-based exactly as R stores them).

R is not installed in this image, so this mirror is what drives configs 1, 2, 3 and 5 end to end; `R/` holds the equivalent
R glue.  With rng = "R" (mcmc_nngp_initialize and mcmc_nngp_run) every random draw is taken from R's own stream in the order
the reference takes it -- GpGp's randomised max-min ordering and jittered neighbour search included -- so a session gives the
numbers the reference gives in R: tests/test_gpu_vignette.py replays the reference's vignette and compares with what it prints.
"""
from __future__ import annotations

import time

import numpy as np

from . import _lib as L
from .context import NNGPContext, chains_run, find_ordered_nn, greedy_coloring, order_maxmin
from .rstream import RStream, find_ordered_nn_gpgp, order_maxmin_gpgp

SHAPE_PARAMS = {
    "exponential_isotropic": lambda d: ["log_range"],
    "exponential_sphere": lambda d: ["log_range"],
    "exponential_scaledim": lambda d: [f"log_range_{k + 1}" for k in range(d)],
    "exponential_spacetime": lambda d: ["log_range_1", "log_range_2"],
    "matern_isotropic": lambda d: ["log_range", "qlogis_smoothness"],
    "matern_sphere": lambda d: ["log_range", "qlogis_smoothness"],
    "matern_scaledim": lambda d: [f"log_range_{k + 1}" for k in range(d)] + ["qlogis_smoothness"],
    "matern_spacetime": lambda d: ["log_range_1", "log_range_2", "qlogis_smoothness"],
}


def plogis(x):
    return 1.0 / (1.0 + np.exp(-x))


def shape_to_covparms(shape, shape_params, smooth_transform):
    """c(1, shape, 0) with "log*" -> exp and "qlogis*" -> the calling file's own transform (the reference uses
    .4+.7 plogis in initialize.R:199, .5+.5 plogis in update_Gaussian.R:70,121,177 and 1.5 plogis in predict.R:37)."""
    out = [1.0]
    for v, name in zip(np.atleast_1d(shape), shape_params):
        out.append(float(np.exp(v)) if name.startswith("log") else float(smooth_transform(v)))
    out.append(0.0)
    return out


def _model_matrix(Xin):
    """model.matrix(~., X)[,-1]: numeric columns as they are, factor columns as treatment-contrast dummies (first level
    dropped).  Accepts a numpy array or a pandas DataFrame.  (initialize.R:124-127)"""
    if Xin is None:
        return None, []
    try:
        import pandas as pd
    except Exception:  # pragma: no cover
        pd = None
    if pd is not None and isinstance(Xin, pd.DataFrame):
        cols, names = [], []
        for name in Xin.columns:
            col = Xin[name]
            if col.dtype == object or str(col.dtype) == "category" or col.dtype == bool:
                cat = col.astype("category")
                for lev in list(cat.cat.categories)[1:]:
                    cols.append((cat == lev).to_numpy(dtype=np.float64))
                    names.append(f"{name}{lev}")
            else:
                cols.append(col.to_numpy(dtype=np.float64))
                names.append(str(name))
        return np.column_stack(cols) if cols else np.zeros((len(Xin), 0)), names
    A = np.asarray(Xin, dtype=np.float64)
    if A.ndim == 1:
        A = A[:, None]
    return A, [f"V{k + 1}" for k in range(A.shape[1])]


# ---------------------------------------------------------------------------------------------------------------------
# mcmc_nngp_initialize  (Scripts/mcmc_nngp_initialize.R:1-240)
# ---------------------------------------------------------------------------------------------------------------------
def mcmc_nngp_initialize(observed_locs, observed_field, X_obs=None, X_locs=None, m=10, reordering="maxmin",
                         stationary_covfun="exponential_isotropic", response_model="Gaussian", n_chains=3, seed=1,
                         device=0, build_adjacency=None, rng="numpy"):
    """rng: see _initialize_host, which prepares everything on the host; the initial fields (:201-208: factor build +
    sqrt(scale) * solve(Linv, z)) are drawn here on the device."""
    lst, pending = _initialize_host(observed_locs, observed_field, X_obs, X_locs, m, reordering, stationary_covfun, response_model,
                                    n_chains, seed, build_adjacency, rng)
    va = lst["vecchia_approx"]
    ctx = NNGPContext(lst["locs"], va["NNarray"], va["coloring"], va["locs_match"], stationary_covfun, device=device)
    try:
        for name, (cp, z) in pending.items():
            params = lst["states"][name]["params"]
            ctx.factor_build(cp)                                                                      # :201
            ctx.field_init(params["beta_0"], params["log_scale"], z)                                  # :202-208
            params["field"] = ctx.field_get()
            lst["records"][name]["iterations"][0, 1] = time.time() - lst["t_begin"]
    finally:
        ctx.close()
    print(f"Setup done, {time.time() - lst['t_begin']:.3f} s elapsed")                                # :236
    return lst


def _initialize_host(observed_locs, observed_field, X_obs, X_locs, m, reordering, stationary_covfun, response_model, n_chains, seed,
                     build_adjacency, rng):
    """Everything of mcmc_nngp_initialize that the reference does on the host too: ordering, vecchia_approx, regressors, the
    scalar starting values of every chain.  Returns (list, pending) with pending[chain] = (covparms, z): the factor parameters
    and the normals of that chain's initial field draw (:201-208), which mcmc_nngp_initialize carries out on the device.
    rng = "numpy": numpy's generator for the initial states, the exact farthest-point ordering for "maxmin" (fast at any n).
    rng = "R": R's stream after set.seed(seed) (:17), consumed exactly as the reference consumes it -- GpGp::order_maxmin
    (:29: jitter rnorm, sample(n)), GpGp::find_ordered_nn (:93: jitter rnorm), sample(., 1) per chain (:154-161), then per chain
    rnorm(p + 1), rbeta, rbeta, rnorm(n) (:189-208) -- so the returned list equals the one R returns."""
    t_begin = time.time()
    if rng not in ("numpy", "R"):
        raise ValueError(f"unknown rng {rng!r}")
    rs = RStream(seed) if rng == "R" else None                                                       # :17 set.seed(seed)
    gen = np.random.default_rng(seed)
    observed_locs = np.asarray(observed_locs, dtype=np.float64)
    if observed_locs.ndim == 1:
        observed_locs = observed_locs[:, None]
    observed_field = np.asarray(observed_field, dtype=np.float64).ravel()
    # ---- de-duplicate (first occurrence kept, original order) and re-order: initialize.R:26-34
    _, first, inverse = np.unique(observed_locs, axis=0, return_index=True, return_inverse=True)
    inverse = np.asarray(inverse).ravel()
    locs = observed_locs[np.sort(first)]
    rkind = reordering if isinstance(reordering, str) else reordering[0]
    if rkind == "maxmin":
        if rs is not None:
            order = order_maxmin_gpgp(locs, rs, lonlat="sphere" in stationary_covfun) - 1             # :29, GpGp's own ordering
        else:
            order = order_maxmin(locs) - 1          # exact farthest-point; GpGp::order_maxmin is a randomised approximation
    elif rkind == "random":
        order = rs.sample_int(locs.shape[0]) - 1 if rs is not None else gen.permutation(locs.shape[0])   # :30
    elif rkind == "coord":
        order = np.argsort(locs[:, int(reordering[1]) - 1], kind="stable")
    elif rkind == "dist_to_point":
        order = np.argsort(((locs - np.asarray(reordering[1], dtype=float)) ** 2).sum(1), kind="stable")
    elif rkind == "middleout":
        order = np.argsort(((locs - locs.mean(0)) ** 2).sum(1), kind="stable")
    elif rkind == "none":
        order = np.arange(locs.shape[0])
    else:
        raise ValueError(f"unknown reordering {reordering!r}")
    locs = locs[order]
    n, d = locs.shape
    if stationary_covfun not in SHAPE_PARAMS:
        raise ValueError(f"unknown stationary_covfun {stationary_covfun!r}")
    shape_params = SHAPE_PARAMS[stationary_covfun](d)
    space_time_model = {"response_model": response_model, "covfun": {"stationary_covfun": stationary_covfun, "shape_params": shape_params}}

    # ---- Vecchia approximation: initialize.R:80-110 (all index arrays 1-based like R)
    va = {"n_locs": n, "n_obs": observed_field.size}
    # match(observed_locs, locs) without a Python loop: unique row -> its rank in first-occurrence order -> its place after reordering
    rank = np.empty(first.size, dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(first.size)
    place = np.empty(n, dtype=np.int64)
    place[order] = np.arange(n)
    locs_match = (place[rank[inverse]] + 1).astype(np.int32)                                         # :85
    va["locs_match"] = locs_match
    order_obs = np.argsort(locs_match, kind="stable")
    counts = np.bincount(locs_match - 1, minlength=n)
    ptr = np.concatenate([[0], np.cumsum(counts)])
    va["hctam_scol"] = np.split(order_obs + 1, ptr[1:-1])                                            # :88
    va["hctam_scol_1"] = (order_obs[ptr[:-1]] + 1).astype(np.int32)                                  # :89
    va["obs_per_loc"] = counts.astype(np.float64)                                                    # :91
    NN = find_ordered_nn_gpgp(locs, m, rs) if rs is not None else find_ordered_nn(locs, m)           # :93 (no lonlat: quirk 3)
    va["NNarray"] = NN
    non_na = NN != L.NA_INT
    va["NNarray_non_NA"] = non_na                                                                    # :97
    va["sparse_chol_column_idx"] = NN.T[non_na.T]                                                    # :99  column-major scan
    va["sparse_chol_row_idx"] = np.tile(np.arange(1, n + 1, dtype=np.int32)[:, None], (1, m + 1)).T[non_na.T]   # :101
    if build_adjacency is None:
        build_adjacency = n <= 200_000
    if build_adjacency:                                                                              # :103-109
        import scipy.sparse as sp
        A = sp.csr_matrix((np.ones(va["sparse_chol_row_idx"].size), (va["sparse_chol_row_idx"] - 1, va["sparse_chol_column_idx"] - 1)), shape=(n, n))
        M = (A.T @ A).tocsc()
        M.data[:] = 1.0
        va["MRF_adjacency_mat"] = M
    else:
        va["MRF_adjacency_mat"] = None   # pattern(A^T A) is never needed by the device path; skipped for very large n
    va["coloring"] = greedy_coloring(NN)                                                             # :110, Coloring.R:2-20

    # ---- regressors: initialize.R:116-137
    X = {"arg": {"X_locs": X_locs, "X_obs": X_obs}, "X": None, "locs": [], "names": []}
    Xl, nl = _model_matrix(X_locs)
    Xo, no_ = _model_matrix(X_obs)
    parts = [a for a in (Xl, Xo) if a is not None]
    if parts:
        XX = np.column_stack(parts)
        X["names"] = nl + no_
        # reference quirk: X$locs = seq(ncol(X_locs)) counts the ORIGINAL columns of X_locs (factors not yet expanded)
        n_loc_cols = 0 if X_locs is None else (np.asarray(X_locs).shape[1] if np.asarray(X_locs).ndim > 1 else 1)
        X["locs"] = list(range(min(n_loc_cols, XX.shape[1])))
        X["X_mean"] = XX.mean(0)
        XX = XX - X["X_mean"]
        X["X"] = XX
        X["solve_XTX"] = np.linalg.inv(XX.T @ XX)
        one = np.column_stack([np.ones(XX.shape[0]), XX])
        X["solve_1XT1X"] = np.linalg.inv(one.T @ one)
        X["chol_solve_XTX"] = np.linalg.cholesky(X["solve_XTX"]).T        # R's chol() is the upper factor
        X["chol_solve_1XT1X"] = np.linalg.cholesky(X["solve_1XT1X"]).T

    # ---- chain states: initialize.R:143-210
    design = np.ones((observed_field.size, 1)) if X["X"] is None else np.column_stack([np.ones(observed_field.size), X["X"]])
    coef, *_ = np.linalg.lstsq(design, observed_field, rcond=None)
    resid = observed_field - design @ coef
    dof = max(observed_field.size - design.shape[1], 1)
    vcov = np.linalg.inv(design.T @ design) * (resid @ resid) / dof
    var_resid = resid.var(ddof=1)
    head = locs[:100]

    def max_dist(P):
        return float(np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)).max())

    def draw_shape():                                                                                # :154-161
        shape = []
        for name in shape_params:
            if name.startswith("log_range"):
                if "scaledim" in stationary_covfun:
                    P = head[:, [int(name.split("_")[-1]) - 1]]
                elif "spacetime" in stationary_covfun:
                    P = head[:, :-1] if name.endswith("_1") else head[:, [-1]]
                else:
                    P = head
                if rs is not None:
                    shape.append(rs.sample_one(np.log(max_dist(P)) - np.log(np.arange(20.0, 201.0))))
                else:
                    shape.append(np.log(max_dist(P)) - np.log(gen.integers(20, 201)))
            else:
                shape.append(rs.rnorm(1)[0] if rs is not None else gen.standard_normal())
        return np.array(shape)

    normal = (lambda k: rs.rnorm(k)) if rs is not None else (lambda k: gen.standard_normal(k))
    beta_10_10 = (lambda: rs.rbeta(1, 10, 10)[0]) if rs is not None else (lambda: gen.beta(10, 10))
    # the reference draws the shapes of ALL chains first (:152-162), then the rest chain by chain (:181-209)
    shapes = [draw_shape() for _ in range(n_chains)] if rs is not None else None
    states, records, pending = {}, {}, {}
    for i in range(n_chains):
        shape = shapes[i] if shapes is not None else draw_shape()
        perturb = np.linalg.cholesky(vcov) @ normal(coef.size)                                        # :189
        params = {"shape": shape, "beta_0": float(coef[0] + perturb[0])}
        if X["X"] is not None:
            params["beta"] = coef[1:] + perturb[1:]
        params["log_scale"] = float(np.log(beta_10_10() * var_resid))                                 # :193
        params["log_noise_variance"] = float(np.log(beta_10_10() * var_resid))                        # :194
        cp = shape_to_covparms(shape, shape_params, lambda v: .4 + .7 * plogis(v))                    # :196-200
        pending[f"chain_{i + 1}"] = (cp, normal(n))                                                   # z of :208
        states[f"chain_{i + 1}"] = {
            "transition_kernels": {"covariance_params_sufficient": {"logvar": -2.0}, "covariance_params_ancillary": {"logvar": -2.0},
                                   "log_noise_variance": {"logvar": -1.0}},                          # :184-187
            "params": params}
        records[f"chain_{i + 1}"] = {"iterations": np.array([[0.0, time.time() - t_begin]]), "params": {}}   # :225-227
    return {"locs": locs, "X": X, "observed_field": observed_field, "observed_locs": observed_locs,
            "space_time_model": space_time_model, "vecchia_approx": va, "states": states, "records": records,
            "diagnostics": {"Gelman_Rubin_Brooks": []}, "t_begin": t_begin, "seed": seed}, pending


# ---------------------------------------------------------------------------------------------------------------------
# mcmc_nngp_update_Gaussian  (Scripts/mcmc_nngp_update_Gaussian.R:14-319)
# ---------------------------------------------------------------------------------------------------------------------
def _chain_context(vecchia_approx, locs, covfun, chain, device):
    cache = vecchia_approx.setdefault("_b200_contexts", {})
    k = (chain, device)
    if k not in cache:
        cache[k] = NNGPContext(locs, vecchia_approx["NNarray"], vecchia_approx["coloring"], vecchia_approx["locs_match"], covfun, device=device)
    return cache[k]


def release_contexts(mcmc_nngp_list):
    """frees the device-side copies of vecchia_approx (one per chain) held between cycles"""
    for ctx in mcmc_nngp_list["vecchia_approx"].pop("_b200_contexts", {}).values():
        ctx.close()


def mcmc_nngp_update_Gaussian(locs, X, observed_field, space_time_model, vecchia_approx, states, n_iterations_update,
                              n_cores=None, field_thinning=1, ancillary=True, n_chromatic=10, iterations=None, n_gpus=None,
                              rng="philox", regressor_engine="device"):
    """Returns [ {"state": ..., "records": ...} per chain ] like the reference (:315).  `ancillary` is accepted and ignored,
    as in the reference (quirk 1).  Chains are not forked: chain i lives on GPU i mod n_gpus inside this process and all chains
    advance concurrently behind one nngp_chains_run call (at most n_cores in flight), like the reference's mclapply (:22-26).
    regressor_engine (models with X): "device" = the whole loop behind nngp_chain_run_regressors (X resident in HBM);
    "host" = the loop driven from Python over the device primitives (kept as a cross-check of the former)."""
    n_dev = L.device_count()
    if n_dev < 1:
        raise L.NNGPError(2, "no CUDA device: libnngp_b200 has no CPU fallback")
    n_gpus = n_dev if n_gpus is None else max(1, min(n_gpus, n_dev))
    iter_start = int(iterations[-1, 0])
    covfun = space_time_model["covfun"]["stationary_covfun"]
    shape_params = space_time_model["covfun"]["shape_params"]
    var_y = float(np.var(observed_field, ddof=1))
    names = list(states.keys())
    ctxs = [_chain_context(vecchia_approx, locs, covfun, i, i % n_gpus) for i in range(len(names))]
    if X["X"] is not None and regressor_engine == "host":
        return [_update_chain_regressors(ctxs[i], states[nm], X, observed_field, vecchia_approx, shape_params, var_y, n_iterations_update,
                                         field_thinning, n_chromatic, iter_start, i + 1) for i, nm in enumerate(names)]
    # all chains advance concurrently behind ONE call (nngp_chains_run: a host thread and a stream per chain; chain i on GPU
    # i mod n_gpus), as mclapply does in the reference (:22-26); n_cores bounds the number in flight
    params_list = []
    for ctx, nm in zip(ctxs, names):
        p, tk = states[nm]["params"], states[nm]["transition_kernels"]
        if X["X"] is None:
            ctx.obs_set(observed_field)
        elif not getattr(ctx, "_regressors_loaded", False):
            ctx.regressors_set(X["X"], observed_field, xlocs=[int(c) + 1 for c in X["locs"]], first_obs=vecchia_approx["hctam_scol_1"])
            ctx._regressors_loaded = True
        ctx.field_set(p["field"])
        params_list.append({"shape": p["shape"], "beta_0": p["beta_0"], "log_scale": p["log_scale"], "log_noise_variance": p["log_noise_variance"],
                            "logvar_sufficient": tk["covariance_params_sufficient"]["logvar"],
                            "logvar_ancillary": tk["covariance_params_ancillary"]["logvar"]})
    rng_mode = L.RNG_SUPPLIED if rng == "R" else L.RNG_PHILOX
    kw = dict(thin=field_thinning, n_chromatic=n_chromatic, iter_start=iter_start, chain_indices=list(range(1, len(names) + 1)),
              rng_mode=rng_mode, max_concurrent=n_cores)
    if X["X"] is None:
        res = chains_run(ctxs, params_list, n_iterations_update, var_y, **kw)
    else:
        res = chains_run(ctxs, params_list, n_iterations_update, var_y, betas=[states[nm]["params"]["beta"] for nm in names],
                         solve_1XT1X=X["solve_1XT1X"], chol_solve_1XT1X=X["chol_solve_1XT1X"], **kw)
    out = []
    for ctx, nm, r in zip(ctxs, names, res):
        tk = states[nm]["transition_kernels"]
        po, rec, frec = r[0], r[1], r[-2]
        new_params = {"shape": po["shape"], "beta_0": po["beta_0"], "log_scale": po["log_scale"],
                      "log_noise_variance": po["log_noise_variance"], "field": ctx.field_get()}
        brec = None
        if X["X"] is not None:
            new_params["beta"] = po["beta"]
            brec = r[2]
        new_state = {"transition_kernels": {"covariance_params_sufficient": {"logvar": po["logvar_sufficient"]},
                                            "covariance_params_ancillary": {"logvar": po["logvar_ancillary"]},
                                            "log_noise_variance": dict(tk["log_noise_variance"])},
                     "params": new_params}
        out.append({"state": new_state, "records": _records_dict(rec, shape_params, frec, beta=brec, beta_names=X["names"] if X["X"] is not None else None)})
    return out


def _records_dict(rec, shape_params, field_records, beta=None, beta_names=None):
    # in the order the reference creates them (update_Gaussian.R:42-56): it is the column order of the diagnostics
    r = {"beta_0": rec[:, [0]].copy()}
    if beta is not None:
        r["beta"] = beta
    r.update({"log_scale": rec[:, [1]].copy(), "log_noise_variance": rec[:, [2]].copy(),
              "shape": rec[:, 3:3 + len(shape_params)].copy(), "field": field_records})
    r["_shape_names"] = list(shape_params)
    r["_beta_names"] = list(beta_names or [])
    return r


def _update_chain_regressors(ctx, state, X, y, va, shape_params, var_y, n_iter, thin, n_chromatic, iter_start, chain_index):
    """Regressor model: the loop of :101-314 driven from the host over the device primitives (factor build, log-lik,
    ancillary proposal, sweeps, SSR stay on the GPU; the (p+1)-dimensional regression algebra of :226-246 is host numpy)."""
    rng = np.random.default_rng(iter_start + chain_index)                                            # :36 (numpy stream, not R's)
    p = {k: (np.array(v, dtype=float) if isinstance(v, np.ndarray) else v) for k, v in state["params"].items()}
    tk = {k: dict(v) for k, v in state["transition_kernels"].items()}
    XX = X["X"]
    xl = list(X["locs"])
    n, n_obs, ns = va["n_locs"], va["n_obs"], len(shape_params)
    lm = va["locs_match"] - 1
    first = va["hctam_scol_1"] - 1
    smooth = lambda v: .5 + .5 * plogis(v)
    one_X = np.column_stack([np.ones(n_obs), XX])
    Xl_sites = XX[first][:, xl] if xl else None
    rec = np.zeros((n_iter, 3 + ns))
    rec_beta = np.zeros((n_iter, XX.shape[1]))
    n_frec = int(round(n_iter * thin))
    frec = np.zeros((n_frec, n))
    acc_s, acc_a = np.zeros(n_iter + 1), np.zeros(n_iter + 1)

    field = p["field"].copy()
    ctx.factor_build(shape_to_covparms(p["shape"], shape_params, smooth), L.SLOT_CURRENT)              # :72
    ctx.factor_commit()

    def interweave_matrices():                                                                        # :77-83, 145-151, 200-206
        B = np.column_stack([ctx.spmv(np.ones(n))] + [ctx.spmv(Xl_sites[:, k]) for k in range(len(xl))])
        prec = B.T @ B
        cov = np.linalg.inv(prec)
        return B, cov, np.linalg.cholesky(cov)

    if xl:
        scXl, cov_iw, chol_iw = interweave_matrices()
    mu = p["beta_0"] + XX @ p["beta"]
    for it in range(1, n_iter + 1):
        ctx.field_set(field)
        ctx.obs_set(y - (mu - p["beta_0"]))
        # ---- (A) ancillary :113-157
        inn = rng.standard_normal(ns + 1) * np.exp(.5 * tk["covariance_params_ancillary"]["logvar"])
        new_ls, new_shape = p["log_scale"] + inn[0], p["shape"] + inn[1:]
        bad = ctx.factor_build(shape_to_covparms(new_shape, shape_params, smooth), L.SLOT_PROPOSAL)
        ratio = ctx.ancillary_propose(p["beta_0"], new_ls - p["log_scale"], p["log_noise_variance"])
        if bad == 0 and ratio > np.log(rng.random()):
            ctx.ancillary_accept()
            p["shape"], p["log_scale"] = new_shape, new_ls
            field = ctx.field_get()
            acc_a[it] = 1
            if xl:
                scXl, cov_iw, chol_iw = interweave_matrices()
        if 0 <= iter_start <= 2000 and it % 25 == 0:
            a = acc_a[it - 24:it + 1].mean()
            if a < .05:
                tk["covariance_params_ancillary"]["logvar"] -= rng.normal(.4, .05)
            if a > .15:
                tk["covariance_params_ancillary"]["logvar"] += rng.normal(.4, .05)
        # ---- (B) sufficient :165-213
        inn = rng.standard_normal(ns + 1) * np.exp(.5 * tk["covariance_params_sufficient"]["logvar"])
        new_ls = p["log_scale"] + inn[0]
        if np.exp(new_ls) < var_y:
            new_shape = p["shape"] + inn[1:]
            bad = ctx.factor_build(shape_to_covparms(new_shape, shape_params, smooth), L.SLOT_PROPOSAL)
            gp_ratio = ctx.loglik(p["beta_0"], new_ls, L.SLOT_PROPOSAL) - ctx.loglik(p["beta_0"], p["log_scale"], L.SLOT_CURRENT)
            if bad == 0 and gp_ratio > np.log(rng.random()):
                ctx.factor_accept()
                p["shape"], p["log_scale"] = new_shape, new_ls
                acc_s[it] = 1
                if xl:
                    scXl, cov_iw, chol_iw = interweave_matrices()
        if 0 <= iter_start <= 2000 and it % 25 == 0:
            a = acc_s[it - 24:it + 1].mean()
            if a < .05:
                tk["covariance_params_sufficient"]["logvar"] -= rng.normal(.2, .05)
            if a > .15:
                tk["covariance_params_sufficient"]["logvar"] += rng.normal(.2, .05)
        # ---- (C) mean parameters :219-250
        if not xl:
            bm, bv = ctx.beta0_moments(p["log_scale"])
            p["beta_0"] = bm + np.sqrt(bv) * rng.standard_normal()
        beta_mean = (y - field[lm] + p["beta_0"]) @ one_X @ X["solve_1XT1X"]                          # :229
        innov = beta_mean + np.exp(.5 * p["log_noise_variance"]) * (X["chol_solve_1XT1X"].T @ rng.standard_normal(XX.shape[1] + 1))
        field = field - p["beta_0"] + innov[0]                                                        # :232
        p["beta_0"], p["beta"] = float(innov[0]), innov[1:].copy()
        if xl:                                                                                        # :237-246
            other = field + Xl_sites @ p["beta"][xl]
            bmean = cov_iw @ (scXl.T @ ctx.spmv(other))
            innov = bmean + np.exp(.5 * p["log_scale"]) * (chol_iw @ rng.standard_normal(len(xl) + 1))
            p["beta_0"] = float(innov[0])
            p["beta"][xl] = innov[1:]
            field = other - Xl_sites @ p["beta"][xl]
        mu = p["beta_0"] + XX @ p["beta"]
        # ---- (D) chromatic sweeps :257-275
        ctx.field_set(field)
        ctx.obs_set(y - (mu - p["beta_0"]))
        ctx.gibbs_sweep(p["beta_0"], p["log_scale"], p["log_noise_variance"], n_sweeps=n_chromatic, seed=iter_start * 1009 + chain_index)
        # ---- (E) noise variance :281-293
        ssr = ctx.ssr()
        for _ in range(10):
            d = rng.normal(0, .01)
            if np.exp(p["log_noise_variance"] + d) < var_y:
                if -.5 * n_obs * d - .5 * ssr * (np.exp(-p["log_noise_variance"] - d) - np.exp(-p["log_noise_variance"])) > np.log(rng.random()):
                    p["log_noise_variance"] += d
        field = ctx.field_get()
        # ---- (F) records :305-311
        rec[it - 1, :3] = (p["beta_0"], p["log_scale"], p["log_noise_variance"])
        rec[it - 1, 3:] = p["shape"]
        rec_beta[it - 1] = p["beta"]
        t = it * thin
        if round(t) == t and 1 <= int(t) <= n_frec:
            frec[int(t) - 1] = field
    p["field"] = field
    return {"state": {"transition_kernels": tk, "params": p},
            "records": _records_dict(rec, shape_params, frec, beta=rec_beta, beta_names=X["names"])}


# ---------------------------------------------------------------------------------------------------------------------
# diagnostics  (Scripts/mcmc_nngp_diagnose.R:1-25, 108-122)
# ---------------------------------------------------------------------------------------------------------------------
def _scalar_samples(chain, burn_in):
    cols = [chain["params"][k] for k in chain["params"] if k not in ("field",) and not k.startswith("_")]
    M = np.column_stack(cols)
    n = M.shape[0]
    return M[max(int(np.floor(burn_in * n)) - 1, 0):n]


def Gelman_Rubin_Brooks(records, burn_in=.5):
    samples = [_scalar_samples(c, burn_in) for c in records.values()]
    m, n = len(samples), next(iter(records.values()))["params"]["beta_0"].shape[0]
    W = sum(np.atleast_2d(np.cov(s.T)) for s in samples) / m
    means = np.array([s.mean(0) for s in samples])
    B = np.atleast_2d(np.cov(means.T)) if m > 1 else np.zeros_like(W)
    try:
        mpsrf = (n - 1) / n + (m + 1) / m * np.linalg.svd(np.linalg.solve(W, B), compute_uv=False)[0]
    except np.linalg.LinAlgError:
        mpsrf = np.inf
    with np.errstate(divide="ignore", invalid="ignore"):
        ind = ((m + 1) / m) * ((n - 1) / n) * (np.diag(B) / np.diag(W)) + (n + 1) / n
    return {"R_hat": np.concatenate([[mpsrf], ind]), "within_variance": W}


def _ess_1d(x):
    """initial-positive-sequence estimator (coda::effectiveSize fits an AR spectrum instead; same quantity)"""
    x = np.asarray(x, dtype=float) - np.mean(x)
    n = x.size
    if n < 4 or np.allclose(x, 0):
        return float(n)
    f = np.fft.rfft(x, 2 * n)
    ac = np.fft.irfft(f * np.conj(f))[:n] / (x @ x)
    s = 0.0
    for k in range(1, n - 1, 2):
        pair = ac[k] + ac[k + 1]
        if pair < 0:
            break
        s += pair
    return float(n / max(1.0 + 2.0 * s, 1e-12))


def ESS(records, burn_in=.5):
    E = np.array([[_ess_1d(col) for col in _scalar_samples(c, burn_in).T] for c in records.values()])
    return np.vstack([E, E.sum(0)])


# ---------------------------------------------------------------------------------------------------------------------
# mcmc_nngp_run  (Scripts/mcmc_nngp_run.R:1-52)
# ---------------------------------------------------------------------------------------------------------------------
def mcmc_nngp_run(mcmc_nngp_list, Gelman_Rubin_Brooks_stop=(1.1, 1.1), burn_in=.5, n_cores=None, field_thinning=1,
                  n_iterations_update=200, ancillary=True, n_chromatic=10, save_name=None, n_cycles=1, plot_beta=False,
                  n_gpus=None, rng="philox", verbose=True):
    cycle = 1
    while cycle <= n_cycles:
        if verbose:
            print(f"cycle = {cycle}")
        res = mcmc_nngp_update_Gaussian(mcmc_nngp_list["locs"], mcmc_nngp_list["X"], mcmc_nngp_list["observed_field"],
                                        mcmc_nngp_list["space_time_model"], mcmc_nngp_list["vecchia_approx"], mcmc_nngp_list["states"],
                                        n_iterations_update, n_cores, field_thinning, ancillary, n_chromatic,
                                        iterations=next(iter(mcmc_nngp_list["records"].values()))["iterations"], n_gpus=n_gpus, rng=rng)
        its = np.arange(1, n_iterations_update + 1)
        saved = its[np.round(its * field_thinning) == its * field_thinning]                          # :26
        for i, name in enumerate(mcmc_nngp_list["records"].keys()):
            r = mcmc_nngp_list["records"][name]
            mcmc_nngp_list["states"][name] = res[i]["state"]                                          # :24
            iter_start = r["iterations"][-1, 0]
            r["saved_field"] = np.concatenate([r.get("saved_field", np.zeros(0)), iter_start + saved])   # :27
            r["iterations"] = np.vstack([r["iterations"], [iter_start + n_iterations_update, time.time() - mcmc_nngp_list["t_begin"]]])
            for k, v in res[i]["records"].items():                                                    # :29-32 (rbind)
                if k.startswith("_"):
                    r["params"][k] = v
                else:
                    r["params"][k] = v if k not in r["params"] else np.vstack([r["params"][k], v])
        grb = Gelman_Rubin_Brooks(mcmc_nngp_list["records"], burn_in)                                 # :38
        mcmc_nngp_list["diagnostics"]["Gelman_Rubin_Brooks"].append(grb)
        mcmc_nngp_list["diagnostics"].setdefault("ESS", []).append(ESS(mcmc_nngp_list["records"], burn_in))
        if verbose:
            print("Gelman-Rubin-Brooks R-hat : ", np.round(grb["R_hat"], 4))
        if grb["R_hat"][0] < Gelman_Rubin_Brooks_stop[0] or np.all(grb["R_hat"][1:] < Gelman_Rubin_Brooks_stop[1]):   # :42-43 (OR, quirk 11)
            break
        cycle += 1
    return mcmc_nngp_list


# ---------------------------------------------------------------------------------------------------------------------
# mcmc_nngp_estimate  (Scripts/mcmc_nngp_estimate.R:1-100)
# ---------------------------------------------------------------------------------------------------------------------
def get_summary(samples):
    """mean, 2.5 / 50 / 97.5 % quantiles (R type 7 = numpy 'linear'), sd (n-1)   (estimate.R:1-6)"""
    S = np.asarray(samples, dtype=float)
    if S.ndim == 1:
        S = S[:, None]
    q = np.quantile(S, [0.025, 0.5, 0.975], axis=0)
    return np.column_stack([S.mean(0), q[0], q[1], q[2], S.std(0, ddof=1)])


SUMMARY_COLUMNS = ["mean", "q0.025", "median", "q0.975", "sd"]


def mcmc_nngp_estimate(mcmc_nngp_list, burn_in=.5):
    recs = mcmc_nngp_list["records"]
    first = next(iter(recs.values()))
    it = int(first["iterations"][-1, 0])
    lo = max(int(burn_in * it), 1)
    shape_names = first["params"]["_shape_names"]
    names = ["log_scale", "log_noise_variance"] + list(shape_names)
    samples = np.vstack([np.column_stack([c["params"]["log_scale"], c["params"]["log_noise_variance"], c["params"]["shape"]])[lo - 1:it] for c in recs.values()])
    res = {"covariance_params": {"sampled_covparams": {"names": names, "summary": get_summary(samples)}}}
    g = samples.copy()                                                                                # :34-46 GpGp parametrisation
    gnames = []
    for k, nm in enumerate(names):
        if nm.startswith("log_"):
            g[:, k] = np.exp(g[:, k]); gnames.append(nm[4:])
        elif nm.startswith("qlogis_"):
            g[:, k] = 1.5 * plogis(g[:, k]); gnames.append(nm[7:])
        else:
            gnames.append(nm)
    res["covariance_params"]["GpGp_covparams"] = {"names": gnames, "summary": get_summary(g)}
    inla = g.copy()                                                                                   # :48-65 INLA parametrisation
    inames = list(gnames)
    covfun = mcmc_nngp_list["space_time_model"]["covfun"]["stationary_covfun"]
    rcols = [k for k, nm in enumerate(gnames) if "range" in nm]
    if "exponential" in covfun:
        inla[:, rcols] *= 2
    keep = list(range(len(inames)))
    if "matern" in covfun:
        sm = gnames.index("smoothness")
        inla[:, rcols] *= np.sqrt(8 * inla[:, [sm]])
        keep.remove(sm)
    for k, nm in enumerate(inames):
        if "noise" in nm:
            inla[:, k] = 1.0 / inla[:, k]; inames[k] = "precision_of_Gaussian_obs"
        elif "scale" in nm:
            inla[:, k] = np.sqrt(inla[:, k]); inames[k] = "sd_for_spatial"
    res["covariance_params"]["INLA_covparams"] = {"names": [inames[k] for k in keep], "summary": get_summary(inla[:, keep])}
    # fixed effects :70-80 (beta_0 de-centred with X_mean)
    fx = []
    for c in recs.values():
        cols = [c["params"]["beta_0"]] + ([c["params"]["beta"]] if "beta" in c["params"] else [])
        o = np.column_stack(cols)[lo - 1:it]
        if o.shape[1] > 1:
            o[:, 0] = o[:, 0] - o[:, 1:] @ mcmc_nngp_list["X"]["X_mean"]
        fx.append(o)
    fs = get_summary(np.vstack(fx))
    res["fixed_effects"] = {"names": ["beta_0"] + list(first["params"].get("_beta_names", [])), "summary": fs,
                            "zero_out_of_ci": np.sign(fs[:, 1]) * np.sign(fs[:, 3]) > 0}
    # field :88-94 (stored samples after burn-in, minus beta_0 of the same iteration)
    saved = first["saved_field"]
    sel = saved > it * burn_in
    fsamp = np.vstack([c["params"]["field"][sel] - c["params"]["beta_0"][saved[sel].astype(int) - 1] for c in recs.values()])
    res["field"] = get_summary(fsamp)
    return res


# ---------------------------------------------------------------------------------------------------------------------
# prediction  (Scripts/mcmc_nngp_predict.R:1-104)
# ---------------------------------------------------------------------------------------------------------------------
def mcmc_nngp_predict_field(mcmc_nngp_list, predicted_locs, burn_in=.5, n_cores=1, m=10, device=0, seed=0):
    predicted_locs = np.asarray(predicted_locs, dtype=np.float64)
    if predicted_locs.ndim == 1:
        predicted_locs = predicted_locs[:, None]
    locs = np.vstack([mcmc_nngp_list["locs"], predicted_locs])                                        # :4
    NN = find_ordered_nn(locs, m)                                                                     # :5
    stm = mcmc_nngp_list["space_time_model"]
    n, n_pred = mcmc_nngp_list["vecchia_approx"]["n_locs"], predicted_locs.shape[0]
    first = next(iter(mcmc_nngp_list["records"].values()))
    stored = first["saved_field"]
    stored = stored[stored > burn_in * stored.max()].astype(int)                                      # :13-14
    rng = np.random.default_rng(seed)
    out = {}
    with NNGPContext(locs, NN, np.zeros(n + n_pred, dtype=np.int32), np.zeros(0, dtype=np.int32), stm["covfun"]["stationary_covfun"], device=device) as ctx:
        for name, chain in mcmc_nngp_list["records"].items():
            samples = np.zeros((stored.size, n_pred))
            prev_shape = None
            saved = chain["saved_field"].astype(int)
            for k, i_chain in enumerate(stored):
                i_field = int(np.where(saved == i_chain)[0][0])                                       # :30
                shape = chain["params"]["shape"][i_chain - 1]
                if prev_shape is None or not np.array_equal(shape, prev_shape):                       # :23,32 (rebuild only when the shape changed)
                    ctx.factor_build(shape_to_covparms(shape, stm["covfun"]["shape_params"], lambda v: 1.5 * plogis(v)))   # :37 quirk 2
                    prev_shape = shape.copy()
                samples[k] = ctx.predict_sample(n, chain["params"]["field"][i_field], float(chain["params"]["beta_0"][i_chain - 1, 0]),
                                                float(chain["params"]["log_scale"][i_chain - 1, 0]), rng.standard_normal(n_pred))   # :43-53
            out[name] = samples
    summary = get_summary(np.vstack(list(out.values())))                                              # :57-58
    return {"predicted_locs": predicted_locs, "predicted_field_samples": out, "predicted_field_summary": summary}


mcmc_nngp_predict = mcmc_nngp_predict_field   # north_star's name for it


def mcmc_nngp_predict_fixed_effects(mcmc_nngp_list, X_predicted, burn_in=.5, n_cores=1, match_field_thinning=True, add_intercept=False):
    first = next(iter(mcmc_nngp_list["records"].values()))
    stored = first["saved_field"] if match_field_thinning else np.arange(1, int(first["iterations"][-1, 0]) + 1)   # :70-71
    stored = stored[stored > burn_in * stored.max()].astype(int)
    Xp, names = _model_matrix(X_predicted)
    Xp = Xp - 0.0
    if add_intercept:
        Xp = np.column_stack([np.ones(Xp.shape[0]), Xp]); names = ["beta_0"] + names
    all_names = ["beta_0"] + list(first["params"].get("_beta_names", []))
    subset = [all_names.index(nm) for nm in names]                                                    # :85
    out = {}
    for name, chain in mcmc_nngp_list["records"].items():
        B = np.column_stack([chain["params"]["beta_0"], chain["params"]["beta"]])[stored - 1]
        if B.shape[1] > 1:
            B[:, 0] = B[:, 0] - B[:, 1:] @ mcmc_nngp_list["X"]["X_mean"]                              # :93
        out[name] = B[:, subset] @ Xp.T                                                               # :96
    return {"X_predicted": X_predicted, "predicted_fixed_effects_samples": out,
            "predicted_fixed_effects_summary": get_summary(np.vstack(list(out.values())))}
