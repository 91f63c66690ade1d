"""Line-by-line dense transcription of one chain of Scripts/mcmc_nngp_update_Gaussian.R:34-315 (test infrastructure).

The reference is interpreted R, so the closest thing to "running it" in this image is to transcribe the loop statement by
statement into numpy with DENSE matrices standing in for Matrix::sparseMatrix objects (sparse_chol is an n x n array, solve()
is numpy.linalg.solve, crossprod(a, b) is a.T @ b, ...) and R's own random stream (Mersenne-Twister + inversion, pinned
against the vignette in test_oracle_golden.py).  The only call that is not transcribed is GpGp::vecchia_Linv (third-party,
restated in oracle/ and pinned by the dense-GP identities).  It exists to check the C oracle's chain -- in particular the
regression block :226-250 -- against an independently written form; it is O(n^2)-O(n^3) per iteration and only usable for
n of a few hundred.
"""
from __future__ import annotations

import numpy as np

from oracle import oracle as O

NA = np.iinfo(np.int32).min


def dnorm_log(x, mean, sd):
    return -0.5 * np.log(2 * np.pi) - np.log(sd) - 0.5 * ((x - mean) / sd) ** 2


def sparse_matrix(Linv, NNarray):
    """Matrix::sparseMatrix(i = sparse_chol_row_idx, j = sparse_chol_column_idx, x = Linv[NNarray_non_NA], triangular = T)"""
    n, M = NNarray.shape
    S = np.zeros((n, n))
    for j in range(M):
        ok = NNarray[:, j] != NA
        S[np.nonzero(ok)[0], NNarray[ok, j] - 1] = Linv[ok, j]
    return S


def ll_compressed_sparse_chol(Linv, field, NNarray, log_scale):                                       # :8-12
    chol_field = sparse_matrix(Linv, NNarray) @ field                                                  # GpGp::Linv_mult
    return np.sum(np.log(Linv[NNarray[:, 0] - 1, 0])) - NNarray.shape[0] * 0.5 * log_scale - 0.5 * np.sum(chol_field ** 2) / np.exp(log_scale)


def update_gaussian_chain(locs, NNarray, coloring, locs_match, observed_field, covfun, params, field, n_iterations_update,
                          field_thinning, n_chromatic, iter_start, i, X=None, X_locs_cols=(), beta=None, solve_1XT1X=None,
                          chol_solve_1XT1X=None):
    n_locs, n_obs = locs.shape[0], observed_field.size
    lm = np.asarray(locs_match) - 1
    obs_per_loc = np.bincount(lm, minlength=n_locs).astype(float)
    hctam_scol_1 = np.array([np.nonzero(lm == s)[0][0] for s in range(n_locs)])                        # initialize.R:88-90
    Xl_idx = [c - 1 for c in X_locs_cols]
    st = dict(shape=np.array(params["shape"], dtype=float), log_scale=params["log_scale"], beta_0=params["beta_0"],
              log_noise_variance=params["log_noise_variance"], field=np.array(field, dtype=float),
              beta=None if X is None else np.array(beta, dtype=float))
    logvar = dict(suf=params.get("logvar_sufficient", -2.0), anc=params.get("logvar_ancillary", -2.0))
    O.set_seed(iter_start + i)                                                                         # :36
    ns = st["shape"].size
    rec = np.zeros((n_iterations_update, 3 + ns))
    rec_beta = None if X is None else np.zeros((n_iterations_update, X.shape[1]))
    rec_field = np.zeros((int(round(n_iterations_update * field_thinning)), n_locs))
    acc_suf, acc_anc = np.zeros(n_iterations_update), np.zeros(n_iterations_update)
    def vecchia(shape):                                                                                # :67-72
        # shape_params "log_*" -> exp, "qlogis_*" (the Matern smoothness, last) -> .5 + .5 plogis
        sh = np.exp(shape)
        if covfun.startswith("matern"):
            sh[-1] = .5 + .5 / (1.0 + np.exp(-shape[-1]))
        return O.vecchia_Linv(np.concatenate([[1.0], sh, [0.0]]), covfun, locs, NNarray)
    compressed = vecchia(st["shape"])
    sparse_chol = sparse_matrix(compressed, NNarray)                                                   # :73
    precision_diag = (sparse_chol ** 2).sum(axis=0)                                                    # :74
    O.rnorm(n_locs)                                                                                    # :75 current_p

    def interweaved(sc):                                                                               # :77-83
        one_Xl = np.column_stack([np.ones(n_locs), X[hctam_scol_1][:, Xl_idx]])
        prec = (sc @ one_Xl).T @ (sc @ one_Xl)
        covmat = np.linalg.solve(prec, np.eye(prec.shape[0]))
        return covmat, np.linalg.cholesky(covmat).T, sc @ one_Xl                                       # chol() = upper factor

    if len(Xl_idx) > 0:
        iw_covmat, iw_chol, sc_X_locs = interweaved(sparse_chol)
    mu = st["beta_0"] + X @ st["beta"] if X is not None else np.full(n_obs, st["beta_0"])              # :85-86
    residuals_sum_matrix = np.zeros((n_locs, n_obs))                                                   # :90
    residuals_sum_matrix[lm, np.arange(n_obs)] = 1.0
    var_y = np.var(observed_field, ddof=1)
    colours = list(dict.fromkeys(coloring.tolist()))                                                   # unique(), order of appearance

    for it in range(1, n_iterations_update + 1):
        # ---- ancillary :113-157
        innovation = O.rnorm(ns + 1, 0, np.exp(.5 * logvar["anc"]))
        new_log_scale = st["log_scale"] + innovation[0]
        new_shape = st["shape"] + innovation[1:]
        new_compressed = vecchia(new_shape)
        new_sparse_chol = sparse_matrix(new_compressed, NNarray)
        new_field = st["beta_0"] + np.exp(.5 * (new_log_scale - st["log_scale"])) * np.linalg.solve(new_sparse_chol, sparse_chol @ (st["field"] - st["beta_0"]))
        sd = np.exp(0.5 * st["log_noise_variance"])
        field_response_ratio = np.sum(dnorm_log(observed_field, new_field[lm] + mu - st["beta_0"], sd) -
                                      dnorm_log(observed_field, st["field"][lm] + mu - st["beta_0"], sd))
        if field_response_ratio + 0 > np.log(O.runif(1)[0]):
            st["shape"], st["log_scale"], st["field"] = new_shape, new_log_scale, new_field
            compressed, sparse_chol = new_compressed, new_sparse_chol
            precision_diag = (sparse_chol ** 2).sum(axis=0)
            acc_anc[it - 1] = 1
            if len(Xl_idx) > 0:
                iw_covmat, iw_chol, sc_X_locs = interweaved(sparse_chol)
        if 0 <= iter_start <= 2000 and it / 25 == it // 25:
            a = acc_anc[it - 25:it].mean()
            if a < .05:
                logvar["anc"] = logvar["anc"] - O.rnorm(1, .4, .05)[0]
            if a > .15:
                logvar["anc"] = logvar["anc"] + O.rnorm(1, .4, .05)[0]
        # ---- sufficient :165-213
        innovation = O.rnorm(ns + 1, 0, np.exp(.5 * logvar["suf"]))
        new_log_scale = st["log_scale"] + innovation[0]
        if np.exp(new_log_scale) < var_y:
            new_shape = st["shape"] + innovation[1:]
            new_compressed = vecchia(new_shape)
            new_sparse_chol = sparse_matrix(new_compressed, NNarray)
            GP_ratio = (ll_compressed_sparse_chol(new_compressed, st["field"] - st["beta_0"], NNarray, new_log_scale) -
                        ll_compressed_sparse_chol(compressed, st["field"] - st["beta_0"], NNarray, st["log_scale"]))
            if GP_ratio > np.log(O.runif(1)[0]):
                st["shape"], st["log_scale"] = new_shape, new_log_scale
                compressed, sparse_chol = new_compressed, new_sparse_chol
                precision_diag = (sparse_chol ** 2).sum(axis=0)
                acc_suf[it - 1] = 1
                if len(Xl_idx) > 0:
                    iw_covmat, iw_chol, sc_X_locs = interweaved(sparse_chol)
        if 0 <= iter_start <= 2000 and it / 25 == it // 25:
            a = acc_suf[it - 25:it].mean()
            if a < .05:
                logvar["suf"] = logvar["suf"] - O.rnorm(1, .2, .05)[0]
            if a > .15:
                logvar["suf"] = logvar["suf"] + O.rnorm(1, .2, .05)[0]
        # ---- field mean :219-250
        if len(Xl_idx) == 0 or X is None:
            v = sparse_chol @ np.ones(n_locs)
            beta_covmat = (1.0 / (v @ v)) * np.exp(st["log_scale"])
            beta_mean = np.exp(-st["log_scale"]) * ((sparse_chol @ st["field"]) @ v) * beta_covmat
            st["beta_0"] = beta_mean + np.sqrt(beta_covmat) * O.rnorm(1)[0]
        if X is not None:
            one_X = np.column_stack([np.ones(n_obs), X])
            beta_mean = (observed_field - st["field"][lm] + st["beta_0"]) @ one_X @ solve_1XT1X        # :229
            innovation = beta_mean + np.exp(.5 * st["log_noise_variance"]) * (chol_solve_1XT1X.T @ O.rnorm(X.shape[1] + 1))
            st["field"] = st["field"] - st["beta_0"] + innovation[0]
            st["beta_0"] = innovation[0]
            st["beta"] = innovation[1:].copy()
            if len(Xl_idx) > 0:                                                                        # :237-246
                other_field = st["field"] + X[hctam_scol_1][:, Xl_idx] @ st["beta"][Xl_idx]
                beta_mean = iw_covmat @ ((sparse_chol @ other_field) @ sc_X_locs)
                innovation = beta_mean + np.exp(.5 * st["log_scale"]) * (iw_chol.T @ O.rnorm(len(Xl_idx) + 1))
                st["beta_0"] = innovation[0]
                st["beta"][Xl_idx] = innovation[1:]
                st["field"] = other_field - X[hctam_scol_1][:, Xl_idx] @ st["beta"][Xl_idx]
        mu = st["beta_0"] + X @ st["beta"] if X is not None else np.full(n_obs, st["beta_0"])
        # ---- chromatic sweeps :257-275
        for _ in range(n_chromatic):
            residuals_sum = residuals_sum_matrix @ (observed_field - mu)
            for color_idx in colours:
                sel = np.nonzero(coloring == color_idx)[0]
                posterior_precision = np.exp(-st["log_scale"]) * precision_diag[sel] + np.exp(-st["log_noise_variance"]) * obs_per_loc[sel]
                cond_mean = st["beta_0"] - (1 / posterior_precision) * (
                    (sparse_chol[:, sel].T @ (sparse_chol @ ((st["field"] - st["beta_0"]) * (coloring != color_idx)))) * np.exp(-st["log_scale"])
                    - np.exp(-st["log_noise_variance"]) * residuals_sum[sel])
                st["field"][sel] = cond_mean + O.rnorm(sel.size) / np.sqrt(posterior_precision)
        # ---- noise variance :281-293
        ssr = np.sum((observed_field - st["field"][lm] - mu + st["beta_0"]) ** 2)
        for _ in range(10):
            innovation = O.rnorm(1, 0, .01)[0]
            if np.exp(st["log_noise_variance"] + innovation) < var_y:
                if (-.5 * n_obs * innovation - .5 * ssr * (np.exp(-st["log_noise_variance"] - innovation) - np.exp(-st["log_noise_variance"]))
                        > np.log(O.runif(1)[0])):
                    st["log_noise_variance"] = st["log_noise_variance"] + innovation
        # ---- records :305-311
        rec[it - 1] = np.concatenate([[st["beta_0"], st["log_scale"], st["log_noise_variance"]], st["shape"]])
        if X is not None:
            rec_beta[it - 1] = st["beta"]
        if round(it * field_thinning) == it * field_thinning:
            rec_field[int(it * field_thinning) - 1] = st["field"]
    st["logvar_sufficient"], st["logvar_ancillary"] = logvar["suf"], logvar["anc"]
    return st, rec, rec_beta, rec_field, np.column_stack([acc_anc, acc_suf]).astype(np.int32)
