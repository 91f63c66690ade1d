"""Pins the CPU oracle: (i) against the printed values of the reference's rendered vignette (tests/golden/
vignette_golden.json, extracted by tests/golden/make_vignette_golden.py), (ii) against dense-GP identities that do not
depend on GpGp."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.special as sps

from oracle import oracle as O

NA = O.NA_INT


def toy_locs():
    """Vignette.rmd:26-29: set.seed(1); locs = cbind(500*runif(2000), 1); locs[1,2] = 1.01"""
    O.set_seed(1)
    u = O.runif(2000)
    locs = np.column_stack([500.0 * u, np.ones(2000)])
    locs[0, 1] = 1.01
    return locs


def test_runif_matches_vignette(golden):
    locs = toy_locs()
    g = np.array(golden["observed_locs_head"])
    assert np.allclose(locs[:6], g, rtol=0, atol=5e-5)          # printed with 7 significant digits


def test_rnorm_matches_vignette(golden):
    """Vignette.rmd:31,40: rnorm(2000) for the field, then X = (locs[,1], rnorm(2000)); Vignette.md:180-186 prints the
    centred columns, which pins R's inversion normal generator (MT + AS241)."""
    locs = toy_locs()
    O.rnorm(2000)
    wn = O.rnorm(2000)
    X = np.column_stack([locs[:, 0], wn])
    X = X - X.mean(axis=0)
    g = np.array(golden["X_head"])
    assert np.allclose(X[:6, 0], g[:, 0], rtol=0, atol=5e-6)
    assert np.allclose(X[:6, 1], g[:, 1], rtol=0, atol=5e-8)


def test_qnorm_against_cephes():
    ps = np.concatenate([np.linspace(1e-12, 1 - 1e-12, 4001), 10.0 ** -np.arange(1, 300, 7)])
    ps = ps[np.abs(ps - 0.5) > 1e-9]
    q = np.array([O.qnorm(p) for p in ps])
    assert np.max(np.abs(q - sps.ndtri(ps)) / np.abs(sps.ndtri(ps))) < 5e-15


def test_sample_perm_is_permutation():
    O.set_seed(3)
    p = O.sample_perm(1000)
    assert sorted(p.tolist()) == list(range(1, 1001))


def test_ordering_prefix_nnarray_and_moral_graph(golden):
    """The max-min permutation itself needs GpGp's RNG use and cannot be regenerated; its first 100 entries are printed
    (Vignette.md:406-419) and are used as data.  Everything derived from them must match the printed structures."""
    locs = toy_locs()
    perm = np.array(golden["hctam_scol_1_100"]) - 1
    pl = locs[perm]
    assert np.allclose(pl[:6], np.array(golden["locs_head"]), rtol=0, atol=5e-6)       # Vignette.md:148-154
    nn = O.find_ordered_nn(pl, 5)
    want = np.array([[NA if v is None else v for v in r] for r in golden["NNarray_head"]], dtype=np.int32)
    assert np.array_equal(nn[:6], want)                                                # Vignette.md:221-227
    adj_p, adj_i = O.moral_graph(nn)
    A = np.zeros((30, 30), dtype=int)
    for s in range(30):
        for t in adj_i[adj_p[s]:adj_p[s + 1]]:
            if t < 30:
                A[t, s] = 1
    assert np.array_equal(A, np.array(golden["MRF_adjacency_30"]))                     # Vignette.md:275-306
    # locs_match (obs -> position in reordered locs) is the inverse of hctam_scol_1 where both are printed
    lm = np.array(golden["locs_match_100"])
    for k, obs in enumerate(golden["hctam_scol_1_100"]):
        if obs <= 100:
            assert lm[obs - 1] == k + 1


def rand_problem(n, m, d=2, seed=0):
    rng = np.random.default_rng(seed)
    locs = rng.random((n, d))
    nn = O.find_ordered_nn(locs, m)
    return rng, locs, nn


def dense_from_linv(Linv, nn):
    n = nn.shape[0]
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(nn.shape[1]):
            if nn[i, j] != NA:
                A[i, nn[i, j] - 1] = Linv[i, j]
    return A


@pytest.mark.parametrize("covfun,cp", [("exponential_isotropic", [1.0, 0.3, 0.0]),
                                       ("matern_isotropic", [1.0, 0.3, 0.75, 0.0]),
                                       ("exponential_scaledim", [1.0, 0.3, 0.5, 0.0]),
                                       ("exponential_spacetime", [1.0, 0.3, 0.6, 0.0])])
def test_dense_gp_identity(covfun, cp):
    """m = n-1 makes the Vecchia factor exact: Linv == chol(Sigma)^-1 and the log-lik equals the dense Gaussian one."""
    n = 120
    rng, locs, nn = rand_problem(n, n - 1, seed=1)
    Linv = O.vecchia_Linv(cp, covfun, locs, nn)
    A = dense_from_linv(Linv, nn)
    if covfun == "exponential_isotropic":
        D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1)) / cp[1]
        S = np.exp(-D)
    elif covfun == "matern_isotropic":
        D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1)) / cp[1]
        nu = cp[2]
        with np.errstate(invalid="ignore"):
            S = 2 ** (1 - nu) / sps.gamma(nu) * D ** nu * sps.kv(nu, D)
        S[np.diag_indices(n)] = 1.0
    elif covfun == "exponential_scaledim":
        sc = locs / np.array(cp[1:3])
        S = np.exp(-np.sqrt(((sc[:, None, :] - sc[None, :, :]) ** 2).sum(-1)))
    else:
        sc = locs / np.array([cp[1], cp[2]])
        S = np.exp(-np.sqrt(((sc[:, None, :] - sc[None, :, :]) ** 2).sum(-1)))
    Lc = np.linalg.cholesky(S)
    Ainv_true = sla.solve_triangular(Lc, np.eye(n), lower=True)
    assert np.max(np.abs(A - Ainv_true)) / np.max(np.abs(Ainv_true)) < 1e-9
    z = rng.standard_normal(n)
    log_scale = 0.37
    ll = O.ll_compressed_sparse_chol(Linv, z, nn, log_scale)
    sign, logdet = np.linalg.slogdet(S)
    quad = z @ np.linalg.solve(S, z)
    ll_dense = -0.5 * logdet - 0.5 * n * log_scale - 0.5 * quad / np.exp(log_scale)
    assert abs(ll - ll_dense) < 1e-8 * abs(ll_dense)
    pd = O.precision_diag(Linv, nn)
    assert np.allclose(pd, np.diag(np.linalg.inv(S)), rtol=1e-8)


def test_sphere_matches_chordal_isotropic():
    rng = np.random.default_rng(5)
    n, m = 200, 6
    lonlat = np.column_stack([rng.uniform(-120, -70, n), rng.uniform(25, 50, n)])
    nn = O.find_ordered_nn(lonlat, m)
    lon, lat = np.deg2rad(lonlat[:, 0]), np.deg2rad(lonlat[:, 1])
    xyz = np.column_stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
    a = O.vecchia_Linv([1.0, 0.05, 0.0], "exponential_sphere", lonlat, nn)
    b = O.vecchia_Linv([1.0, 0.05, 0.0], "exponential_isotropic", xyz, nn)
    assert np.allclose(a, b, rtol=1e-9, atol=1e-12)   # deg->rad rounding differs (x*pi/180 vs deg2rad)


def test_row_identity_general_m():
    """Linv[i,0]^-2 = C_ii - c^T C_NN^-1 c and -Linv[i,1:]/Linv[i,0] = C_NN^-1 c."""
    n, m = 400, 10
    rng, locs, nn = rand_problem(n, m, seed=2)
    rng_ = 0.2
    Linv = O.vecchia_Linv([1.0, rng_, 0.0], "exponential_isotropic", locs, nn)
    for i in [0, 1, 5, 10, 11, 57, 399]:
        par = [p - 1 for p in nn[i, 1:] if p != NA]
        if not par:
            assert abs(Linv[i, 0] - 1.0) < 1e-14
            continue
        P = locs[par]
        C = np.exp(-np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1)) / rng_)
        c = np.exp(-np.sqrt(((P - locs[i]) ** 2).sum(-1)) / rng_)
        b = np.linalg.solve(C, c)
        F = 1.0 - c @ b
        assert abs(Linv[i, 0] ** -2 - F) < 1e-10 * F
        assert np.allclose(-Linv[i, 1:1 + len(par)] / Linv[i, 0], b, rtol=1e-8, atol=1e-12)


def test_solve_inverts_mult():
    n, m = 500, 8
    rng, locs, nn = rand_problem(n, m, seed=3)
    Linv = O.vecchia_Linv([1.0, 0.1, 0.0], "exponential_isotropic", locs, nn)
    w = rng.standard_normal(n)
    back = O.sparse_chol_solve(Linv, nn, O.Linv_mult(Linv, w, nn))
    assert np.max(np.abs(back - w)) < 1e-11
    A = dense_from_linv(Linv, nn)
    u = rng.standard_normal(n)
    assert np.allclose(O.sparse_chol_tmult(Linv, nn, u), A.T @ u, rtol=1e-12, atol=1e-12)


def coloring_of(nn):
    adj_p, adj_i = O.moral_graph(nn)
    return O.naive_greedy_coloring(adj_p, adj_i), adj_p, adj_i


def test_coloring_proper_and_first_fit():
    n, m = 1500, 10
    _, locs, nn = rand_problem(n, m, seed=4)
    cols, adj_p, adj_i = coloring_of(nn)
    for s in range(n):
        nb = adj_i[adj_p[s]:adj_p[s + 1]]
        nb = nb[nb != s]
        assert not np.any(cols[nb] == cols[s])                       # proper
        lower = set(cols[nb[nb < s]].tolist())
        assert all(c in lower for c in range(1, cols[s]))            # first-fit minimal => unique => bit-exact
    assert cols.min() == 1 and len(np.unique(cols)) == cols.max()
    # unique(coloring) in first-appearance order is 1..K (the reference iterates colours in that order)
    _, first = np.unique(cols, return_index=True)
    assert np.all(np.diff(first) > 0)


def test_sweep_forms_agree_and_match_dense_conditional():
    n, m = 300, 6
    rng, locs, nn = rand_problem(n, m, seed=6)
    Linv = O.vecchia_Linv([1.0, 0.15, 0.0], "exponential_isotropic", locs, nn)
    cols, _, _ = coloring_of(nn)
    pd = O.precision_diag(Linv, nn)
    n_obs = n + 40
    lm = np.concatenate([np.arange(1, n + 1), rng.integers(1, n + 1, 40)]).astype(np.int32)
    opl = np.bincount(lm - 1, minlength=n).astype(float)
    y = rng.standard_normal(n_obs)
    beta_0, ls, lnv = 0.4, 0.3, -0.7
    mu = np.full(n_obs, beta_0)
    rs = O.residuals_sum(lm, n, y, mu)
    field = beta_0 + rng.standard_normal(n)
    z = rng.standard_normal(n)
    f_ref = O.chromatic_sweep(Linv, nn, cols, pd, opl, rs, beta_0, ls, lnv, z, field, form="reference")
    f_res = O.chromatic_sweep(Linv, nn, cols, pd, opl, rs, beta_0, ls, lnv, z, field, form="residual")
    assert np.max(np.abs(f_ref - f_res)) < 1e-11
    # with z = 0 a sweep is a coloured Gauss-Seidel pass on (Q/s2 + D/t2) w = D-weighted residuals: check colour 1
    A = dense_from_linv(Linv, nn)
    Q = A.T @ A
    f0 = O.chromatic_sweep(Linv, nn, cols, pd, opl, rs, beta_0, ls, lnv, np.zeros(n), field, form="reference")
    sel = np.where(cols == 1)[0]
    w = field - beta_0
    prec = np.exp(-ls) * np.diag(Q)[sel] + np.exp(-lnv) * opl[sel]
    off = Q[sel] @ w - np.diag(Q)[sel] * w[sel]
    want = beta_0 - (off * np.exp(-ls) - np.exp(-lnv) * rs[sel]) / prec
    assert np.allclose(f0[sel], want, rtol=1e-10, atol=1e-12)


def test_bessel_against_scipy():
    for nu in (0.3, 0.5, 0.75, 1.2, 1.49):
        for x in (1e-6, 1e-3, 0.1, 0.7, 2.0, 2.5, 9.0, 40.0):
            assert abs(O.bessel_k(nu, x) - sps.kv(nu, x)) < 1e-12 * sps.kv(nu, x)


def test_predict_sample_matches_conditional_formula():
    n, n_pred, m = 200, 50, 5
    rng = np.random.default_rng(8)
    locs = rng.random((n + n_pred, 2))
    nn = O.find_ordered_nn(locs, m)
    Linv = O.vecchia_Linv([1.0, 0.2, 0.0], "exponential_isotropic", locs, nn)
    field = 0.3 + rng.standard_normal(n)
    z = rng.standard_normal(n_pred)
    ls = 0.5
    out = O.predict_field_sample(Linv, nn, n, field, 0.3, ls, z)
    x = np.concatenate([(field - 0.3) / np.exp(0.5 * ls), np.zeros(n_pred)])
    for i in range(n, n + n_pred):
        acc = z[i - n]
        for j in range(1, m + 1):
            if nn[i, j] != NA:
                acc -= Linv[i, j] * x[nn[i, j] - 1]
        x[i] = acc / Linv[i, 0]
    assert np.allclose(out, np.exp(0.5 * ls) * x[n:], rtol=1e-11, atol=1e-12)


def test_chain_forms_agree_and_are_sane():
    n, m = 400, 5
    rng, locs, nn = rand_problem(n, m, seed=9)
    cols, _, _ = coloring_of(nn)
    Linv = O.vecchia_Linv([1.0, 0.1, 0.0], "exponential_isotropic", locs, nn)
    w = O.sparse_chol_solve(Linv, nn, rng.standard_normal(n))
    y = 1.0 + w + np.sqrt(0.1) * rng.standard_normal(n)
    lm = np.arange(1, n + 1, dtype=np.int32)
    opl = np.ones(n)
    p0 = dict(shape=[np.log(0.1)], beta_0=0.9, log_scale=0.1, log_noise_variance=np.log(0.2))
    a = O.update_gaussian_chain(locs, nn, cols, lm, opl, y, "exponential_isotropic", p0, 0.9 + w, 30, 0.5, 3, 0, 1, 0)
    b = O.update_gaussian_chain(locs, nn, cols, lm, opl, y, "exponential_isotropic", p0, 0.9 + w, 30, 0.5, 3, 0, 1, 1)
    assert np.allclose(a[2], b[2], rtol=1e-8, atol=1e-8)          # scalar records
    assert np.allclose(a[1], b[1], rtol=1e-7, atol=1e-7)          # final field
    assert a[3].shape == (15, n)
    assert np.all(np.isfinite(a[2]))


@pytest.mark.parametrize("p_locs,p_obs,extra", [(2, 1, 25), (0, 2, 0), (3, 0, 10)])
def test_chain_with_regressors_matches_dense_transcription(p_locs, p_obs, extra):
    """The C oracle's chain with the regression block (update_Gaussian.R:226-250: block update of (beta_0, beta), interweaved
    centred update for location-level regressors, interweaving matrices refreshed on every accept) against the line-by-line
    dense numpy transcription of the R loop driven by the same R random stream: same accept/reject decisions, same records."""
    import dense_transcription as T
    from problems import make_regression_problem
    n, m, n_iter = 90, 4, 30
    P = make_regression_problem(n, m, seed=17, n_extra_obs=extra, p_locs=p_locs, p_obs=p_obs)
    p0 = dict(shape=[np.log(0.12)], beta_0=0.5, log_scale=-0.2, log_noise_variance=np.log(0.15), logvar_sufficient=-1.0,
              logvar_ancillary=-1.0)
    beta0 = 0.1 * np.ones(P["X"].shape[1])
    field0 = 0.5 + P["w"]
    reg = dict(X=P["X"], xlocs=P["xlocs"], first_obs=P["first_obs"], solve_1XT1X=P["solve_1XT1X"],
               chol_solve_1XT1X=P["chol_solve_1XT1X"], beta=beta0)
    for form in (0, 1):
        po, fo, reco, freco, acco, breco = O.update_gaussian_chain(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], P["obs_per_loc"],
                                                                   P["y"], "exponential_isotropic", p0, field0, n_iter, 0.5, 2, 0, 2, form,
                                                                   regressors=reg)
        st, rec, rec_beta, rec_field, acc = T.update_gaussian_chain(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], P["y"],
                                                                    "exponential_isotropic", p0, field0, n_iter, 0.5, 2, 0, 2, X=P["X"],
                                                                    X_locs_cols=list(P["xlocs"]), beta=beta0, solve_1XT1X=P["solve_1XT1X"],
                                                                    chol_solve_1XT1X=P["chol_solve_1XT1X"])
        assert np.array_equal(acc, acco)
        assert acc.sum() > 0                                      # some proposals were accepted: the refresh paths ran
        assert np.max(np.abs(rec - reco)) < 1e-8
        assert np.max(np.abs(rec_beta - breco)) < 1e-8
        assert np.max(np.abs(st["field"] - fo)) < 1e-7
        assert np.max(np.abs(rec_field - freco)) < 1e-7
        assert abs(st["logvar_ancillary"] - po["logvar_ancillary"]) < 1e-12
        assert np.max(np.abs(st["beta"] - po["beta"])) < 1e-8


@pytest.mark.parametrize("covfun,shape", [("exponential_isotropic", [np.log(0.12)]), ("matern_isotropic", [np.log(0.12), 0.4])])
def test_chain_without_regressors_matches_dense_transcription(covfun, shape):
    """Same check for the no-regressor loop (the path nngp_chain_run implements), exponential and Matern (qlogis smoothness)."""
    import dense_transcription as T
    from problems import make_regression_problem
    n, m, n_iter = 90, 4, 50
    P = make_regression_problem(n, m, seed=3, n_extra_obs=15, p_locs=1, p_obs=0)
    p0 = dict(shape=shape, beta_0=0.5, log_scale=-0.2, log_noise_variance=np.log(0.15), logvar_sufficient=-1.0,
              logvar_ancillary=-1.0)
    field0 = 0.5 + P["w"]
    po, fo, reco, freco, acco = O.update_gaussian_chain(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], P["obs_per_loc"], P["y"],
                                                        covfun, p0, field0, n_iter, 1.0, 2, 0, 1, 0)
    st, rec, _, rec_field, acc = T.update_gaussian_chain(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], P["y"],
                                                         covfun, p0, field0, n_iter, 1.0, 2, 0, 1)
    assert np.array_equal(acc, acco)
    assert np.max(np.abs(rec - reco)) < 1e-8
    assert np.max(np.abs(rec_field - freco)) < 1e-7


def test_matern_kernel_against_50_digit_arithmetic():
    """SURVEY 8(f4) "Matern Bessel-K accuracy study": the oracle's Matern correlation 2^(1-nu)/Gamma(nu) s^nu K_nu(s) (GpGp's
    parametrisation: no sqrt(2 nu) factor, value 1 at s = 0) over the smoothness range the reference's transforms can reach --
    .4 + .7 plogis in initialize.R:199, .5 + .5 plogis in update_Gaussian.R:70, 1.5 plogis in predict.R:37, i.e. (0, 1.5) -- and 12
    decades of scaled distance, against mpmath at 50 digits.  It is read off a 2-point Vecchia factor row, for which
    Linv[1, 0]^-2 = 1 - rho^2 and -Linv[1, 1] / Linv[1, 0] = rho, so the test goes through the same code the chain uses."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    worst = 0.0
    for nu in (0.05, 0.3, 0.4, 0.5, 0.75, 1.0, 1.1, 1.45, 1.499):
        for s in (1e-9, 1e-6, 1e-3, 0.03, 0.4, 1.0, 3.0, 9.0, 25.0, 60.0):
            rho = 2 ** (1 - mp.mpf(nu)) / mp.gamma(nu) * mp.mpf(s) ** nu * mp.besselk(nu, s)
            if float(1 - rho ** 2) < 1e-10:
                continue      # rho = 1 to double precision: the 2 x 2 block is singular (the chain rejects such proposals)
            locs = np.array([[0.0, 0.0], [s, 0.0]])
            nn = np.array([[1, NA], [2, 1]], dtype=np.int32)
            Linv = O.vecchia_Linv([1.0, 1.0, nu, 0.0], "matern_isotropic", locs, nn)
            got = -Linv[1, 1] / Linv[1, 0]
            err = abs(mp.mpf(float(got)) - rho) / max(rho, mp.mpf(10) ** -300)
            if rho > 1e-290:
                worst = max(worst, float(err))
            cond_var = 1.0 / Linv[1, 0] ** 2
            want_var = float(1 - rho ** 2)
            # 1 - rho^2 from a double rho: absolute error of a few ulps of 1 (the cancellation GpGp's double arithmetic has too)
            assert abs(cond_var - want_var) <= 1e-14, (nu, s, cond_var, want_var)
    assert worst < 5e-14, worst          # observed: 6.0e-15 (nu = 0.05, s = 1)
