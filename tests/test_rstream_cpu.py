"""CPU checks of the product's R-compatible random stream and of the reference's own ordering / neighbour search regenerated
on it (include/nngp_b200.h: nngp_rng_*, nngp_host_order_maxmin_gpgp, nngp_host_find_ordered_nn_gpgp), against the values the
reference's vignette prints (tests/golden/vignette_golden.json) and against the oracle (oracle/r_rng.c, oracle/gpgp_order.c).
None of this needs a GPU: these are host set-up utilities in the reference too (Scripts/mcmc_nngp_initialize.R:17-110)."""
import numpy as np
import pytest

import nngp_b200 as nb
from nngp_b200 import api
from oracle import oracle as O

NA = nb.NA_INT


def toy(rs):
    """Vignette.rmd:26-47 on the product's stream"""
    rs.set_seed(1)
    locs = np.column_stack([500.0 * rs.runif(2000), np.ones(2000)])
    locs[0, 1] = 1.01
    D = np.sqrt(((locs[:, None, :] - locs[None, :, :]) ** 2).sum(-1))
    field = np.sqrt(10.0) * (np.linalg.cholesky(np.exp(-D / 5.0)) @ rs.rnorm(2000))
    X = np.column_stack([locs[:, 0], rs.rnorm(2000)])
    beta = np.array([0.01, rs.rnorm(1)[0]])
    beta_0 = rs.rnorm(1)[0]
    noise = np.sqrt(5.0) * rs.rnorm(2000)
    return locs, field + noise + X @ beta + beta_0, X


def test_stream_equals_the_oracles_draw_for_draw():
    rs = nb.RStream(20211)
    O.set_seed(20211)
    assert np.array_equal(rs.runif(1000), O.runif(1000))
    assert np.array_equal(rs.rnorm(1000), O.rnorm(1000))
    assert np.array_equal(rs.sample_int(777), O.sample_perm(777))
    assert np.array_equal(rs.sample_int(181, 1), O.sample_int(181, 1))
    assert np.array_equal(rs.sample_int(100000, 10), O.sample_int(100000, 10))          # 17 bits: two uniforms per draw
    assert np.array_equal(rs.rbeta(50, 10, 10), np.array([O.rbeta(10, 10) for _ in range(50)]))
    assert np.array_equal(rs.rbeta(20, 2.5, 7.0), np.array([O.rbeta(2.5, 7.0) for _ in range(20)]))
    assert np.array_equal(rs.runif(3), O.runif(3))                                      # still in step
    with pytest.raises(nb.NNGPError):
        rs.rbeta(1, 0.5, 2.0)                                                           # Cheng's BC branch is not carried


def test_state_is_the_callers():
    """the stream state is 625 caller-owned ints (R's .Random.seed[2:626]): copying it forks the stream"""
    rs = nb.RStream(5)
    rs.runif(10)
    fork = nb.RStream()
    fork.state[:] = rs.state
    assert np.array_equal(rs.rnorm(7), fork.rnorm(7))
    assert rs.state[0] == fork.state[0] and np.array_equal(rs.state, fork.state)


def test_runif_rnorm_match_the_vignette(golden):
    rs = nb.RStream(1)
    locs, _, X = toy(rs)
    assert np.allclose(locs[:6], golden["observed_locs_head"], rtol=0, atol=5e-5)              # Vignette.md:136-142
    Xc = X - X.mean(axis=0)
    assert np.allclose(Xc[:6], golden["X_head"], rtol=0, atol=5e-6)                            # :180-186


def test_order_maxmin_gpgp_reproduces_the_printed_ordering(golden):
    rs = nb.RStream(1)
    locs, _, _ = toy(rs)
    rs.set_seed(1)                                                                             # initialize.R:17
    order = nb.order_maxmin_gpgp(locs, rs)                                                     # :29
    assert np.array_equal(order[:100], golden["hctam_scol_1_100"])                             # Vignette.md:406-419
    lm = np.empty(2000, dtype=int)
    lm[order - 1] = np.arange(1, 2001)
    assert np.array_equal(lm[:100], golden["locs_match_100"])                                  # :322-328: the whole permutation
    nn = nb.find_ordered_nn_gpgp(locs[order - 1], 5, rs)                                       # :93
    want = np.array([[NA if v is None else v for v in r] for r in golden["NNarray_head"]], dtype=np.int32)
    assert np.array_equal(nn[:6], want)                                                        # :221-227
    # and the stream is where R's is: the next draw is chain 1's sample(., 1)  (initialize.R:154)
    O.set_seed(1)
    assert np.array_equal(O.order_maxmin_gpgp(locs), order)
    assert np.array_equal(O.find_ordered_nn_gpgp(locs[order - 1], 5), nn)
    assert np.array_equal(rs.runif(4), O.runif(4))


@pytest.mark.parametrize("n,d,seed", [(2, 2, 1), (3, 1, 2), (50, 2, 3), (1234, 2, 4), (2500, 3, 5), (900, 1, 6), (6000, 2, 7)])
def test_order_maxmin_gpgp_equals_the_oracle(n, d, seed):
    """grid queries on demand (product) against the n x sqrt(n) neighbour table GpGp builds (oracle): same permutation, same
    stream position afterwards"""
    locs = np.random.default_rng(seed).random((n, d))
    rs = nb.RStream(seed)
    O.set_seed(seed)
    a = nb.order_maxmin_gpgp(locs, rs)
    assert sorted(a.tolist()) == list(range(1, n + 1))
    assert np.array_equal(a, O.order_maxmin_gpgp(locs))
    nn = nb.find_ordered_nn_gpgp(locs[a - 1], 7, rs)
    assert np.array_equal(nn, O.find_ordered_nn_gpgp(locs[a - 1], 7))
    assert np.array_equal(rs.runif(2), O.runif(2))


def test_order_maxmin_gpgp_lonlat_is_a_permutation_with_spread_out_head():
    """the lon/lat branch is not pinned by any reference output; check what must hold: a permutation whose first points are
    farther apart on the sphere than those of any of 25 random prefixes"""
    g = np.random.default_rng(0)
    locs = np.column_stack([g.uniform(-180, 180, 3000), np.degrees(np.arcsin(g.uniform(-1, 1, 3000)))])
    rs = nb.RStream(11)
    o = nb.order_maxmin_gpgp(locs, rs, lonlat=True)
    assert sorted(o.tolist()) == list(range(1, 3001))

    def min_chord(idx):
        lon, lat = np.radians(locs[idx, 0]), np.radians(locs[idx, 1])
        P = np.column_stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
        D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1)) + 10 * np.eye(len(idx))
        return D.min()
    rand = [min_chord(g.permutation(3000)[:30]) for _ in range(25)]
    assert min_chord(o[:30] - 1) > 1.5 * np.median(rand) and min_chord(o[:30] - 1) > max(rand)


def test_initialize_host_part_reproduces_the_vignette(golden):
    """api.mcmc_nngp_initialize(rng = "R") minus the device step: ordering, vecchia_approx, regressors and chain 1's starting
    values as the vignette prints them (Vignette.md:133-501); the normals of the initial field draw are the stream's next n"""
    rs = nb.RStream(1)
    locs, y, X = toy(rs)
    lst, pending = api._initialize_host(locs, y, None, X, 5, "maxmin", "exponential_isotropic", "Gaussian", 3, 1, None, "R")
    va = lst["vecchia_approx"]
    assert np.array_equal(va["hctam_scol_1"][:100], golden["hctam_scol_1_100"])
    assert np.array_equal(va["locs_match"][:100], golden["locs_match_100"])
    assert np.allclose(lst["locs"][:6], golden["locs_head"], rtol=0, atol=5e-6)
    assert np.array_equal(va["sparse_chol_column_idx"][:100], np.arange(1, 101)) and np.array_equal(va["sparse_chol_row_idx"][:100], np.arange(1, 101))
    A = va["MRF_adjacency_mat"][:30, :30].toarray().astype(int)
    assert np.array_equal(A, np.array(golden["MRF_adjacency_30"]))                             # Vignette.md:275-306
    assert np.allclose(lst["X"]["X"][:6], golden["X_head"], rtol=0, atol=5e-6)
    g, p = golden["init_chain_1"], lst["states"]["chain_1"]["params"]
    assert abs(p["beta_0"] - g["beta_0"]) < 6e-6 and np.allclose(p["beta"], g["beta"], rtol=0, atol=6e-10)
    assert abs(p["log_scale"] - g["log_scale"]) < 6e-7 and abs(p["shape"][0] - g["shape"][0]) < 6e-7
    assert abs(p["log_noise_variance"] - g["log_noise_variance"]) < 6e-7
    # the device step is checked on the GPU (tests/test_gpu_vignette.py); here: the pending draw, solved by the oracle, gives
    # the printed field, so covparms and normals handed to the device are the reference's
    cp, z = pending["chain_1"]
    Linv = O.vecchia_Linv(cp, "exponential_isotropic", lst["locs"], va["NNarray"])
    f = p["beta_0"] + np.sqrt(np.exp(p["log_scale"])) * O.sparse_chol_solve(Linv, va["NNarray"], z)
    assert np.allclose(f[:100], g["field_100"], rtol=0, atol=6e-9)                             # :507-523
    assert np.array_equal(va["coloring"], O.naive_greedy_coloring(*O.moral_graph(va["NNarray"])))


@pytest.mark.parametrize("covfun,d,reordering,with_obs_reg", [("matern_isotropic", 2, "maxmin", False), ("exponential_scaledim", 3, "maxmin", True),
                                                               ("matern_spacetime", 3, "random", True), ("exponential_isotropic", 2, "random", False)])
def test_initialize_host_part_equals_the_oracle_driver(covfun, d, reordering, with_obs_reg):
    """every covariance family's starting values (sample(., 1) per range parameter, rnorm(1) per smoothness, :154-161), both
    seeded reorderings, duplicated locations, regressors at both levels: the product's host part and the oracle's driver are
    separate implementations of initialize.R on separate implementations of R's stream"""
    from oracle import reference_driver as R
    g = np.random.default_rng(17)
    base = g.random((700, d))
    obs_locs = np.vstack([base, base[g.integers(0, 700, 60)]])                  # 60 repeated observations
    n_obs = obs_locs.shape[0]
    y = g.standard_normal(n_obs) + 2.0 * obs_locs[:, 0]
    X_locs = np.column_stack([obs_locs[:, 0], np.sin(7 * obs_locs[:, 1])])      # functions of the location
    X_obs = g.standard_normal((n_obs, 1)) if with_obs_reg else None
    lst, pending = api._initialize_host(obs_locs, y, X_obs, X_locs, 6, reordering, covfun, "Gaussian", 2, 5, None, "R")
    ref = R.initialize(obs_locs, y, X_obs=X_obs, X_locs=X_locs, m=6, reordering=reordering, stationary_covfun=covfun, n_chains=2, seed=5)
    va, vr = lst["vecchia_approx"], ref["vecchia_approx"]
    assert np.array_equal(lst["locs"], ref["locs"])
    for k in ("locs_match", "hctam_scol_1", "obs_per_loc", "NNarray", "coloring", "sparse_chol_column_idx", "sparse_chol_row_idx"):
        assert np.array_equal(va[k], vr[k]), k
    assert lst["space_time_model"]["covfun"]["shape_params"] == ref["space_time_model"]["shape_params"]
    assert [c + 1 for c in lst["X"]["locs"]] == ref["X"]["locs"]
    assert np.allclose(lst["X"]["X"], ref["X"]["X"], rtol=0, atol=1e-14)
    assert np.allclose(lst["X"]["chol_solve_1XT1X"], ref["X"]["chol_solve_1XT1X"], rtol=1e-10, atol=1e-14)
    for name in ("chain_1", "chain_2"):
        p, q = lst["states"][name]["params"], ref["states"][name]["params"]
        assert np.array_equal(p["shape"], q["shape"])
        assert abs(p["beta_0"] - q["beta_0"]) < 1e-10 and np.allclose(p["beta"], q["beta"], rtol=0, atol=1e-10)
        assert abs(p["log_scale"] - q["log_scale"]) < 1e-12 and abs(p["log_noise_variance"] - q["log_noise_variance"]) < 1e-12
        cp, z = pending[name]
        Linv = O.vecchia_Linv(cp, covfun, lst["locs"], va["NNarray"])
        f = p["beta_0"] + np.sqrt(np.exp(p["log_scale"])) * O.sparse_chol_solve(Linv, va["NNarray"], z)
        assert np.allclose(f, q["field"], rtol=0, atol=1e-9)
