"""A spatially sharded field must reproduce the unsharded one: same sweep given the same normals, same Philox draws
(keyed by global site id), same log-likelihood.  On one GPU the shards live in one process and the halo is routed through the
host (colour-stepping ABI); with >= 2 GPUs the NCCL path is run as well."""
import os
import socket

import numpy as np
import pytest

import nngp_b200 as nb
from problems import make_problem

pytestmark = pytest.mark.gpu
CP = [1.0, 0.05, 0.0]
B0, LS, LNV = 0.2, 0.1, -1.0


def reference(P, n_sweeps, z=None, seed=3):
    with nb.NNGPContext(P["locs"], P["NNarray"], P["coloring"], P["locs_match"]) as ctx:
        ctx.factor_build(CP)
        ctx.factor_commit()
        ctx.field_set(P["field"])
        ctx.obs_set(P["y"])
        ll = ctx.loglik(B0, LS)
        ctx.gibbs_sweep(B0, LS, LNV, n_sweeps=n_sweeps, z=z, seed=seed)
        return ctx.field_get(), ll, ctx.loglik(B0, LS), ctx.ssr(), ctx.precision_diag()


def nb_solve_ref(P, b):
    with nb.NNGPContext(P["locs"], P["NNarray"], P["coloring"], P["locs_match"]) as ctx:
        ctx.factor_build(CP)
        return ctx.sptrsv(b)


def shard_contexts(P, parts):
    owner = nb.spatial_blocks(P["locs"], parts)
    ctxs = []
    for r in range(parts):
        plan = nb.shard_plan(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], owner, r, parts)
        c = nb.ShardedContext(plan, device=0, comm_id=None)
        c.factor_build(CP)
        c.factor_commit()
        c.field_set(P["field"][plan["local_sites"]])
        c.obs_set(P["y"][plan["obs_index"]])
        ctxs.append(c)
    return ctxs


def gather_owned(ctxs, n):
    out = np.full(n, np.nan)
    for c in ctxs:
        f = c.field_get()
        own = c.plan["owned"] == 1
        out[c.plan["local_sites"][own]] = f[own]
    return out


@pytest.mark.parametrize("parts", [2, 4])
def test_host_routed_sharded_sweep_equals_unsharded(parts):
    P = make_problem(30000, 10, seed=5, n_extra_obs=500)
    n = P["n"]
    z = P["rng"].standard_normal(2 * n)
    f_ref, ll0_ref, ll1_ref, ssr_ref, pd_ref = reference(P, 2, z=z)
    ctxs = shard_contexts(P, parts)
    try:
        # log-likelihood: partial sums of the owned rows add up to the whole (each context reports ll with n_global)
        parts_ll = [c.loglik(B0, LS) for c in ctxs]
        const = -n * 0.5 * LS
        assert abs(sum(l - const for l in parts_ll) + const - ll0_ref) < 1e-10 * abs(ll0_ref)
        for s in range(2):
            nb.host_routed_sweep(ctxs, B0, LS, LNV, z=z[s * n:(s + 1) * n])
        f = gather_owned(ctxs, n)
        assert np.max(np.abs(f - f_ref)) < 1e-10 * np.max(np.abs(f_ref))
        # ghosts carry the owners' values after the sweep
        for c in ctxs:
            assert np.max(np.abs(c.field_get() - f_ref[c.plan["local_sites"]])) < 1e-10 * np.max(np.abs(f_ref))
        assert abs(sum(c.ssr() for c in ctxs) - ssr_ref) < 1e-10 * ssr_ref
        for c in ctxs:
            own = c.plan["owned"] == 1
            assert np.allclose(c.precision_diag()[own], pd_ref[c.plan["local_sites"][own]], rtol=1e-12)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("parts,n", [(2, 30000), (3, 20000), (4, 60000), (8, 40000)])
def test_fused_peer_to_peer_sweep_on_one_gpu_equals_unsharded(parts, n):
    """The production transport -- boundary tiles push into the peers' ghost slots from inside the sweep kernel, trailing ghost
    CTAs wait on the peers' flags and apply -- with every shard in this process on one GPU (nngp_shard_connect_local): the
    flag / epoch / parity protocol, the per-colour peer masks and the PDL chain are exactly those of the multi-GPU run."""
    P = make_problem(n, 10, seed=15, n_extra_obs=300)
    n_sweeps = 3
    z = P["rng"].standard_normal(n_sweeps * n)
    f_ref, ll0_ref, ll1_ref, ssr_ref, _ = reference(P, n_sweeps, z=z)
    ctxs = shard_contexts(P, parts)
    try:
        nb.connect_local(ctxs)
        ll = nb.group_loglik(ctxs, B0, LS)
        assert np.all(np.abs(ll - ll0_ref) < 1e-10 * abs(ll0_ref)) and np.all(ll == ll[0])   # rank-ordered sum: identical on every member
        nb.group_sweep(ctxs, B0, LS, LNV, n_sweeps=n_sweeps, z=z)
        f = gather_owned(ctxs, n)
        assert np.max(np.abs(f - f_ref)) < 1e-10 * np.max(np.abs(f_ref))
        for c in ctxs:   # ghosts carry the owners' values
            assert np.max(np.abs(c.field_get() - f_ref[c.plan["local_sites"]])) < 1e-10 * np.max(np.abs(f_ref))
        ll = nb.group_loglik(ctxs, B0, LS)
        assert np.all(np.abs(ll - ll1_ref) < 1e-10 * abs(ll1_ref))
        # Philox sweeps, several calls (epochs keep counting across calls, parities alternate)
        with nb.NNGPContext(P["locs"], P["NNarray"], P["coloring"], P["locs_match"]) as ctx:
            ctx.factor_build(CP)
            ctx.factor_commit()
            ctx.field_set(P["field"])
            ctx.obs_set(P["y"])
            ctx.gibbs_sweep(B0, LS, LNV, n_sweeps=n_sweeps, z=z)   # same per-context sweep counter (part of the Philox key) as the shards
            for s in range(3):
                ctx.gibbs_sweep(B0, LS, LNV, n_sweeps=1 + s, seed=21)
            f_ref2 = ctx.field_get()
        for s in range(3):
            nb.group_sweep(ctxs, B0, LS, LNV, n_sweeps=1 + s, seed=21)
        f2 = gather_owned(ctxs, n)
        assert np.max(np.abs(f2 - f_ref2)) < 1e-10 * np.max(np.abs(f_ref2))
    finally:
        for c in ctxs:
            c.close()


def run_concurrently(fns):
    """the blocking entry points of the members of a locally connected field must be in flight together (they wait for each
    other inside their kernels): one host thread per member (ctypes releases the GIL during the call)"""
    import threading
    out, err = [None] * len(fns), [None] * len(fns)

    def work(k):
        try:
            out[k] = fns[k]()
        except Exception as e:   # noqa: BLE001
            err[k] = e

    ts = [threading.Thread(target=work, args=(k,)) for k in range(len(fns))]
    for th in ts:
        th.start()
    for th in ts:
        th.join()
    if any(e is not None for e in err):
        raise RuntimeError("; ".join(f"member {k}: {e!r}" for k, e in enumerate(err) if e is not None))
    return out


@pytest.mark.parametrize("parts", [2, 3, 4])
def test_sharded_triangular_solve_ancillary_step_and_whole_chain(parts):
    """Scripts/mcmc_nngp_update_Gaussian.R:127 (ancillary proposal = SpMV + triangular solve with the proposal factor) and the whole
    iteration loop (:101-314) on a sharded field: every rank solves its owned rows with the synchronisation-free kernel, boundary
    values go straight into the peers' solution vectors; the chain makes the same accept / reject decisions as the unsharded one."""
    n = 24000
    P = make_problem(n, 10, seed=25)
    b = P["rng"].standard_normal(n)
    cp2 = [1.0, 0.06, 0.0]
    var_y = float(np.var(P["y"], ddof=1))
    p0 = dict(shape=[np.log(0.05)], beta_0=B0, log_scale=LS, log_noise_variance=LNV)
    n_iter = 8
    with nb.NNGPContext(P["locs"], P["NNarray"], P["coloring"], P["locs_match"]) as ctx:
        ctx.factor_build(CP)
        ctx.factor_commit()
        ctx.field_set(P["field"])
        ctx.obs_set(P["y"])
        x_ref = ctx.sptrsv(b)
        ctx.factor_build(cp2, slot=nb.SLOT_PROPOSAL)
        ratio_ref = ctx.ancillary_propose(B0, 0.05, LNV)
        ctx.field_init(B0, LS, b)
        f_init_ref = ctx.field_get()
        ctx.field_set(P["field"])
        chain_ref = {}
        for mode in (nb.RNG_SUPPLIED, nb.RNG_PHILOX):
            ctx.field_set(P["field"])
            chain_ref[mode] = ctx.chain_run(p0, n_iter, var_y, thin=0.5, n_chromatic=2, iter_start=0, chain_index=2, rng_mode=mode) + (ctx.field_get(),)
    ctxs = shard_contexts(P, parts)
    try:
        nb.connect_local(ctxs)
        xs = run_concurrently([lambda c=c: c.sptrsv(b[c.plan["local_sites"]]) for c in ctxs])
        for c, x in zip(ctxs, xs):   # owned AND ghost entries hold the solution
            assert np.max(np.abs(x - x_ref[c.plan["local_sites"]])) < 1e-9 * np.max(np.abs(x_ref))
        for c in ctxs:
            c.factor_build(cp2, slot=nb.SLOT_PROPOSAL)
        ratios = run_concurrently([lambda c=c: c.ancillary_propose(B0, 0.05, LNV) for c in ctxs])
        assert all(r == ratios[0] for r in ratios) and abs(ratios[0] - ratio_ref) < 1e-9 * max(1.0, abs(ratio_ref))
        run_concurrently([lambda c=c: c.field_init(B0, LS, b[c.plan["local_sites"]]) for c in ctxs])
        for c in ctxs:
            assert np.max(np.abs(c.field_get() - f_init_ref[c.plan["local_sites"]])) < 1e-9 * np.max(np.abs(f_init_ref))
        for mode in (nb.RNG_SUPPLIED, nb.RNG_PHILOX):
            for c in ctxs:
                c.field_set(P["field"][c.plan["local_sites"]])
            po, rec, frecs, acc = nb.group_chain_run(ctxs, p0, n_iter, var_y, thin=0.5, n_chromatic=2, iter_start=0, chain_index=2, rng_mode=mode)
            po_r, rec_r, frec_r, acc_r, f_r = chain_ref[mode]
            assert np.array_equal(acc, acc_r)                                  # the same accept / reject sequence
            assert np.allclose(rec, rec_r, rtol=1e-8, atol=1e-10)
            for c, fr in zip(ctxs, frecs):
                assert np.allclose(fr, frec_r[:, c.plan["local_sites"]], rtol=1e-8, atol=1e-9)
                assert np.allclose(c.field_get(), f_r[c.plan["local_sites"]], rtol=1e-8, atol=1e-9)
    finally:
        for c in ctxs:
            c.close()


def test_sharded_matern_m20_factor_loglik_and_sweep_equal_unsharded():
    """config 4's shape (m = 20, matern_isotropic nu 0.75, range 0.02 scaled to this n) on a sharded field: the factor rows of the
    owned sites, the all-reduced log-likelihood and a fused peer-to-peer sweep reproduce the unsharded context (Scripts/
    mcmc_nngp_initialize.R:62-69 covariance families; update_Gaussian.R:70 smoothness in (0.5, 1))."""
    n, m, parts = 30000, 20, 4
    cp = [1.0, 0.06, 0.75, 0.0]
    P = make_problem(n, m, seed=35)
    z = P["rng"].standard_normal(2 * n)
    with nb.NNGPContext(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], "matern_isotropic") as ctx:
        assert ctx.factor_build(cp) == 0
        L_ref = ctx.factor_get()
        ctx.factor_commit()
        ctx.field_set(P["field"])
        ctx.obs_set(P["y"])
        ll_ref = ctx.loglik(B0, LS)
        ctx.gibbs_sweep(B0, LS, LNV, n_sweeps=2, z=z)
        f_ref = ctx.field_get()
    owner = nb.spatial_blocks(P["locs"], parts)
    ctxs = []
    try:
        for r in range(parts):
            plan = nb.shard_plan(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], owner, r, parts)
            c = nb.ShardedContext(plan, "matern_isotropic", device=0, comm_id=None)
            assert c.factor_build(cp) == 0
            own = plan["owned"] == 1
            Lg = c.factor_get()[own]
            Lr = L_ref[plan["local_sites"][own]]
            # an owned row's parents are all local, in the same order: the same row up to the (local) column layout
            assert np.max(np.linalg.norm(Lg - Lr, axis=1) / np.linalg.norm(Lr, axis=1)) < 1e-12
            c.factor_commit()
            c.field_set(P["field"][plan["local_sites"]])
            c.obs_set(P["y"][plan["obs_index"]])
            ctxs.append(c)
        nb.connect_local(ctxs)
        ll = nb.group_loglik(ctxs, B0, LS)
        assert np.all(np.abs(ll - ll_ref) < 1e-10 * abs(ll_ref))
        nb.group_sweep(ctxs, B0, LS, LNV, n_sweeps=2, z=z)
        f = gather_owned(ctxs, n)
        assert np.max(np.abs(f - f_ref)) < 1e-10 * np.max(np.abs(f_ref))
    finally:
        for c in ctxs:
            c.close()


def test_philox_draws_do_not_depend_on_the_sharding():
    P = make_problem(20000, 10, seed=6)
    f_ref, *_ = reference(P, 1, z=None, seed=9)
    ctxs = shard_contexts(P, 3)
    try:
        nb.host_routed_sweep(ctxs, B0, LS, LNV, z=None, seed=9)
        f = gather_owned(ctxs, P["n"])
        assert np.max(np.abs(f - f_ref)) < 1e-10 * np.max(np.abs(f_ref))
    finally:
        for c in ctxs:
            c.close()


def test_single_rank_sharded_context_is_the_plain_one():
    P = make_problem(5000, 10, seed=7)
    z = P["rng"].standard_normal(P["n"])
    f_ref, ll0, ll1, *_ = reference(P, 1, z=z)
    plan = nb.shard_plan(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], nb.spatial_blocks(P["locs"], 1), 0, 1)
    with nb.ShardedContext(plan) as c:
        c.factor_build(CP)
        c.factor_commit()
        c.field_set(P["field"])
        c.obs_set(P["y"])
        assert abs(c.loglik(B0, LS) - ll0) < 1e-10 * abs(ll0)
        c.gibbs_sweep(B0, LS, LNV, n_sweeps=1, z=z)              # world = 1: the ordinary entry point works
        assert np.max(np.abs(c.field_get() - f_ref)) < 1e-10
        assert np.max(np.abs(c.sptrsv(z) - nb_solve_ref(P, z))) < 1e-9    # a one-rank "sharded" field solves like the plain one


def _nccl_worker(rank, world, port, q, transport="nccl"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        P = make_problem(40000, 10, seed=8)
        n = P["n"]
        z = np.random.default_rng(1).standard_normal(2 * n)
        ctx, plan = nb.create_sharded_distributed(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], "exponential_isotropic", rank, dist, transport=transport)
        ctx.factor_build(CP)
        ctx.factor_commit()
        ctx.field_set(P["field"][plan["local_sites"]])
        ctx.obs_set(P["y"][plan["obs_index"]])
        ll = ctx.loglik(B0, LS)                                   # all-reduced inside the library
        ctx.gibbs_sweep(B0, LS, LNV, n_sweeps=2, z=z)             # halo exchange over NCCL after every colour
        ctx.gibbs_sweep(B0, LS, LNV, n_sweeps=1, seed=4)
        f = ctx.field_get()
        ssr = ctx.ssr()
        own = plan["owned"] == 1
        extra = None
        if transport == "p2p":   # the solve-based entry points and the whole chain across two GPUs
            bvec = np.random.default_rng(2).standard_normal(n)
            x = ctx.sptrsv(bvec[plan["local_sites"]])
            ctx.field_set(P["field"][plan["local_sites"]])
            po, rec, frec, acc = ctx.chain_run(dict(shape=[np.log(0.05)], beta_0=B0, log_scale=LS, log_noise_variance=LNV), 6, float(np.var(P["y"], ddof=1)),
                                               thin=0.0, n_chromatic=2, iter_start=0, chain_index=1, rng_mode=nb.RNG_SUPPLIED, keep_field=False)
            extra = (x[own], rec, acc, ctx.field_get()[own])
        q.put((rank, ll, plan["local_sites"][own], f[own], ssr, extra))
        ctx.close()
    except Exception as e:   # report instead of leaving the parent waiting on the queue
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_nccl_sharded_sweep_on_two_gpus(transport):
    if nb.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_nccl_worker, args=(r, 2, port, q, transport)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(len(r) == 6 for r in res), res
    P = make_problem(40000, 10, seed=8)
    n = P["n"]
    z = np.random.default_rng(1).standard_normal(2 * n)
    with nb.NNGPContext(P["locs"], P["NNarray"], P["coloring"], P["locs_match"]) as ctx:
        ctx.factor_build(CP)
        ctx.factor_commit()
        ctx.field_set(P["field"])
        ctx.obs_set(P["y"])
        ll_ref = ctx.loglik(B0, LS)
        ctx.gibbs_sweep(B0, LS, LNV, n_sweeps=2, z=z)
        ctx.gibbs_sweep(B0, LS, LNV, n_sweeps=1, seed=4)
        f_ref, ssr_ref = ctx.field_get(), ctx.ssr()
        if transport == "p2p":
            bvec = np.random.default_rng(2).standard_normal(n)
            x_ref = ctx.sptrsv(bvec)
            ctx.field_set(P["field"])
            _, rec_ref, _, acc_ref = ctx.chain_run(dict(shape=[np.log(0.05)], beta_0=B0, log_scale=LS, log_noise_variance=LNV), 6, float(np.var(P["y"], ddof=1)),
                                                   thin=0.0, n_chromatic=2, iter_start=0, chain_index=1, rng_mode=nb.RNG_SUPPLIED, keep_field=False)
            fc_ref = ctx.field_get()
    f = np.full(n, np.nan)
    for rank, ll, sites, vals, ssr, extra in res:
        assert abs(ll - ll_ref) < 1e-10 * abs(ll_ref)
        assert abs(ssr - ssr_ref) < 1e-10 * ssr_ref
        f[sites] = vals
        if transport == "p2p":
            x, rec, acc, fc = extra
            assert np.max(np.abs(x - x_ref[sites])) < 1e-9 * np.max(np.abs(x_ref))
            assert np.array_equal(acc, acc_ref) and np.allclose(rec, rec_ref, rtol=1e-8, atol=1e-10)
            assert np.allclose(fc, fc_ref[sites], rtol=1e-8, atol=1e-9)
    assert np.max(np.abs(f - f_ref)) < 1e-10 * np.max(np.abs(f_ref))
