"""Host-side logic of the spatially sharded field: partition, ghost sets, exchange lists.  The N > 1 path is
exercised with a world_size-2 gloo job on the CPU (no GPU, no compute call)."""
import os
import socket

import numpy as np
import pytest

import nngp_b200 as nb
from nngp_b200.partition import NA_INT, shard_plan, spatial_blocks


def problem(n=6000, m=8, seed=0):
    rng = np.random.default_rng(seed)
    locs = rng.random((n, 2))
    nn = nb.find_ordered_nn(locs, m)
    col = nb.greedy_coloring(nn)
    lm = np.concatenate([np.arange(1, n + 1), rng.integers(1, n + 1, 200)]).astype(np.int32)
    return locs, nn, col, lm


@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_blocks_are_balanced_and_cover_everything(P):
    locs, *_ = problem()
    owner = spatial_blocks(locs, P)
    counts = np.bincount(owner, minlength=P)
    assert counts.sum() == locs.shape[0] and counts.max() - counts.min() <= P


@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_native_plan_equals_the_numpy_prototype(P):
    """nngp_host_shard_plan_build (O(n (m+1)), csrc/host_shard.cpp) against the straightforward numpy derivation"""
    import partition_reference as PR
    locs, nn, col, lm = problem(5000, 7, seed=11)
    owner = spatial_blocks(locs, P)
    for r in range(P):
        a = shard_plan(locs, nn, col, lm, owner, r, P)
        b = PR.shard_plan(locs, nn, col, lm, owner, r, P)
        for k, vb in b.items():
            if k == "n_rows_needed":
                continue
            if isinstance(vb, np.ndarray):
                assert a[k].shape == vb.shape and np.array_equal(a[k], vb), k
            else:
                assert a[k] == vb, k


@pytest.mark.parametrize("P", [2, 4])
def test_plans_are_closed_and_mutually_consistent(P):
    locs, nn, col, lm = problem()
    owner = spatial_blocks(locs, P)
    plans = [shard_plan(locs, nn, col, lm, owner, r, P) for r in range(P)]
    n, K = locs.shape[0], int(col.max())
    assert sum(p["n_owned"] for p in plans) == n                             # every site has exactly one owner
    assert sum(p["obs_index"].size for p in plans) == lm.size               # every observation is counted once
    for p in plans:
        NN = p["NNarray"]
        gl = p["local_sites"]
        assert np.all(np.diff(gl) > 0)                                       # local order = global order
        valid = NN != NA_INT
        assert np.all(NN[:, 0] == np.arange(1, gl.size + 1))
        assert np.all((NN[:, 1:][valid[:, 1:]] >= 1))
        rows, cols = np.nonzero(valid[:, 1:])
        assert np.all(NN[:, 1:][rows, cols] <= rows)                         # parents precede their row (lower triangular)
        # every row that contains an owned site is present with all its parents, unchanged
        own_g = gl[p["owned"] == 1]
        contains_owned = np.isin(np.where(nn != NA_INT, nn - 1, -1), own_g).any(axis=1)
        for i in np.nonzero(contains_owned)[0][::97]:
            li = np.searchsorted(gl, i)
            assert gl[li] == i
            want = [v for v in nn[i] if v != NA_INT]
            got = [gl[v - 1] + 1 for v in NN[li] if v != NA_INT]
            assert got == want
    # what r sends to h for colour c is exactly what h expects from r, in the same order
    for r in range(P):
        for h in range(P):
            if r == h:
                continue
            for c in range(K):
                a, b = plans[r]["send_ptr"][c * P + h], plans[r]["send_ptr"][c * P + h + 1]
                ra, rb = plans[h]["recv_ptr"][c * P + r], plans[h]["recv_ptr"][c * P + r + 1]
                sent = plans[r]["global_id"][plans[r]["send_site"][a:b] - 1]
                expected = plans[h]["global_id"][plans[h]["recv_site"][ra:rb] - 1]
                assert np.array_equal(sent, expected)
                assert np.all(col[sent] == c + 1)
    # every ghost is received from exactly one owner
    for p in plans:
        ghosts = np.nonzero(p["owned"] == 0)[0] + 1
        assert sorted(p["recv_site"].tolist()) == sorted(ghosts.tolist())


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    try:
        locs, nn, col, lm = problem(4000, 6, seed=3)
        owner = spatial_blocks(locs, world)
        plan = shard_plan(locs, nn, col, lm, owner, rank, world)
        K = plan["n_colors"]
        truth = np.cos(np.arange(locs.shape[0]) * 0.37)                      # "new field value" of every site, by global id
        local = np.where(plan["owned"] == 1, truth[plan["global_id"]], np.nan)
        for c in range(K):                                                   # the per-colour halo exchange, over gloo
            reqs, bufs = [], {}
            for h in range(world):
                if h == rank:
                    continue
                a, b = plan["send_ptr"][c * world + h], plan["send_ptr"][c * world + h + 1]
                ra, rb = plan["recv_ptr"][c * world + h], plan["recv_ptr"][c * world + h + 1]
                if b > a:
                    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(local[plan["send_site"][a:b] - 1])), h))
                if rb > ra:
                    bufs[h] = (torch.empty(rb - ra, dtype=torch.float64), ra, rb)
                    reqs.append(dist.irecv(bufs[h][0], h))
            for r in reqs:
                r.wait()
            for h, (t, ra, rb) in bufs.items():
                local[plan["recv_site"][ra:rb] - 1] = t.numpy()
        ok = bool(np.all(np.isfinite(local)) and np.array_equal(local, truth[plan["global_id"]]))
        t = torch.tensor([float(plan["n_owned"])], dtype=torch.float64)
        dist.all_reduce(t)                                                   # the scalar all-reduce of the log-lik partials
        q.put((rank, ok, int(t.item()) == locs.shape[0]))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)
