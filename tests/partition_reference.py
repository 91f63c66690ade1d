"""numpy prototype of the shard plan (test infrastructure: the checker of nngp_host_shard_plan_build / nngp_host_spatial_blocks).

Spatial sharding of one latent field across the GPUs of a box (BASELINE.json config 4).

Rank g OWNS the sites of one spatial block.  To sweep them it needs, locally:
  * the factor rows of every row that contains an owned site (its own rows + "ghost rows": children that live elsewhere),
  * the field value of every site appearing in those rows ("ghost sites" = moral-graph neighbours across the cut), kept
    current by a per-colour halo exchange of boundary values.
Local numbering preserves the global order, so the local factor stays lower triangular and the local NNarray obeys the same
invariants as a global one.  Everything here is host-side set-up (numpy); it is exercised without a GPU by the gloo tests.
"""
from __future__ import annotations

import numpy as np

NA_INT = -2147483648


def spatial_blocks(locs: np.ndarray, n_parts: int) -> np.ndarray:
    """recursive coordinate bisection into n_parts blocks of (almost) equal counts; returns owner[n] in 0..n_parts-1"""
    locs = np.asarray(locs, dtype=np.float64)
    owner = np.zeros(locs.shape[0], dtype=np.int32)

    def split(idx, lo, parts):
        if parts == 1:
            owner[idx] = lo
            return
        left = parts // 2
        ext = locs[idx].max(0) - locs[idx].min(0)
        k = int(np.argmax(ext[:min(2, locs.shape[1])]))
        order = idx[np.argsort(locs[idx, k], kind="stable")]
        cut = (order.size * left) // parts
        split(order[:cut], lo, left)
        split(order[cut:], lo + left, parts - left)

    split(np.arange(locs.shape[0]), 0, n_parts)
    return owner


def _needed_masks(NNarray: np.ndarray, owner: np.ndarray, n_parts: int):
    """rows_needed[h, i]: row i contains a site owned by h;  site_local[h, s]: site s appears in a row needed by h"""
    n = NNarray.shape[0]
    valid = NNarray != NA_INT
    par = np.where(valid, NNarray - 1, 0)
    own_of_entry = owner[par]
    rows_needed = np.zeros((n_parts, n), dtype=bool)
    site_local = np.zeros((n_parts, n), dtype=bool)
    for h in range(n_parts):
        rows_needed[h] = ((own_of_entry == h) & valid).any(axis=1)
        sel = par[rows_needed[h]][valid[rows_needed[h]]]
        site_local[h, sel] = True
    return rows_needed, site_local


def _levels(NNarray):
    """depth of every row in the solve DAG: 0 without parents, else 1 + max over the parents"""
    n = NNarray.shape[0]
    lvl = np.zeros(n, dtype=np.int64)
    for i in range(n):
        par = NNarray[i, 1:]
        par = par[par != NA_INT]
        if par.size:
            lvl[i] = lvl[par - 1].max() + 1
    return lvl


def shard_plan(locs, NNarray, coloring, locs_match, owner, rank, n_parts):
    """Everything rank `rank` needs to build its sharded context (indices 1-based / NA like the C ABI expects)."""
    locs = np.asarray(locs, dtype=np.float64)
    NNarray = np.asarray(NNarray, dtype=np.int32)
    coloring = np.asarray(coloring, dtype=np.int32)
    n, M = NNarray.shape
    K = int(coloring.max())
    rows_needed, site_local = _needed_masks(NNarray, owner, n_parts)
    mine = site_local[rank]
    local_sites = np.nonzero(mine)[0]                       # global 0-based ids, ascending = global order preserved
    g2l = np.full(n, -1, dtype=np.int64)
    g2l[local_sites] = np.arange(local_sites.size)
    nl = local_sites.size
    # local NNarray: real rows where needed, a trivial self-only row for ghost sites whose own row is not needed
    NN_loc = np.full((nl, M), NA_INT, dtype=np.int32)
    NN_loc[:, 0] = np.arange(1, nl + 1)
    need = rows_needed[rank][local_sites]
    src = NNarray[local_sites[need]]
    valid = src != NA_INT
    mapped = np.where(valid, g2l[np.where(valid, src - 1, 0)] + 1, NA_INT).astype(np.int32)
    assert np.all(mapped[valid] >= 1), "a parent of a needed row is not local"
    NN_loc[need] = mapped
    owned = (owner[local_sites] == rank)
    # observations of owned sites only (every observation is counted by exactly one rank)
    lm = np.asarray(locs_match, dtype=np.int64) - 1
    obs_sel = np.nonzero(owner[lm] == rank)[0]
    lm_loc = (g2l[lm[obs_sel]] + 1).astype(np.int32)
    # global position of every site in the reference's rnorm() hand-out order (colour 1..K, ascending index)
    cstart = np.concatenate([[0], np.cumsum(np.bincount(coloring, minlength=K + 1)[1:])])
    order = np.argsort(coloring, kind="stable")
    zpos_g = np.empty(n, dtype=np.int64)
    zpos_g[order] = np.arange(n)
    # halo exchange lists per (colour, peer), ordered by global id on both sides
    send_site, send_ptr, recv_site, recv_ptr = [], [0], [], [0]
    for c in range(1, K + 1):
        for h in range(n_parts):
            if h == rank:
                s_idx = r_idx = np.zeros(0, dtype=np.int64)
            else:
                s_idx = np.nonzero((owner == rank) & site_local[h] & (coloring == c))[0]       # my sites that h ghosts
                r_idx = np.nonzero((owner == h) & mine & (coloring == c))[0]                   # h's sites that I ghost
            send_site.append(g2l[s_idx] + 1)
            recv_site.append(g2l[r_idx] + 1)
            send_ptr.append(send_ptr[-1] + s_idx.size)
            recv_ptr.append(recv_ptr[-1] + r_idx.size)
    return {
        "rank": rank, "world": n_parts, "n_global": n, "n_colors": K,
        "local_sites": local_sites, "locs": locs[local_sites], "NNarray": NN_loc, "coloring": coloring[local_sites],
        "owned": owned.astype(np.int32), "global_id": local_sites.astype(np.int32), "global_zpos": zpos_g[local_sites].astype(np.int32),
        "global_level": _levels(NNarray)[local_sites].astype(np.int32),
        "obs_index": obs_sel, "locs_match": lm_loc,
        "send_site": np.concatenate(send_site).astype(np.int32) if send_site else np.zeros(0, np.int32),
        "send_ptr": np.array(send_ptr, dtype=np.int32),
        "recv_site": np.concatenate(recv_site).astype(np.int32) if recv_site else np.zeros(0, np.int32),
        "recv_ptr": np.array(recv_ptr, dtype=np.int32),
        "n_owned": int(owned.sum()), "n_ghost": int(nl - owned.sum()), "n_rows_needed": int(need.sum()),
    }
