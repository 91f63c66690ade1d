"""Spatial sharding of one latent field across the GPUs of a box (BASELINE.json config 4).

Rank g OWNS the sites of one spatial block.  To sweep them it needs, locally:
  * the factor rows of every row that contains an owned site (its own rows + "ghost rows": children that live elsewhere),
  * the field value of every site appearing in those rows ("ghost sites" = moral-graph neighbours across the cut), kept
    current by a per-colour halo exchange of boundary values.
Local numbering preserves the global order, so the local factor stays lower triangular and the local NNarray obeys the same
invariants as a global one.  The work is done by the native host utilities (csrc/host_shard.cpp, O(n (m+1)) per rank);
this module is their ctypes binding.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

NA_INT = -2147483648


def spatial_blocks(locs: np.ndarray, n_parts: int) -> np.ndarray:
    """recursive coordinate bisection into n_parts blocks of (almost) equal counts; returns owner[n] in 0..n_parts-1"""
    locs = np.asarray(locs, dtype=np.float64)
    if locs.ndim == 1:
        locs = locs[:, None]
    n, d = locs.shape
    owner = np.zeros(n, dtype=np.int32)
    st = C.c_int(0)
    L.load().nngp_host_spatial_blocks(L.dptr(L.f64(locs)), L.ci(n), L.ci(d), L.ci(n_parts), L.iptr(owner), C.byref(st))
    L.check(st)
    return owner


def shard_plan(locs, NNarray, coloring, locs_match, owner, rank, n_parts):
    """Everything rank `rank` needs to build its sharded context (indices 1-based / NA like the C ABI expects)."""
    locs = np.asarray(locs, dtype=np.float64)
    if locs.ndim == 1:
        locs = locs[:, None]
    n, d = locs.shape
    nn = L.i32(NNarray)
    M = np.asarray(NNarray).shape[1]
    col = L.i32(coloring)
    lm = L.i32(locs_match)
    own = L.i32(owner)
    pid, st = C.c_int(-1), C.c_int(0)
    sizes = (C.c_int * 6)()
    lib = L.load()
    lib.nngp_host_shard_plan_build(L.dptr(L.f64(locs)), L.iptr(nn), L.iptr(col), L.ci(n), L.ci(d), L.ci(M - 1), L.ci(lm.size), L.iptr(lm),
                                   L.iptr(own), L.ci(rank), L.ci(n_parts), C.byref(pid), sizes, C.byref(st))
    L.check(st)
    nl, nol, ns, nr, K, n_owned = [int(v) for v in sizes]
    out_locs = np.zeros(nl * d)
    NN_loc = np.zeros(nl * M, dtype=np.int32)
    ints = {k: np.zeros(max(sz, 1), dtype=np.int32) for k, sz in
            dict(coloring=nl, owned=nl, global_id=nl, global_zpos=nl, global_level=nl, obs_index=nol, locs_match=nol, send_site=ns, recv_site=nr,
                 send_ptr=K * n_parts + 1, recv_ptr=K * n_parts + 1).items()}
    lib.nngp_host_shard_plan_get(C.byref(pid), L.dptr(out_locs), L.iptr(NN_loc), L.iptr(ints["coloring"]), L.iptr(ints["owned"]),
                                 L.iptr(ints["global_id"]), L.iptr(ints["global_zpos"]), L.iptr(ints["global_level"]), L.iptr(ints["obs_index"]), L.iptr(ints["locs_match"]),
                                 L.iptr(ints["send_site"]), L.iptr(ints["send_ptr"]), L.iptr(ints["recv_site"]), L.iptr(ints["recv_ptr"]),
                                 C.byref(st))
    L.check(st)
    gid = ints["global_id"][:nl]
    return {
        "rank": rank, "world": n_parts, "n_global": n, "n_colors": K,
        "local_sites": gid.astype(np.int64), "locs": out_locs.reshape((nl, d), order="F"), "NNarray": NN_loc.reshape((nl, M), order="F"),
        "coloring": ints["coloring"][:nl], "owned": ints["owned"][:nl], "global_id": gid, "global_zpos": ints["global_zpos"][:nl], "global_level": ints["global_level"][:nl],
        "obs_index": ints["obs_index"][:nol].astype(np.int64), "locs_match": ints["locs_match"][:nol],
        "send_site": ints["send_site"][:ns], "send_ptr": ints["send_ptr"], "recv_site": ints["recv_site"][:nr], "recv_ptr": ints["recv_ptr"],
        "n_owned": n_owned, "n_ghost": nl - n_owned, "n_rows_needed": None,
    }
