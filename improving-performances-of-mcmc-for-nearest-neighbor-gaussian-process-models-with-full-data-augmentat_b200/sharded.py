"""Sharded contexts: one latent field split over the GPUs of a box by spatial blocks."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .context import NNGPContext
from .partition import shard_plan, spatial_blocks


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    st = C.c_int(0)
    L.load().nngp_comm_unique_id(buf, C.byref(st))
    L.check(st)
    return buf.raw


class ShardedContext(NNGPContext):
    """Context over the LOCAL site set of one rank (owned + ghost sites).  Vectors passed to field_set / field_get have
    plan["local_sites"].size entries; observations are those of the owned sites (plan["obs_index"])."""

    def __init__(self, plan: dict, covfun_name="exponential_isotropic", device=0, comm_id: bytes | None = None,
                 layout=L.LAYOUT_MORTON):
        self.plan = plan
        locs = np.asarray(plan["locs"], dtype=np.float64)
        self.n, self.d = locs.shape
        self.m = plan["NNarray"].shape[1] - 1
        self.covfun_name = covfun_name
        lm = L.i32(plan["locs_match"])
        self.n_obs = lm.size
        self.world, self.rank = plan["world"], plan["rank"]
        self.n_z = int(plan["n_global"])
        self._id = None
        cid, st = C.c_int(-1), C.c_int(0)
        idbuf = C.create_string_buffer(comm_id if comm_id else b"\0" * 128, 128)
        lib = L.load()
        lib.nngp_ctx_create_sharded(
            L.ci(self.n), L.ci(self.d), L.ci(self.m), L.dptr(L.f64(locs)), L.iptr(L.i32(plan["NNarray"])),
            L.iptr(L.i32(plan["coloring"])), L.ci(plan["n_colors"]), L.iptr(L.i32(plan["owned"])), L.iptr(L.i32(plan["global_id"])),
            L.iptr(L.i32(plan["global_zpos"])), L.iptr(L.i32(plan["global_level"])) if plan.get("global_level") is not None else None,
            L.cd(plan["n_global"]), L.ci(self.n_obs), L.iptr(lm), L.ci(L.COVFUN_IDS[covfun_name]),
            L.ci(device), L.ci(layout), L.ci(self.world), L.ci(self.rank), L.iptr(L.i32(plan["send_site"])), L.iptr(L.i32(plan["send_ptr"])),
            L.iptr(L.i32(plan["recv_site"])), L.iptr(L.i32(plan["recv_ptr"])), idbuf, C.byref(cid), C.byref(st))
        L.check(st)
        self._id = cid.value
        info = (C.c_int * 8)()
        lib.nngp_ctx_info(L.ci(self._id), info, C.byref(st))
        L.check(st)
        self.n_colors, self.n_levels, self.nnz, self.max_col = info[2], info[3], info[4], info[5]
        self.device, self.layout = info[6], info[7]

    # ---- peer-to-peer transport (CUDA IPC over NVLink)
    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._call("nngp_shard_p2p_export", buf)
        return buf.raw

    def p2p_connect(self, all_handles: list, all_recv_ptr: list):
        W, K = self.world, self.plan["n_colors"]
        base = np.zeros(K * W, dtype=np.int32)
        for c in range(K):
            for h in range(W):
                if h != self.rank:
                    base[c * W + h] = all_recv_ptr[h][c * W + self.rank]
        blob = C.create_string_buffer(b"".join(all_handles), 64 * W)
        self._call("nngp_shard_p2p_connect", blob, L.iptr(base))

    # ---- colour-stepping sweep (caller-moved halo)
    def sweep_begin(self, beta_0, log_scale, log_noise_variance, z=None, seed=0):
        if z is None:
            self._call("nngp_shard_sweep_begin", L.cd(beta_0), L.cd(log_scale), L.cd(log_noise_variance), L.ci(L.RNG_PHILOX), None, L.cd(seed))
        else:
            zz = L.f64(z)
            assert zz.size == self.plan["n_global"]
            self._call("nngp_shard_sweep_begin", L.cd(beta_0), L.cd(log_scale), L.cd(log_noise_variance), L.ci(L.RNG_SUPPLIED), L.dptr(zz), L.cd(seed))

    def sweep_colour(self, colour: int):
        self._call("nngp_shard_sweep_colour", L.ci(colour))

    def halo_get(self, colour: int) -> np.ndarray:
        W = self.world
        sp = self.plan["send_ptr"]
        out = np.zeros(max(int(sp[colour * W] - sp[(colour - 1) * W]), 1))
        self._call("nngp_shard_halo_get", L.ci(colour), L.dptr(out))
        return out[: int(sp[colour * W] - sp[(colour - 1) * W])]

    def halo_put(self, colour: int, values: np.ndarray):
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.size == 0:
            v = np.zeros(1)
        self._call("nngp_shard_halo_put", L.ci(colour), L.dptr(v))

    def sweep_end(self):
        self._call("nngp_shard_sweep_end")

    def send_segment(self, colour: int, peer: int):
        W, sp = self.world, self.plan["send_ptr"]
        base = sp[(colour - 1) * W]
        return int(sp[(colour - 1) * W + peer] - base), int(sp[(colour - 1) * W + peer + 1] - base)

    def recv_segment(self, colour: int, peer: int):
        W, rp = self.world, self.plan["recv_ptr"]
        base = rp[(colour - 1) * W]
        return int(rp[(colour - 1) * W + peer] - base), int(rp[(colour - 1) * W + peer + 1] - base)


def host_routed_sweep(contexts, beta_0, log_scale, log_noise_variance, z=None, seed=0):
    """One sweep of a sharded field whose shards live in ONE process (any devices): the halo is routed through the host.
    This is the reference driver for the colour-stepping ABI and the way the sharded arithmetic is tested on one GPU."""
    K = contexts[0].plan["n_colors"]
    for c in contexts:
        c.sweep_begin(beta_0, log_scale, log_noise_variance, z=z, seed=seed)
    for colour in range(1, K + 1):
        for c in contexts:
            c.sweep_colour(colour)
        outs = [c.halo_get(colour) for c in contexts]
        for h, dst in enumerate(contexts):
            r0, r1 = 0, dst.plan["recv_ptr"][colour * dst.world] - dst.plan["recv_ptr"][(colour - 1) * dst.world]
            buf = np.zeros(int(r1))
            for g, src in enumerate(contexts):
                if g == h:
                    continue
                a, b = src.send_segment(colour, h)
                ra, rb = dst.recv_segment(colour, g)
                assert b - a == rb - ra
                buf[ra:rb] = outs[g][a:b]
            dst.halo_put(colour, buf)
    for c in contexts:
        c.sweep_end()


def connect_local(contexts):
    """Connects the shards of one field that live in this process (one per GPU, or several on one GPU): afterwards
    group_sweep / group_loglik run the fused peer-to-peer halo exchange between them."""
    ids = (C.c_int * len(contexts))(*[c._id for c in contexts])
    st = C.c_int(0)
    L.load().nngp_shard_connect_local(ids, L.ci(len(contexts)), C.byref(st))
    L.check(st)


def group_sweep(contexts, beta_0, log_scale, log_noise_variance, n_sweeps=1, z=None, seed=0):
    ids = (C.c_int * len(contexts))(*[c._id for c in contexts])
    st = C.c_int(0)
    if z is None:
        L.load().nngp_shard_group_sweep(ids, L.ci(len(contexts)), L.ci(n_sweeps), L.cd(beta_0), L.cd(log_scale), L.cd(log_noise_variance),
                                        L.ci(L.RNG_PHILOX), None, L.cd(seed), C.byref(st))
    else:
        zz = L.f64(z)
        assert zz.size == n_sweeps * contexts[0].plan["n_global"]
        L.load().nngp_shard_group_sweep(ids, L.ci(len(contexts)), L.ci(n_sweeps), L.cd(beta_0), L.cd(log_scale), L.cd(log_noise_variance),
                                        L.ci(L.RNG_SUPPLIED), L.dptr(zz), L.cd(seed), C.byref(st))
    L.check(st)


def group_loglik(contexts, beta_0, log_scale, slot=L.SLOT_CURRENT) -> np.ndarray:
    ids = (C.c_int * len(contexts))(*[c._id for c in contexts])
    st = C.c_int(0)
    out = np.zeros(len(contexts))
    L.load().nngp_shard_group_loglik(ids, L.ci(len(contexts)), L.ci(slot), L.cd(beta_0), L.cd(log_scale), L.dptr(out), C.byref(st))
    L.check(st)
    return out


def group_chain_run(contexts, params: dict, n_iter, var_y, thin=1.0, n_chromatic=10, iter_start=0, chain_index=1, rng_mode=L.RNG_PHILOX,
                    keep_field=True):
    """nngp_shard_group_chain_run: one chain on a locally connected sharded field.  Returns (params, records, [field records per
    member], accepts) laid out like NNGPContext.chain_run."""
    W = len(contexts)
    shape = np.atleast_1d(np.asarray(params["shape"], dtype=np.float64))
    p = np.ascontiguousarray(np.concatenate([[params["beta_0"], params["log_scale"], params["log_noise_variance"],
                                              params.get("logvar_sufficient", -2.0), params.get("logvar_ancillary", -2.0)], shape]))
    n_iter = int(n_iter)
    n_frec = int(round(n_iter * thin))
    rec = np.zeros(n_iter * (3 + shape.size))
    sizes = [c.n for c in contexts]
    frec = np.zeros(max(n_frec * sum(sizes), 1)) if keep_field else None
    acc = np.zeros(2 * n_iter, dtype=np.int32)
    ids = (C.c_int * W)(*[c._id for c in contexts])
    st = C.c_int(0)
    L.load().nngp_shard_group_chain_run(ids, L.ci(W), L.ci(shape.size), L.dptr(p), L.ci(n_iter), L.cd(thin), L.ci(n_chromatic), L.ci(iter_start),
                                        L.ci(chain_index), L.ci(rng_mode), L.cd(var_y), L.dptr(rec), L.dptr(frec) if keep_field else None,
                                        L.iptr(acc), C.byref(st))
    L.check(st)
    out = dict(beta_0=p[0], log_scale=p[1], log_noise_variance=p[2], logvar_sufficient=p[3], logvar_ancillary=p[4], shape=p[5:].copy())
    frecs = None
    if keep_field:
        frecs, off = [], 0
        for nl in sizes:
            frecs.append(frec[off: off + n_frec * nl].reshape((n_frec, nl), order="F"))
            off += n_frec * nl
    return out, rec.reshape((n_iter, 3 + shape.size), order="F"), frecs, acc.reshape((n_iter, 2), order="F")


def create_sharded_distributed(locs, NNarray, coloring, locs_match, covfun_name, device, dist, transport="p2p"):
    """torch.distributed driver: every rank calls this with the same (replicated) global structure.  transport "p2p" maps the
    peers' receive areas through CUDA IPC (halo values are stored straight into the peers' memory over NVLink); "nccl"
    uses ncclSend/ncclRecv per colour (rank 0's NCCL id is broadcast).  Returns (ShardedContext, plan)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    owner = spatial_blocks(locs, world)
    plan = shard_plan(locs, NNarray, coloring, locs_match, owner, rank, world)
    del owner
    box = [comm_unique_id() if (rank == 0 and transport == "nccl") else None]
    dist.broadcast_object_list(box, src=0)
    ctx = ShardedContext(plan, covfun_name, device=device, comm_id=box[0])
    if transport == "p2p" and world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, (ctx.p2p_export(), plan["recv_ptr"]))
        ctx.p2p_connect([h[0] for h in handles], [h[1] for h in handles])
        dist.barrier()   # nobody pushes before everybody has mapped everybody
    return ctx, plan
