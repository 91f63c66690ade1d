"""Development aid (torchrun, >= 2 ranks): where a colour of the sharded sweep spends its time (nngp_shard_timeline).
torchrun --nproc-per-node 2 scripts/shard_timeline.py [--sites-per-gpu 1000000] [--m 10]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=10)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    bench.pin_to_gpu_numa(local)
    if rank == 0:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import nngp_b200 as nb
    from nngp_b200 import _lib as L
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.sites_per_gpu * world
    beta_0, ls, lnv = 0.0, float(np.log(bench.SIGMA2)), float(np.log(bench.TAU2))
    locs, nn, coloring, locs_match, _ = bench.shared_problem(dist, rank, n, a.m, 1, "maxmin", "tl")
    ctx, plan = nb.create_sharded_distributed(locs, nn, coloring, locs_match, "exponential_isotropic", local, dist, transport="p2p")
    assert ctx.factor_build(bench.covparms("exponential_isotropic", bench.RANGE)) == 0
    ctx.factor_commit()
    w = np.random.default_rng(7).standard_normal(n) * 0.5
    ctx.field_set(w[plan["local_sites"]])
    ctx.obs_set((w + np.sqrt(bench.TAU2) * np.random.default_rng(8).standard_normal(n))[plan["obs_index"]])
    ctx.gibbs_sweep(beta_0, ls, lnv, n_sweeps=3, seed=1)
    K = ctx.n_colors
    acc = np.zeros((K, 6))
    cnt = np.zeros((K, 6))
    reps = 20
    for _ in range(reps):
        out, st = np.zeros(K * 6), C.c_int(0)
        dist.barrier()
        L.load().nngp_shard_timeline(L.ci(ctx._id), L.cd(beta_0), L.cd(ls), L.cd(lnv), L.dptr(out), C.byref(st))
        L.check(st)
        o = out.reshape(K, 6)
        ok = o >= 0
        rel = o - o[:, :1]          # relative to the colour's first tile past its wait
        acc += np.where(ok, rel, 0.0)
        cnt += ok
        start = o[:, 0]
    mean = acc / np.maximum(cnt, 1) / 1e3
    if rank == 0:
        print(f"n={n} world={world} m={a.m} colours={K}; per colour, us after the first tile passed griddepcontrol.wait (mean of {reps} sweeps, rank 0)")
        print("colour  next_colour_starts  last_push  first_ghost_seen  last_ghost_seen  last_ghost_patched  last_tile_done")
        nxt = np.append(np.diff(start), np.nan) / 1e3
        for c in range(K):
            print(f"{c + 1:6d} {nxt[c]:19.2f} {mean[c, 1]:10.2f} {mean[c, 2]:17.2f} {mean[c, 3]:16.2f} {mean[c, 4]:19.2f} {mean[c, 5]:15.2f}")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
