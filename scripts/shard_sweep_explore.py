"""Development aid (run under torchrun, >= 2 ranks): the sharded sweep of an n = sites-per-gpu x N field under the launch options
of the fused halo exchange.  torchrun --nproc-per-node 2 scripts/shard_sweep_explore.py [--sites-per-gpu 1000000] [--m 10]"""
import argparse
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
    ap.add_argument("--covfun", default="exponential_isotropic")
    ap.add_argument("--reordering", default="maxmin")
    ap.add_argument("--reps", type=int, default=100)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    bench.pin_to_gpu_numa(local)
    if rank == 0:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import nngp_b200 as nb
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.sites_per_gpu * world
    cp = bench.covparms(a.covfun, bench.RANGE)
    beta_0, ls, lnv = 0.0, float(np.log(bench.SIGMA2)), float(np.log(bench.TAU2))
    locs, nn, coloring, locs_match, _ = bench.shared_problem(dist, rank, n, a.m, 1, a.reordering, "explore")
    ctx, plan = nb.create_sharded_distributed(locs, nn, coloring, locs_match, a.covfun, local, dist, transport="p2p")
    assert ctx.factor_build(cp) == 0
    ctx.factor_commit()
    w = np.random.default_rng(7).standard_normal(n) * 0.5
    y = w + np.sqrt(bench.TAU2) * np.random.default_rng(8).standard_normal(n)
    ctx.field_set(w[plan["local_sites"]])
    ctx.obs_set(y[plan["obs_index"]])
    ctx.gibbs_sweep(beta_0, ls, lnv, n_sweeps=1, seed=1)

    def maxred(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sp = np.asarray(plan["send_ptr"])
    rp = np.asarray(plan["recv_ptr"])
    per_col_recv = np.diff(rp.reshape(-1)[:: world]) if rp.size % world == 1 else None
    if rank == 0:
        print(f"n={n} world={world} m={a.m} colours={ctx.n_colors}; rank 0: local {plan['local_sites'].size}, ghosts {plan['n_ghost']}, sends {int(sp[-1])}", flush=True)
    for first, ctas in ((0, 296), (2, 74), (2, 148), (2, 296), (1, 74), (1, 296)):
        if True:
            ctx.set_option("shard_ghost_first", first)
            ctx.set_option("shard_ghost_ctas", ctas)
            ctx.time_op("gibbs_sweep", reps=5)
            dist.barrier()
            torch.cuda.synchronize()
            ms = ctx.time_op("gibbs_sweep", reps=a.reps)[0]
            med, mn = maxred(float(np.median(ms))), maxred(float(ms.min()))
            if rank == 0:
                print(f"  ghost_first={first} ghost_ctas={ctas:4d}: sweep median {med * 1e3:8.1f} us  min {mn * 1e3:8.1f} us", flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
