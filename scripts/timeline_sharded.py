"""2-rank sharded sweep: prints rank 0's per-stage timestamps of the fused exchange kernel (development aid).
torchrun --nproc-per-node 2 scripts/timeline_sharded.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import nngp_b200 as nb
from nngp_b200.context import debug_timeline

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, m = 1_000_000, 10
rng = np.random.default_rng(1)
locs = rng.random((n, 2)); nn = nb.find_ordered_nn(locs, m); col = nb.greedy_coloring(nn)
lm = np.arange(1, n + 1, dtype=np.int32)
ctx, plan = nb.create_sharded_distributed(locs, nn, col, lm, "exponential_isotropic", local, dist, transport="p2p")
ctx.factor_build([1.0, 0.05, 0.0]); ctx.factor_commit()
w = np.random.default_rng(7).standard_normal(n) * 0.5
ctx.field_set(w[plan["local_sites"]]); ctx.obs_set(w[plan["obs_index"]])
ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), 3, seed=1)
for g in (1, 0):
    ctx.set_option("use_graph", g)
    ctx.set_option("debug_timeline", 1)
    ms, nl = ctx.time_op("gibbs_sweep", reps=3)
    ctx.set_option("debug_timeline", 0)
    if rank == 0:
        t = debug_timeline()
        print(f"graph={g}: sweep {ms.mean()*1e3:.1f} us, {len(t)} stamps, launches {nl}")
        names = {0: "kernel start", 1: "after griddep wait", 2: "after push", 3: "flags seen", 4: "end"}
        d = {}
        last = t[0, 0]
        for ts, st in t:
            d.setdefault(int(st), []).append(ts - last); last = ts
        for st in sorted(d):
            a = np.array(d[st][1:] if st == 0 else d[st])
            print(f"   -> {names[st]:20s} mean {a.mean():9.0f} ns  median {np.median(a):9.0f}  max {a.max():9.0f}  n={a.size}")
dist.barrier()
ctx.close()
dist.destroy_process_group()
