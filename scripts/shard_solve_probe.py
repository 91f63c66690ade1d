"""Development aid: the sharded triangular solve with all shards of a field on ONE GPU (nngp_shard_connect_local), for a few
problem sizes / shard counts / poll back-offs.  python scripts/shard_solve_probe.py"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import nngp_b200 as nb  # noqa: E402
from problems import make_problem  # noqa: E402

CP = [1.0, 0.05, 0.0]


def run(n, parts, seed, sleep_ns=0, window=0):
    P = make_problem(n, 10, seed=seed)
    b = np.random.default_rng(1).standard_normal(n)
    with nb.NNGPContext(P["locs"], P["NNarray"], P["coloring"], P["locs_match"]) as ctx:
        ctx.factor_build(CP)
        x_ref = ctx.sptrsv(b)
    owner = nb.spatial_blocks(P["locs"], parts)
    ctxs = []
    for r in range(parts):
        plan = nb.shard_plan(P["locs"], P["NNarray"], P["coloring"], P["locs_match"], owner, r, parts)
        c = nb.ShardedContext(plan, device=0, comm_id=None)
        c.factor_build(CP)
        c.set_option("solve_sleep_ns", sleep_ns)
        c.set_option("solve_window_ctas", window)
        ctxs.append(c)
    nb.connect_local(ctxs)
    out, err = [None] * parts, [None] * parts

    def work(k):
        try:
            out[k] = ctxs[k].sptrsv(b[ctxs[k].plan["local_sites"]])
        except Exception as e:  # noqa: BLE001
            err[k] = e

    for rep in range(2):
        ts = [threading.Thread(target=work, args=(k,)) for k in range(parts)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        dt = time.perf_counter() - t0
        ok = all(e is None for e in err) and all(np.max(np.abs(out[k] - x_ref[ctxs[k].plan["local_sites"]])) < 1e-9 * np.max(np.abs(x_ref)) for k in range(parts))
        print(f"n={n} parts={parts} seed={seed} sleep={sleep_ns} window={window} rep={rep}: {'OK' if ok else 'FAIL'} in {dt * 1e3:.1f} ms"
              + ("" if ok else "  " + "; ".join(str(e)[-160:] for e in err if e is not None)), flush=True)
        if not ok:
            break
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    run(24000, 3, 25)
    run(24000, 2, 25)
    run(24000, 2, 25, sleep_ns=500)
    run(24000, 2, 25, window=16)
    run(24000, 2, 26)
    run(40000, 2, 8)
    run(6000, 2, 25)
    run(100000, 2, 25)
