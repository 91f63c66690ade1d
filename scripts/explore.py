"""Times the kernel options of the hot path on config 3 (development aid; not part of the bench contract).

    python scripts/explore.py [--n 1000000] [--m 10] [--reps 20] [--order random|maxmin] [--covfun ...]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nngp_b200 as nb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=10)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--order", default="random", choices=["random", "maxmin"])
    ap.add_argument("--covfun", default="exponential_isotropic")
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    locs = rng.random((a.n, 2))
    t0 = time.perf_counter()
    if a.order == "maxmin":
        locs = locs[nb.order_maxmin(locs) - 1]
    t1 = time.perf_counter()
    nn = nb.find_ordered_nn(locs, a.m)
    t2 = time.perf_counter()
    col = nb.greedy_coloring(nn)
    t3 = time.perf_counter()
    lm = np.arange(1, a.n + 1, dtype=np.int32)
    cp = [1.0, 0.05, 0.0] if a.covfun.startswith("exp") else [1.0, 0.05, 0.75, 0.0]
    ctx = nb.NNGPContext(locs, nn, col, lm, a.covfun)
    t4 = time.perf_counter()
    print(f"n={a.n} m={a.m} order={a.order} {a.covfun}: ordering {t1 - t0:.2f} s, neighbours {t2 - t1:.2f} s, colouring {t3 - t2:.2f} s, "
          f"ctx_create {t4 - t3:.2f} s; colours {ctx.n_colors}, solve levels {ctx.n_levels}, longest column {ctx.max_col}", flush=True)
    assert ctx.factor_build(cp) == 0
    ctx.factor_commit()
    ctx.field_init(0.0, 0.0, rng.standard_normal(a.n))
    w = ctx.field_get()
    ctx.obs_set(w + np.sqrt(0.1) * rng.standard_normal(a.n))
    ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), 1, seed=1)

    def line(label, op):
        ctx.time_op(op, reps=3)
        ms, nl = ctx.time_op(op, reps=a.reps)
        print(f"  {label}: median {np.median(ms) * 1e3:8.1f} us  min {ms.min() * 1e3:8.1f} us  launches {nl}", flush=True)

    for sv in (0, 4, 1, 2, 3):
        for g in ((1, 0) if sv == 0 else (1,)):
            ctx.set_option("sweep_variant", sv)
            ctx.set_option("use_graph", g)
            line(f"sweep variant={sv} graph={g}", "gibbs_sweep")
    ctx.set_option("sweep_variant", 0)
    ctx.set_option("use_graph", 1)
    for sv, win in ((0, 0), (0, 296), (0, 592), (1, 0)):
        ctx.set_option("solve_variant", sv)
        ctx.set_option("solve_window_ctas", win)
        line(f"spmv+sptrsv variant={sv} window_ctas={win}", "sptrsv")
    ctx.set_option("solve_window_ctas", 0)
    ctx.set_option("solve_variant", 0)
    for lc in (0, 1):   # the factor build also writes the factor in the solve's row order (1, default) or not (0)
        ctx.set_option("solve_level_copy", lc)
        assert ctx.factor_build(cp) == 0
        ctx.factor_commit()
        line(f"spmv+sptrsv level_copy={lc}", "sptrsv")
        line(f"factor_build level_copy={lc}", "factor_build")
    for cv in (0, 1):
        ctx.set_option("commit_variant", cv)
        line(f"transpose+precision_diag variant={cv}", "commit")
    ctx.set_option("commit_variant", 0)
    line("factor_build", "factor_build")
    for lv in (1, 0):
        ctx.set_option("loglik_variant", lv)
        line(f"loglik variant={lv} ({'plain' if lv else 'TMA ring'})", "loglik")
    ctx.set_option("loglik_variant", 1)
    line("spmv", "spmv")
    line("sweep + loglik", "sweep_loglik")
    # several chains sharing the GPU: aggregate sweeps/s with the dependents triggered at the top (0) or after the wait (4)
    others = []
    for k in range(2):
        c2 = nb.NNGPContext(locs, nn, col, lm, a.covfun)
        c2.factor_build(cp)
        c2.factor_commit()
        c2.field_set(w)
        c2.obs_set(w)
        c2.gibbs_sweep(0.0, 0.0, np.log(0.1), 1, seed=k + 2)
        others.append(c2)
    for sv in (0, 4):
        for c in [ctx] + others:
            c.set_option("sweep_variant", sv)
        for nc in (1, 2, 3):
            ms = nb.time_op_group(([ctx] + others)[:nc], "gibbs_sweep", reps=100)
            print(f"  {nc} chain(s) on one GPU, sweep variant {sv}: {nc * 100 / ms * 1e3:9.0f} sweeps/s aggregate", flush=True)
    for c2 in others:
        c2.close()
    ctx.close()


if __name__ == "__main__":
    main()
