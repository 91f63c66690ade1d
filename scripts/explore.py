"""Times every kernel variant of the hot path on config 3 (development aid; not part of the bench contract).

    python scripts/explore.py [--n 1000000] [--m 10] [--reps 20]
"""
import argparse
import os
import sys

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
    ap.add_argument("--only-sweep", action="store_true", help="time the sweep variants only (Morton layout)")
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    locs = rng.random((a.n, 2))
    if a.order == "maxmin":
        locs = locs[nb.order_maxmin(locs) - 1]
    nn = nb.find_ordered_nn(locs, a.m)
    col = nb.greedy_coloring(nn)
    lm = np.arange(1, a.n + 1, dtype=np.int32)
    cp = [1.0, 0.05, 0.0] if a.covfun.startswith("exp") else [1.0, 0.05, 0.75, 0.0]
    for layout in ((nb.LAYOUT_MORTON,) if a.only_sweep else (nb.LAYOUT_MORTON, nb.LAYOUT_COLOR_MORTON)):
        ctx = nb.NNGPContext(locs, nn, col, lm, a.covfun, layout=layout)
        assert ctx.factor_build(cp) == 0
        ctx.factor_commit()
        ctx.field_init(0.0, 0.0, rng.standard_normal(a.n))
        w = ctx.field_get()
        ctx.obs_set(w + np.sqrt(0.1) * rng.standard_normal(a.n))
        ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), 1, seed=1)
        print(f"layout={layout} n={a.n} m={a.m} colors={ctx.n_colors} levels={ctx.n_levels} nnz={ctx.nnz} max_col={ctx.max_col}")
        for sv, slp in ():
            ctx.set_option("sweep_variant", sv)
            ctx.set_option("chain_sleep_ns", slp)
            ctx.set_option("use_graph", 1)
            ctx.time_op("gibbs_sweep", reps=3)
            ms, nl = ctx.time_op("gibbs_sweep", reps=a.reps)
            print(f"  sweep variant={sv} (flag-chained) sleep={slp}: mean {ms.mean()*1e3:8.1f} us  min {ms.min()*1e3:8.1f} us  launches {nl}")
        ctx.set_option("chain_sleep_ns", 0)
        for sv in (6, 22, 23):
            for g in ((1,) if sv in (0, 4, 5) else (1, 0)):
                ctx.set_option("sweep_variant", sv)
                ctx.set_option("use_graph", g)
                ctx.time_op("gibbs_sweep", reps=3)
                ms, nl = ctx.time_op("gibbs_sweep", reps=a.reps)
                print(f"  sweep variant={sv} graph={g}: mean {ms.mean()*1e3:8.1f} us  min {ms.min()*1e3:8.1f} us  launches {nl}")
        ctx.set_option("sweep_variant", 6)
        ctx.set_option("use_graph", 1)
        if a.only_sweep:
            # stress check: the flag-chained variants must reproduce the PDL chain bit for bit (same tiles, same Philox keys)
            f0 = ctx.field_get()
            zz = rng.standard_normal(20 * a.n)
            outs = {}
            for sv in (6, 9, 18):
                ctx.set_option("sweep_variant", sv)
                ctx.field_set(f0)
                ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), 20, z=zz)
                outs[sv] = ctx.field_get()
            print("  20 sweeps: max |v9 - v6| =", np.max(np.abs(outs[9] - outs[6])), " max |v18 - v6| =", np.max(np.abs(outs[18] - outs[6])))
            ctx.close()
            continue
        for sv, win, slp in ((0, 18, 0), (0, 37, 0), (0, 74, 0), (0, 111, 0), (0, 148, 0), (0, 222, 0), (0, 296, 0), (0, 74, 50), (0, 148, 50)):
            ctx.set_option("solve_variant", sv)
            ctx.set_option("solve_window_ctas", win)
            ctx.set_option("solve_sleep_ns", slp)
            ctx.time_op("sptrsv", reps=2)
            ms, nl = ctx.time_op("sptrsv", reps=max(3, a.reps // 4))
            print(f"  spmv+sptrsv variant={sv} window_ctas={win} sleep={slp}: mean {ms.mean()*1e3:8.1f} us  min {ms.min()*1e3:8.1f} us  launches {nl}")
        ctx.set_option("solve_window_ctas", 0)
        ctx.set_option("solve_ctas_per_sm", 1)
        ctx.set_option("solve_sleep_ns", 0)
        ctx.set_option("solve_variant", 0)
        ctx.set_option("commit_variant", 1)
        ms, nl = ctx.time_op("commit", reps=a.reps)
        print(f"  commit (thread per column): mean {ms.mean()*1e3:8.1f} us")
        ctx.set_option("commit_variant", 2)
        ms, nl = ctx.time_op("commit", reps=a.reps)
        print(f"  commit (tiled, segment sums): mean {ms.mean()*1e3:8.1f} us")
        ctx.set_option("commit_variant", 0)
        for fv in (1, 2, 0):
            ctx.set_option("factor_variant", fv)
            ctx.time_op("factor_build", reps=2)
            ms, nl = ctx.time_op("factor_build", reps=a.reps)
            print(f"  factor_build variant={fv}: mean {ms.mean()*1e3:8.1f} us  min {ms.min()*1e3:8.1f} us")
        ctx.set_option("loglik_variant", 0)
        ms, nl = ctx.time_op("loglik", reps=a.reps)
        ms2, _ = ctx.time_op("loglik", reps=a.reps, flush_l2=True)
        print(f"  loglik (TMA-staged ring)  : mean {ms.mean()*1e3:8.1f} us  min {ms.min()*1e3:8.1f} us  (L2 flushed: {ms2.mean()*1e3:8.1f} us)")
        ctx.set_option("loglik_variant", 1)
        for op in ("loglik", "spmv", "factor_build", "commit", "sweep_loglik"):
            ctx.time_op(op, reps=2)
            ms, nl = ctx.time_op(op, reps=a.reps)
            ms2, _ = ctx.time_op(op, reps=a.reps, flush_l2=True)
            print(f"  {op:14s}: mean {ms.mean()*1e3:8.1f} us  min {ms.min()*1e3:8.1f} us  (L2 flushed: {ms2.mean()*1e3:8.1f} us)  launches {nl}")
        ctx.close()


if __name__ == "__main__":
    main()
