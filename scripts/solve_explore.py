"""Development aid: spmv + triangular solve time for a few head sizes.  python scripts/solve_explore.py [--m 10]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(m):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import nngp_b200 as nb
    import bench
    n = 1_000_000
    _, locs, nn, col, lm, _ = bench.build_problem(n, m, seed=1, reordering="maxmin")
    ctx = nb.NNGPContext(locs, nn, col, lm)
    assert ctx.factor_build([1.0, 0.05, 0.0]) == 0
    ctx.factor_commit()
    ctx.field_init(0.0, 0.0, np.random.default_rng(1).standard_normal(n))
    for hd in (0, 1):
        ctx.set_option("solve_head", hd)
        ctx.time_op("sptrsv", reps=3)
        ms, nl = ctx.time_op("sptrsv", reps=30)
        print(f"  cap={os.environ.get('NNGP_HEAD_CAP_SLOTS', 'default')} width={os.environ.get('NNGP_HEAD_MAX_WIDTH', 'default')} head={hd}: spmv+sptrsv median {np.median(ms) * 1e3:7.1f} us min {ms.min() * 1e3:7.1f} us, launches {nl}, levels {ctx.n_levels}", flush=True)
    ms, _ = ctx.time_op("spmv", reps=30)
    print(f"  spmv alone {np.median(ms) * 1e3:7.1f} us", flush=True)
    ctx.close()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        one(int(sys.argv[2]))
    else:
        m = int(sys.argv[sys.argv.index("--m") + 1]) if "--m" in sys.argv else 10
        for cap, width in ((26000, 2048), (12000, 1024), (5000, 512), (2000, 128), (26000, 512)):
            env = dict(os.environ, NNGP_HEAD_CAP_SLOTS=str(cap), NNGP_HEAD_MAX_WIDTH=str(width))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(m)], env=env, capture_output=True, text=True)
            print(r.stdout[-700:], r.stderr[-300:] if r.returncode else "", flush=True)
