"""Per-colour timeline of the PDL-chained sweep (development aid): when the last CTA of colour c reaches
griddepcontrol.wait, when the first one is released, when the first / last CTA is past its scatter."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nngp_b200 as nb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = 10
rng = np.random.default_rng(1)
locs = rng.random((n, 2))
nn = nb.find_ordered_nn(locs, m)
col = nb.greedy_coloring(nn)
ctx = nb.NNGPContext(locs, nn, col, np.arange(1, n + 1, dtype=np.int32))
ctx.factor_build([1.0, 0.05, 0.0]); ctx.factor_commit()
ctx.field_init(0.0, 0.0, rng.standard_normal(n))
ctx.obs_set(ctx.field_get() + np.sqrt(0.1) * rng.standard_normal(n))
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx.set_option("sweep_variant", variant)
ms, _ = ctx.time_op("gibbs_sweep", reps=10)
print(f"sweep (no stamps) {ms.mean()*1e3:.1f} us")
ctx.set_option("debug_timeline", 1)
for rep in range(3):
    ctx.gibbs_sweep(0.0, 0.0, np.log(0.1), 1, seed=1)
t = ctx.debug_colour_times() / 1e3
sizes = np.bincount(col)[1:]
print("colour   sites   last_ready  first_rel  first_done  last_done | span(rel->last_done)  handoff(prev last_done -> first_rel)")
prev = None
for c in range(ctx.n_colors):
    lr, fr, fd, ld = t[c]
    ho = (fr - prev) if prev is not None else float("nan")
    print(f"{c:4d} {sizes[c]:8d}   {lr:9.2f} {fr:9.2f} {fd:9.2f} {ld:9.2f} | {ld - fr:8.2f} {ho:8.2f}")
    prev = ld
print(f"total {t[-1,3] - t[0,1]:.1f} us")
ph = ctx.debug_colour_phases()
print("mean per-CTA phase times (us): colour  n_ctas   stream+consts   wait   gather+products   reduce+draw   scatter-issue   sum")
for c in range(ctx.n_colors):
    nct = max(ph[c, 7], 1.0)
    v = ph[c, :5] / nct / 1e3
    pv = ph[c, 8:12] / nct / 1e3
    print(f"{c:4d} {int(ph[c,7]):6d}   " + "  ".join(f"{x:8.2f}" for x in v) + f"  {v.sum():8.2f}   | prologue: desc {pv[0]:.2f} site-loads {pv[1]:.2f} draw {pv[2]:.2f} stream-rest {pv[3]:.2f}")
ctx.close()
