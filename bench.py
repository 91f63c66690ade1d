#!/usr/bin/env python
"""bench.py -- headline benchmark of the NNGP hot path on B200 (contract: see the task statement / DESIGN.md section 6).

Metric (BASELINE.json): Gibbs sweeps/sec + Vecchia log-lik evals/sec at n = 1M, m = 10.  One STEP = one full chromatic
Gibbs sweep of the latent field (all colour classes, Philox normals) + one Vecchia log-likelihood evaluation, on
config 3 (synthetic U(0,1)^2 sites, n = 1 000 000, m = 10, exponential_isotropic, range 0.05, sigma^2 = 1, tau^2 = 0.1,
reordering = "random").  `value` = steps/s, whole job, inputs resident in HBM, timed with CUDA events on the library's
stream.  `e2e` = the same step through the C ABI with HOST buffers (field in, field out, host field for the log-lik).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 1000000] [--m 10]

N > 1 (torchrun, one rank per GPU), default --mode chains: one independent chain per GPU (the reference's own parallelism,
mclapply over chains, Scripts/mcmc_nngp_update_Gaussian.R:25); no data-path collective; scaling = "weak".
--mode sharded: ONE field of --n sites split over the N GPUs by spatial blocks, per-colour halo exchange of boundary values
(ncclSend/ncclRecv) and an all-reduce for the log-lik scalars; scaling = "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RANGE, SIGMA2, TAU2 = 0.05, 1.0, 0.1


def algorithmic_bytes(n, m, d=2):
    """FP64 = 8 B, int32 = 4 B, each gathered operand counted once per use, no cache credit."""
    M = m + 1
    return {
        "gibbs_sweep": n * (M * 48 + 40),
        "loglik": n * M * 20,
        "factor_build": n * M * (4 + 8 * d + 8),
        "spmv": n * M * 20 + 8 * n,
    }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(n, m, seed):
    """config 3 generator: sites, exact ordered NN, first-fit colouring; w ~ NNGP prior is drawn on the device."""
    import nngp_b200 as nb
    rng = np.random.default_rng(seed)
    locs = rng.random((n, 2))
    nn = nb.find_ordered_nn(locs, m)
    coloring = nb.greedy_coloring(nn)
    locs_match = np.arange(1, n + 1, dtype=np.int32)
    return rng, locs, nn, coloring, locs_match


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle restatement of the reference's CPU path (R + GpGp + Matrix are not in this image)
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_chain_worker(args):
    n, m, seed, steps, warmup = args
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    locs = rng.random((n, 2))
    import nngp_b200 as nb   # host set-up utilities only (no CUDA call is made in this process)
    nn = nb.find_ordered_nn(locs, m)
    coloring = nb.greedy_coloring(nn)
    Linv = O.vecchia_Linv([1.0, RANGE, 0.0], "exponential_isotropic", locs, nn)
    pd = O.precision_diag(Linv, nn)
    w = O.sparse_chol_solve(Linv, nn, rng.standard_normal(n))
    y = w + np.sqrt(TAU2) * rng.standard_normal(n)
    lm = np.arange(1, n + 1, dtype=np.int32)
    rs = O.residuals_sum(lm, n, y, np.zeros(n))
    field = w.copy()
    times = []
    for s in range(warmup + steps):
        z = rng.standard_normal(n)
        t0 = time.perf_counter()
        field = O.chromatic_sweep(Linv, nn, coloring, pd, np.ones(n), rs, 0.0, np.log(SIGMA2), np.log(TAU2), z, field, form="reference")
        ll = O.ll_compressed_sparse_chol(Linv, field, nn, np.log(SIGMA2))
        times.append(time.perf_counter() - t0)
    assert np.isfinite(ll)
    return times[warmup:]


def oracle_steps_per_sec(n, m, steps, warmup, procs):
    import multiprocessing as mp
    args = [(n, m, 100 + k, steps, warmup) for k in range(procs)]
    if procs == 1:
        res = [_oracle_chain_worker(args[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_oracle_chain_worker, args)
    per_chain = [len(t) / sum(t) for t in res]
    return float(sum(per_chain)), float(np.mean([np.mean(t) for t in res]) * 1e3)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, a.ref_procs))
    # bounded sample: the reference-form sweep costs K full mat-vecs, ~1 s/step/chain at n = 1M
    n = a.n if a.ref_n is None else a.ref_n
    steps = max(1, min(a.steps, 3))
    warmup = min(a.warmup, 1)
    value, ms = oracle_steps_per_sec(n, a.m, steps, warmup, procs)
    scale = n / a.n
    line = {
        "impl": "reference", "metric": "gibbs_sweep_plus_vecchia_loglik_per_sec", "value": value * scale, "unit": "steps/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms / scale, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config3: U(0,1)^2, n={a.n}, m={a.m}, exponential_isotropic range {RANGE}, reordering=random",
                   "step": "1 chromatic Gibbs sweep (reference form: one sparse mat-vec per colour) + 1 Vecchia log-lik"},
        "cpu_baseline": {"value": value * scale, "unit": "steps/s", "cores": procs, "kind": "port",
                         "sample": f"oracle restatement (not R/GpGp): {procs} independent chains (one per core, as mclapply does), "
                                   f"n={n}, {steps} steps each" + ("" if scale == 1 else f", scaled by n/{a.n}")},
        "e2e": {"value": value * scale, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import nngp_b200 as nb
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the NNGP hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n, m = a.n, a.m
    rng, locs, nn, coloring, locs_match = build_problem(n, m, seed=1 + rank)
    ctx = nb.NNGPContext(locs, nn, coloring, locs_match, "exponential_isotropic", device=local)
    assert ctx.factor_build([SIGMA2 * 0 + 1.0, RANGE, 0.0]) == 0
    ctx.factor_commit()
    ctx.field_init(0.0, np.log(SIGMA2), rng.standard_normal(n))          # w ~ NNGP prior by a triangular solve
    w = ctx.field_get()
    y = w + np.sqrt(TAU2) * rng.standard_normal(n)
    ctx.obs_set(y)
    beta_0, ls, lnv = 0.0, float(np.log(SIGMA2)), float(np.log(TAU2))
    ctx.gibbs_sweep(beta_0, ls, lnv, n_sweeps=1, seed=rank)               # sets the sweep parameters used by time_op

    # ---- device-resident timing: W warm-up steps, then exactly K steps, CUDA events on the library's stream ----
    ctx.time_op("sweep_loglik", reps=max(a.warmup, 3))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = nb.launch_count()
    barrier()
    t0 = time.perf_counter()
    ms_steps, launches_per_step = ctx.time_op("sweep_loglik", reps=a.steps)
    barrier()
    wall = time.perf_counter() - t0
    launches = nb.launch_count() - launches0
    total_ms = float(ms_steps.sum())
    # component timings (same stream, same events)
    ms_sweep, nl_sweep = ctx.time_op("gibbs_sweep", reps=max(10, a.steps))
    ms_ll, _ = ctx.time_op("loglik", reps=max(10, a.steps))
    ms_fac, _ = ctx.time_op("factor_build", reps=max(5, a.steps // 2))
    ms_solve, _ = ctx.time_op("sptrsv", reps=5)
    ms_commit, _ = ctx.time_op("commit", reps=5)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * a.steps / (total_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers: field in -> sweep -> field out, host field -> log-lik ----
    # The host side of every step is a page-locked buffer (nngp_host_alloc), as the contract asks; the same loop with
    # ordinary numpy arrays (staged through the library's pinned buffer by a multi-threaded copy) is reported next to it.
    def e2e_loop(field_h, steps):
        ll = 0.0
        for _ in range(steps):
            ctx.field_set(field_h)                            # H2D 8n bytes
            ctx.gibbs_sweep(beta_0, ls, lnv, 1, seed=rank)
            ctx.field_get(out=field_h)                        # D2H 8n bytes
            ll = ctx.loglik_host(field_h, ls)                 # H2D 8n bytes, D2H 2 scalars
        return ll

    def timed_e2e(field_h):
        e2e_loop(field_h, 2)
        barrier()
        t0 = time.perf_counter()
        ll = e2e_loop(field_h, a.steps)
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert np.isfinite(ll)
        return world * a.steps / dt

    # ---- config 5 flavour: the full reference iteration (2 factor rebuilds, ancillary solve, 10 sweeps, ...) behind one call ----
    var_y = float(np.var(y, ddof=1))
    chain_params = {"shape": [np.log(RANGE)], "beta_0": 0.0, "log_scale": ls, "log_noise_variance": lnv}
    ctx.chain_run(chain_params, 3, var_y, thin=0.0, n_chromatic=10, iter_start=5000, chain_index=1 + rank, keep_field=False)
    n_chain_it = 20
    barrier()
    t0 = time.perf_counter()
    ctx.chain_run(chain_params, n_chain_it, var_y, thin=0.0, n_chromatic=10, iter_start=5000, chain_index=1 + rank, keep_field=False)
    barrier()
    chain_it_per_s = world * n_chain_it / (time.perf_counter() - t0)
    ctx.field_set(w)

    # ---- config 5 flavour: mcmc_nngp_predict_field at n new sites (joint context over 2n sites; only the new rows are solved) ----
    pred_per_s = None
    if world == 1 and not a.no_predict:
        new_locs = np.random.default_rng(99).random((n, 2))
        joint = np.vstack([locs, new_locs])
        nn_j = nb.find_ordered_nn(joint, m)
        with nb.NNGPContext(joint, nn_j, np.zeros(2 * n, dtype=np.int32), np.zeros(0, dtype=np.int32), "exponential_isotropic", device=local) as pctx:
            pctx.factor_build([1.0, RANGE, 0.0])
            zp = np.random.default_rng(5).standard_normal(n)
            pctx.predict_sample(n, w, beta_0, ls, zp)
            t0 = time.perf_counter()
            for _ in range(5):
                pctx.predict_sample(n, w, beta_0, ls, zp)
            pred_per_s = 5 / (time.perf_counter() - t0)

    pinned = nb.PinnedArray(n)
    pinned.array[:] = ctx.field_get()
    e2e_value = timed_e2e(pinned.array)
    e2e_pageable = timed_e2e(ctx.field_get())
    pinned.free()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and n == 1_000_000 and m == 10:
        with open(tpath) as f:
            traffic = json.load(f)["bytes"].get("gibbs_sweep_full")
    ab = algorithmic_bytes(n, m)
    sweep_ms = float(np.mean(ms_sweep))
    achieved = ab["gibbs_sweep"] / (sweep_ms * 1e-3) / 1e9
    line = {
        "metric": "gibbs_sweep_plus_vecchia_loglik_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config3: U(0,1)^2, n={n}, m={m}, exponential_isotropic range {RANGE}, sigma2 {SIGMA2}, tau2 {TAU2}, reordering=random",
                   "step": "1 chromatic Gibbs sweep (all colours, Philox normals) + 1 Vecchia log-lik evaluation",
                   "parallelism": "1 chain per GPU" if world > 1 else "single GPU",
                   "l2": "working set (factor + indices + transpose map, ~0.4 GB) exceeds the 126 MB L2; no explicit flush",
                   "n_colors": ctx.n_colors, "solve_levels": ctx.n_levels, "layout": ctx.layout},
        "gibbs_sweeps_per_sec": world * 1e3 / sweep_ms, "loglik_evals_per_sec": world * 1e3 / float(np.mean(ms_ll)),
        "factor_builds_per_sec": world * 1e3 / float(np.mean(ms_fac)),
        "chain_iterations_per_sec": chain_it_per_s,
        "predicted_field_samples_per_sec": pred_per_s,
        "predicted_field_sample": f"nngp_predict_sample through the C ABI with host buffers: one stored sample conditionally simulated at {n} new sites",
        "chain_iteration": "nngp_chain_run: reference loop update_Gaussian.R:101-314 (2 factor rebuilds, ancillary SpMV+SpTRSV, 2 log-liks, beta_0, 10 sweeps, noise steps), whole job",
        "ms": {"sweep": sweep_ms, "loglik": float(np.mean(ms_ll)), "factor_build": float(np.mean(ms_fac)),
               "spmv_plus_sptrsv": float(np.mean(ms_solve)), "accept_transpose_precision_diag": float(np.mean(ms_commit)),
               "wall_timed_region": wall * 1e3},
        "roofline": {"bound": "hbm", "kernel": "gibbs_tile2_kernel (the K colour launches of one sweep, PDL-chained, replayed from one CUDA graph; 6 CTAs/SM build for colours that would not fit one wave at 5)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "traffic_source": "profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum over the colour launches of one sweep, --cache-control none)",
                     "algorithmic_bytes_per_sweep": ab["gibbs_sweep"],
                     "loglik_GBps": ab["loglik"] / (float(np.mean(ms_ll)) * 1e-3) / 1e9,
                     "factor_build_GBps": ab["factor_build"] / (float(np.mean(ms_fac)) * 1e-3) / 1e9},
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": 16 * n + 64, "d2h_bytes_per_step": 8 * n + 16,
                "what": "nngp_field_set + nngp_gibbs_sweep + nngp_field_get + nngp_loglik_host; host side = pinned buffer (nngp_host_alloc)",
                "pageable_numpy_value": e2e_pageable},
        "gpu_launches": int(launches), "launches_per_step": int(launches_per_step), "clocks": clocks,
    }
    if world == 1 and not a.no_cpu_baseline:
        v, ms = oracle_steps_per_sec(n, m, steps=2, warmup=0, procs=1)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": 1, "kind": "port",
                                "sample": f"oracle restatement (not R/GpGp), 1 thread, full n={n}: 2 steps of reference-form sweep + log-lik"}
    print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def run_sharded(a):
    """One field over all ranks.  Every rank builds the (replicated) host-side structure, keeps its block."""
    import torch
    import torch.distributed as dist
    import nngp_b200 as nb
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, m = a.n, a.m
    t_setup = time.perf_counter()
    # the (replicated) host-side structure is built once, by rank 0 with all host cores, and shared through /dev/shm
    shm = f"/dev/shm/nngp_bench_{os.environ.get('MASTER_PORT', '0')}_{n}_{m}"
    if rank == 0:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)           # torchrun pins it to 1; the library is not loaded yet
        rng, locs, nn, coloring, locs_match = build_problem(n, m, seed=1)
        np.save(shm + "_locs.npy", locs); np.save(shm + "_nn.npy", nn); np.save(shm + "_col.npy", coloring)
    dist.barrier()
    if rank != 0:
        locs, nn, coloring = np.load(shm + "_locs.npy"), np.load(shm + "_nn.npy"), np.load(shm + "_col.npy")
        locs_match = np.arange(1, n + 1, dtype=np.int32)
    dist.barrier()
    if rank == 0:
        for suffix in ("_locs.npy", "_nn.npy", "_col.npy"):
            os.remove(shm + suffix)
    ctx, plan = nb.create_sharded_distributed(locs, nn, coloring, locs_match, "exponential_isotropic", local, dist, transport=a.transport)
    t_setup = time.perf_counter() - t_setup
    assert ctx.factor_build([1.0, RANGE, 0.0]) == 0
    ctx.factor_commit()
    w = np.random.default_rng(7).standard_normal(n) * 0.5                    # any field will do for throughput; same on all ranks
    y = w + np.sqrt(TAU2) * np.random.default_rng(8).standard_normal(n)
    ctx.field_set(w[plan["local_sites"]])
    ctx.obs_set(y[plan["obs_index"]])
    beta_0, ls, lnv = 0.0, float(np.log(SIGMA2)), float(np.log(TAU2))
    ctx.gibbs_sweep(beta_0, ls, lnv, n_sweeps=1, seed=1)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    ctx.time_op("sweep_loglik", reps=max(a.warmup, 3))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = nb.launch_count()
    barrier()
    ms_steps, launches_per_step = ctx.time_op("sweep_loglik", reps=a.steps)
    barrier()
    launches = nb.launch_count() - launches0
    ms_sweep, _ = ctx.time_op("gibbs_sweep", reps=max(10, a.steps))
    ms_ll, _ = ctx.time_op("loglik", reps=max(10, a.steps))
    ms_fac, _ = ctx.time_op("factor_build", reps=5)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([float(ms_steps.sum()), float(ms_sweep.mean()), float(ms_ll.mean()), float(ms_fac.mean())], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, sweep_ms, ll_ms, fac_ms = [float(v) for v in t.tolist()]
    halo = torch.tensor([float(plan["send_ptr"][-1]), float(plan["n_ghost"]), float(plan["n_owned"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(halo, op=dist.ReduceOp.SUM)
    if rank == 0:
        peak, peak_src = measured_peaks()
        ab = algorithmic_bytes(n, m)
        achieved = ab["gibbs_sweep"] / (sweep_ms * 1e-3) / 1e9
        line = {
            "metric": "gibbs_sweep_plus_vecchia_loglik_per_sec", "value": a.steps / (total_ms * 1e-3), "unit": "steps/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"one field, U(0,1)^2, n={n}, m={m}, exponential_isotropic range {RANGE}, reordering=random",
                       "step": "1 chromatic Gibbs sweep of the whole field (per-colour NCCL halo exchange) + 1 Vecchia log-lik (all-reduce)",
                       "parallelism": f"field sharded over {world} GPUs by spatial blocks", "transport": a.transport, "n_colors": ctx.n_colors,
                       "halo_values_per_sweep": int(halo[0].item()), "ghost_sites_total": int(halo[1].item()), "setup_s": t_setup},
            "gibbs_sweeps_per_sec": 1e3 / sweep_ms, "loglik_evals_per_sec": 1e3 / ll_ms, "factor_builds_per_sec": 1e3 / fac_ms,
            "ms": {"sweep": sweep_ms, "loglik": ll_ms, "factor_build": fac_ms},
            "roofline": {"bound": "hbm", "kernel": "gibbs_tile2_kernel + halo exchange (whole-field sweep, all ranks)", "achieved": achieved,
                         "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world), "traffic": None, "peak_source": peak_src + f" x {world} GPUs"},
            "e2e": None, "gpu_launches": int(launches), "launches_per_step": int(launches_per_step), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sites", "--n", dest="n", type=int, default=1_000_000)
    ap.add_argument("--nbrs", "--m", dest="m", type=int, default=10)
    ap.add_argument("--ref-n", type=int, default=None, help="reference arm: run on a smaller n and scale (bounded sample)")
    ap.add_argument("--ref-procs", type=int, default=32, help="reference arm: at most this many independent chains (one per host core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-predict", action="store_true")
    ap.add_argument("--mode", default="chains", choices=["chains", "sharded"])
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="sharded mode: halo transport")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.mode == "sharded":
        run_sharded(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
