"""Turns the ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed).

    python scripts/summarize_profiles.py r01
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

PAREN = re.compile(r"\(.*")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "dram__bytes.sum.per_second",
           "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "launch__waves_per_multiprocessor", "smsp__inst_executed.sum"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def launch_list(tag):
    path = os.path.join(SRC, "launches.csv")
    if not os.path.exists(path):
        return
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    total = 0.0
    for x in rows:
        name = PAREN.sub("", x["Kernel Name"]).replace("void ", "").strip()
        us = to_us(x["Metric Value"], x["Metric Unit"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    with open(os.path.join(OUT, f"{tag}_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (python bench.py --steps 2 --warmup 3 --no-cpu-baseline)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'mean_us':>9s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:58]:58s} {v[0]:8d} {v[1]:10.1f} {v[1] / v[0]:9.2f} {100 * v[1] / total:6.1f}%\n")
        # one timed step = K gibbs launches + advance + loglik + reduce
        ids = [i for i, x in enumerate(rows) if "advance_sweep" in x["Kernel Name"]]
        if len(ids) > 6:
            a, b = ids[4], ids[5]
            f.write("\n# launches of one timed step (sweep + log-lik), in order:\n")
            for x in rows[a + 1:b + 1]:
                f.write(f"  {PAREN.sub('', x['Kernel Name'])[:50]:50s} grid={x['Grid Size']:14s} {to_us(x['Metric Value'], x['Metric Unit']):8.2f} us\n")


def full(tag, rep, label):
    path = os.path.join(SRC, rep)
    csv_path = path.replace(".ncu-rep", "_raw.csv")
    if os.path.exists(csv_path):
        raw = open(csv_path).read()
    elif os.path.exists(path):
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    with open(os.path.join(OUT, f"{tag}_{label}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on   ({rep})\n")
        tot_r = tot_w = tot_t = 0.0
        for r in rows[2:]:
            f.write(f"\n## {PAREN.sub('', r[idx['Kernel Name']])}  grid={r[idx['Grid Size']]} block={r[idx['Block Size']]}\n")
            for m in METRICS:
                if m in idx:
                    f.write(f"  {m:62s} {r[idx[m]]:>14s} {units[idx[m]]}\n")
            st = sorted([(float(r[idx[h]]), h) for h in stall if r[idx[h]] not in ("", "n/a")], reverse=True)[:5]
            f.write("  top stalls (warp cycles per issued instruction): " + ", ".join(f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')}={v:.1f}" for v, h in st) + "\n")
            us = to_us(r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]])
            tot_t += us
            if "dram__bytes_read.sum" in idx:
                tot_r += to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
                tot_w += to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            elif "dram__bytes.sum.per_second" in idx:   # section-limited capture: bytes = rate x duration (read + write together)
                u = units[idx["dram__bytes.sum.per_second"]]
                rate = float(r[idx["dram__bytes.sum.per_second"]].replace(",", "")) * {"byte/s": 1, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}.get(u, 1)
                b = rate * us * 1e-6
                tot_r += b
                f.write(f"  dram bytes (rate x duration)                                   {b / 1e6:14.3f} MB\n")
        f.write(f"\n# totals over the {len(rows) - 2} captured launches: dram read {tot_r / 1e6:.1f} MB, dram write {tot_w / 1e6:.1f} MB, "
                f"time {tot_t:.1f} us (cold cache, serialised)\n")
    return tot_r + tot_w


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    os.makedirs(OUT, exist_ok=True)
    launch_list(tag)
    import json
    traffic = {}
    tj = os.path.join(OUT, "traffic.json")
    if os.path.exists(tj):                      # captures that were not re-taken keep their earlier value
        with open(tj) as f:
            traffic = json.load(f).get("bytes", {})
    for rep, label in (("prof_gibbs.ncu-rep", "gibbs_sweep_full"), ("prof_gibbs_first.ncu-rep", "gibbs_first_colour_full"), ("prof_factor.ncu-rep", "factor_full"),
                       ("prof_loglik.ncu-rep", "loglik_full"), ("prof_other.ncu-rep", "transpose_sptrsv_full")):
        t = full(tag, rep, label)
        if t:
            print(label, "dram traffic (MB):", t / 1e6)
            traffic[label] = t
    # bench.py reports roofline.traffic from here (dram__bytes_read.sum + dram__bytes_write.sum of one full sweep)
    sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    import datetime
    with open(os.path.join(OUT, "traffic.json"), "w") as f:
        json.dump({"source": f"{tag}: ncu, n=1M m=10 config 3 (max-min ordering)", "git_sha": sha, "when": datetime.datetime.now(datetime.timezone.utc).isoformat(),
                   "how": "dram__bytes_read.sum + dram__bytes_write.sum per capture; gibbs_sweep_full = the K colour launches of ONE warm sweep (--cache-control none); "
                          "the library that was profiled is the one built from git_sha (scripts/profile_run.sh ran the plain bench first)",
                   "bytes": traffic}, f, indent=1)
