"""Extract the printed values of the reference's rendered vignette into tests/golden/vignette_golden.json.

Run in the build container only (reads /root/reference/Vignette.md, which does not exist on the GPU box):
    python tests/golden/make_vignette_golden.py
The JSON it writes is committed; tests read the JSON, never /root/reference.
"""
import json
import os
import re

SRC = "/root/reference/Vignette.md"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vignette_golden.json")


def body(lines, a, b):
    """lines a..b (1-based, inclusive) with the leading '    ## ' stripped"""
    return [re.sub(r"^\s*##\s?", "", l.rstrip("\n")) for l in lines[a - 1:b]]


def matrix_rows(rows):
    out = []
    for r in rows:
        r = re.sub(r"^\s*\[\d+,\]", "", r)
        out.append([None if t == "NA" else float(t) for t in r.split()])
    return out


def flat_ints(rows):
    out = []
    for r in rows:
        r = re.sub(r"^\s*\[\d+\]", "", r)
        out += [int(t) for t in r.split()]
    return out


def main():
    with open(SRC) as f:
        lines = f.readlines()
    g = {"source": "Vignette.md of the reference (rendered 2021-06-14); line numbers in each key's comment"}
    g["observed_locs_head"] = matrix_rows(body(lines, 137, 142))           # Vignette.md:136-142
    g["locs_head"] = matrix_rows(body(lines, 149, 154))                    # Vignette.md:148-154
    g["X_head"] = matrix_rows(body(lines, 181, 186))                       # Vignette.md:180-186 (slope, white_noise; centred)
    nn = matrix_rows(body(lines, 222, 227))                                # Vignette.md:221-227
    g["NNarray_head"] = [[None if v is None else int(v) for v in r] for r in nn]
    adj = []
    for r in body(lines, 277, 306):                                        # Vignette.md:275-306
        r = re.sub(r"^\s*\[\d+,\]", "", r)
        adj.append([1 if t == "1" else 0 for t in r.split()])
    assert len(adj) == 30 and all(len(r) == 30 for r in adj)
    g["MRF_adjacency_30"] = adj
    g["locs_match_100"] = flat_ints(body(lines, 322, 328))                 # Vignette.md:322-328
    # hctam_scol_1: alternating name / value lines                         # Vignette.md:406-419
    vals = []
    for k, r in enumerate(body(lines, 406, 419)):
        if k % 2 == 1:
            vals += [int(t) for t in r.split()]
    assert len(vals) == 100, len(vals)
    g["hctam_scol_1_100"] = vals
    assert len(g["locs_match_100"]) == 100
    # ---- initial state of chain_1 as printed right after mcmc_nngp_initialize                 # Vignette.md:472-523
    def scalars(a, b):
        return [float(t) for r in body(lines, a, b) for t in re.sub(r"^\s*\[\d+\]", "", r).split() if re.match(r"^-?\d", t)]
    g["init_chain_1"] = {
        "beta_0": scalars(476, 476)[0],                                                        # :475-476
        "beta": scalars(483, 483),                                                             # :482-483
        "log_scale": scalars(489, 489)[0],                                                     # :489
        "shape": scalars(495, 495),                                                            # :495
        "log_noise_variance": scalars(501, 501)[0],                                            # :501
        "field_100": scalars(507, 523),                                                        # :507-523
    }
    assert len(g["init_chain_1"]["field_100"]) == 100 and len(g["init_chain_1"]["beta"]) == 2
    # ---- every Gelman-Rubin-Brooks block the vignette prints: Multivariate, beta_0, slope, white_noise, log_scale,
    # log_noise_variance, shape.  Blocks 1-5: first run (5 x 200 iterations, n_chromatic 5, thinning .01; :642-684);
    # 6-31: the run until all univariate values < 1.05 (26 x 100, thinning .2; :687-873); 32-41: 10 x 100 more (:879-953);
    # 42-46: the second example, the same regressors passed as X_obs (5 x 200; :1139-1178).
    blocks = []
    i = 0
    while i < len(lines):
        if "Multivariate" in lines[i] and "beta_0" in lines[i]:
            blocks.append({"line": i + 1, "R_hat": [float(t) for t in body(lines, i + 2, i + 2)[0].split()]
                           + [float(t) for t in body(lines, i + 4, i + 4)[0].split()]})
            i += 4
        else:
            i += 1
    assert len(blocks) == 46 and all(len(b["R_hat"]) == 7 for b in blocks)
    g["R_hat_blocks"] = blocks
    # ---- mcmc_nngp_estimate(burn_in = .5) after the 4600 iterations                           # Vignette.md:996-1028
    def table(a, b):
        return [[float(t) for t in r.split()[-5:]] for r in body(lines, a, b)]
    g["estimate"] = {"GpGp_covparams": table(1000, 1002),                                       # scale, noise_variance, range
                     "fixed_effects": table(1009, 1011),                                        # beta_0, slope, white_noise
                     "field_head": table(1022, 1027)}
    # posterior summary bands (Vignette.md:999-1002,1008-1011): acceptance bands for the end-to-end toy run
    g["posterior_bands"] = {"scale": [11.32, 8.29, 15.86], "noise_variance": [5.217, 4.83, 5.64], "range": [6.21, 4.19, 9.43]}
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
