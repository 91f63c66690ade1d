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
    # posterior summary bands (Vignette.md:999-1002,1008-1011): acceptance bands for the end-to-end toy run
    g["posterior_bands"] = {"scale": [11.32, 8.29, 15.86], "noise_variance": [5.217, 4.83, 5.64], "range": [6.21, 4.19, 9.43]}
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
