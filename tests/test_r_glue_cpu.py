"""R is not installed in this image, so the .C() glue under R/ cannot be executed here.  This is a static check instead:
every .C("nngp_...") call names an exported ABI function and passes exactly as many arguments as the C prototype in
include/nngp_b200.h has parameters, named and ordered like them, with `status` last (.C() matches by position: a wrong
count or order corrupts memory at run time in R)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def prototypes():
    src = open(os.path.join(ROOT, "include", "nngp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for name, args in re.findall(r"^void\s+(nngp_\w+)\s*\((.*?)\)\s*;", src, flags=re.M | re.S):
        params = [a.strip() for a in args.replace("\n", " ").split(",") if a.strip() and a.strip() != "void"]
        protos[name] = [re.sub(r".*[\s\*]", "", p) for p in params]
    return protos


def split_top_level(s):
    out, depth, cur, quote = [], 0, "", None
    for ch in s:
        if quote:
            cur += ch
            if ch == quote:
                quote = None
            continue
        if ch in "\"'":
            quote = ch
            cur += ch
        elif ch in "([{":
            depth += 1
            cur += ch
        elif ch in ")]}":
            depth -= 1
            cur += ch
        elif ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def dot_c_calls(text):
    calls = []
    for mobj in re.finditer(r"\.C\(", text):
        i, depth = mobj.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        args = split_top_level(text[mobj.end():i - 1])
        name = args[0].strip().strip('"')
        named = [a.split("=")[0].strip() for a in args[1:]]
        calls.append((name, [a for a in named if a != "NAOK"]))
    return calls


def test_every_dot_c_call_matches_its_prototype():
    protos = prototypes()
    seen = set()
    for fn in sorted(os.listdir(os.path.join(ROOT, "R"))):
        if not fn.endswith(".R"):
            continue
        text = "\n".join(l.split("#")[0] for l in open(os.path.join(ROOT, "R", fn)).read().splitlines())
        for name, args in dot_c_calls(text):
            assert name in protos, (fn, name)
            assert len(args) == len(protos[name]), (fn, name, args, protos[name])
            assert args == protos[name], (fn, name, args, protos[name])      # same names in the same order as the prototype
            assert args[-1] == "status" or name == "nngp_last_error", (fn, name)
            seen.add(name)
    # the glue covers the entry points the reference's R code needs
    for need in ("nngp_ctx_create", "nngp_ctx_destroy", "nngp_factor_build", "nngp_loglik", "nngp_gibbs_sweep", "nngp_chain_run",
                 "nngp_regressors_set", "nngp_chain_run_regressors", "nngp_predict_sample", "nngp_field_init",
                 "nngp_host_find_ordered_nn", "nngp_host_greedy_coloring", "nngp_host_order_maxmin", "nngp_last_error"):
        assert need in seen, need
