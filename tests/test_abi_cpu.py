"""CPU-side checks of the boundary: the C-ABI library loads without a GPU, exports every symbol of include/nngp_b200.h,
fails loudly (no CPU fallback) when there is no device, and its host set-up utilities match the oracle bit-exactly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nngp_b200 as nb
from nngp_b200 import _lib as L
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    with open(os.path.join(ROOT, "include", "nngp_b200.h")) as f:
        declared = re.findall(r"^void\s+(nngp_\w+)\s*\(", f.read(), flags=re.M)
    assert sorted(declared) == sorted(L.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_no_cpu_fallback_without_device():
    if nb.device_count() > 0:
        pytest.skip("a GPU is present")
    locs = np.random.default_rng(0).random((50, 2))
    nn = nb.find_ordered_nn(locs, 3)
    with pytest.raises(nb.NNGPError) as e:
        nb.NNGPContext(locs, nn, nb.greedy_coloring(nn), np.arange(1, 51))
    assert e.value.status == 2          # NNGP_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    pkg = L.PKG_DIR
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".R")):
                with open(os.path.join(dirpath, fn), errors="replace") as f:
                    src = f.read()
                assert "liboracle" not in src and "from oracle" not in src and "import oracle" not in src, fn
                assert "nngp_oracle.h" not in src, fn


@pytest.mark.parametrize("n,m,d", [(1, 3, 2), (7, 10, 2), (600, 10, 2), (3000, 10, 2), (2500, 5, 3), (1500, 20, 2), (900, 4, 1)])
def test_find_ordered_nn_bit_exact(n, m, d):
    locs = np.random.default_rng(n + m).random((n, d))
    assert np.array_equal(nb.find_ordered_nn(locs, m), O.find_ordered_nn(locs, m))


def test_find_ordered_nn_degenerate_geometries():
    rng = np.random.default_rng(5)
    # quasi-1-D (the vignette's layout), exact duplicates of coordinates along one axis, gridded sites with many ties
    O.set_seed(1)
    u = O.runif(2000)
    quasi = np.column_stack([500 * u, np.ones(2000)])
    quasi[0, 1] = 1.01
    gx, gy = np.meshgrid(np.arange(40.0), np.arange(40.0))
    grid = np.column_stack([gx.ravel(), gy.ravel()])[rng.permutation(1600)]
    for locs, m in ((quasi, 5), (grid, 8)):
        assert np.array_equal(nb.find_ordered_nn(locs, m), O.find_ordered_nn(locs, m))


@pytest.mark.parametrize("n,m", [(1, 2), (50, 3), (2000, 10), (1500, 20)])
def test_greedy_coloring_bit_exact(n, m):
    locs = np.random.default_rng(n).random((n, 2))
    nn = O.find_ordered_nn(locs, m)
    adj_p, adj_i = O.moral_graph(nn)
    assert np.array_equal(nb.greedy_coloring(nn), O.naive_greedy_coloring(adj_p, adj_i))


def test_order_maxmin_is_exact_farthest_point():
    locs = np.random.default_rng(9).random((400, 2))
    order = nb.order_maxmin(locs) - 1
    assert sorted(order.tolist()) == list(range(400))
    D = np.sqrt(((locs[:, None] - locs[None]) ** 2).sum(-1))
    chosen = [order[0]]
    assert order[0] == np.argmin(((locs - locs.mean(0)) ** 2).sum(1))
    mind = D[order[0]].copy()
    for k in range(1, 400):
        cand = mind.copy()
        cand[chosen] = -1
        assert abs(cand[order[k]] - cand.max()) < 1e-15
        chosen.append(order[k])
        mind = np.minimum(mind, D[order[k]])


def test_bad_arguments_are_reported_not_crashed():
    st = C.c_int(0)
    L.load().nngp_host_find_ordered_nn(None, L.ci(3), L.ci(2), L.ci(1), None, C.byref(st))
    assert st.value == 1
    L.load().nngp_ctx_destroy(L.ci(12345), C.byref(st))
    assert st.value == 1 and "unknown context" in L.last_error()


@pytest.mark.parametrize("n,m,d,order", [(120000, 10, 2, "random"), (30000, 20, 2, "random"), (20000, 10, 2, "maxmin"), (20000, 31, 2, "random")])
def test_parallel_coloring_equals_the_sequential_first_fit(n, m, d, order, monkeypatch):
    """The blocked-parallel colouring (mask of final colours + in-block dependencies resolved in index order) is the
    sequential first-fit loop of Coloring.R:2-20 bit for bit, including the fall-back when 63 colours do not suffice
    (m = 31) and a max-min ordering, whose first sites are moral neighbours of almost everything."""
    rng = np.random.default_rng(n + m)
    locs = rng.random((n, d))
    if order == "maxmin":
        locs = locs[nb.order_maxmin(locs) - 1]
    nn = nb.find_ordered_nn(locs, m)
    par = nb.greedy_coloring(nn)
    monkeypatch.setenv("NNGP_COLORING_MASK_COLOURS", "12")       # fewer mask bits than colours: the fall-back path
    fb = nb.greedy_coloring(nn)
    monkeypatch.delenv("NNGP_COLORING_MASK_COLOURS")
    monkeypatch.setenv("NNGP_COLORING_SEQUENTIAL", "1")
    seq = nb.greedy_coloring(nn)
    assert np.array_equal(par, seq) and np.array_equal(fb, seq)
    assert par.min() == 1


def test_abi_links_and_runs_from_plain_c(tmp_path):
    """include/nngp_b200.h is a C header and the library a C library: tests/c_abi/abi_from_c.c (C11) is compiled with gcc against
    them and run -- version, R's random stream, the host builders, the error convention, and the refusal of a compute entry
    point without a device (no CPU fallback)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    L.load()
    lib_dir = os.path.dirname(L.SO_PATH)
    exe = str(tmp_path / "abi_from_c")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c_abi", "abi_from_c.c"), "-o", exe, "-L" + lib_dir, "-lnngp_b200", "-lm",
                        "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "FAIL" not in r.stdout, r.stdout + r.stderr
    assert r.stdout.count("ok ") >= 7
