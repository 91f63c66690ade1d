"""R's random stream, and the reference's ordering / neighbour search on it, behind the C ABI (include/nngp_b200.h, section
"R-compatible random stream").  A host written in R has its own RNG and GpGp; this is what lets the Python mirror reproduce
mcmc_nngp_initialize(seed) -- ordering, NNarray, colouring, initial states -- exactly as the reference produces them in R
(Scripts/mcmc_nngp_initialize.R:17,29,93,154-161,189-208)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class RStream:
    """Mersenne-Twister + inversion normals + rejection sampling, R's defaults since 3.6.  The state (625 ints) lives here
    and is passed to every call, like R's .Random.seed."""

    def __init__(self, seed: int | None = None):
        self.state = np.zeros(625, dtype=np.int32)
        if seed is not None:
            self.set_seed(seed)

    def _call(self, fn, *args):
        st = C.c_int(0)
        fn(*args, C.byref(st))
        L.check(st)

    def set_seed(self, seed: int) -> None:
        """set.seed(seed)"""
        s = int(seed) & 0xFFFFFFFF
        self._call(L.load().nngp_rng_set_seed, L.ci(s - (1 << 32) if s >= (1 << 31) else s), L.iptr(self.state))

    def runif(self, n: int) -> np.ndarray:
        out = np.empty(int(n))
        self._call(L.load().nngp_rng_runif, L.iptr(self.state), L.ci(n), L.dptr(out))
        return out

    def rnorm(self, n: int) -> np.ndarray:
        out = np.empty(int(n))
        self._call(L.load().nngp_rng_rnorm, L.iptr(self.state), L.ci(n), L.dptr(out))
        return out

    def sample_int(self, n: int, size: int | None = None) -> np.ndarray:
        """sample.int(n, size) without replacement, 1-based; size = n: sample(n)"""
        size = n if size is None else size
        out = np.empty(int(size), dtype=np.int32)
        self._call(L.load().nngp_rng_sample_int, L.iptr(self.state), L.ci(n), L.ci(size), L.iptr(out))
        return out

    def sample_one(self, x):
        """sample(x, 1) for length(x) > 1"""
        return x[int(self.sample_int(len(x), 1)[0]) - 1]

    def rbeta(self, n: int, shape1: float, shape2: float) -> np.ndarray:
        out = np.empty(int(n))
        self._call(L.load().nngp_rng_rbeta, L.iptr(self.state), L.ci(n), L.cd(shape1), L.cd(shape2), L.dptr(out))
        return out


def _locs2(locs):
    locs = np.asarray(locs, dtype=np.float64)
    return locs[:, None] if locs.ndim == 1 else locs


def order_maxmin_gpgp(locs, rs: RStream, lonlat: bool = False) -> np.ndarray:
    """GpGp::order_maxmin(locs, lonlat) on R's stream (initialize.R:29): the reference's ordering itself, 1-based"""
    locs = _locs2(locs)
    n, d = locs.shape
    out = np.empty(n, dtype=np.int32)
    st = C.c_int(0)
    L.load().nngp_host_order_maxmin_gpgp(L.dptr(L.f64(locs)), L.ci(n), L.ci(d), L.ci(1 if lonlat else 0), L.iptr(rs.state), L.iptr(out), C.byref(st))
    L.check(st)
    return out


def find_ordered_nn_gpgp(locs, m: int, rs: RStream) -> np.ndarray:
    """GpGp::find_ordered_nn(locs, m) on R's stream (initialize.R:93): coordinates jittered as GpGp does before the search"""
    locs = _locs2(locs)
    n, d = locs.shape
    out = np.empty(n * (m + 1), dtype=np.int32)
    st = C.c_int(0)
    L.load().nngp_host_find_ordered_nn_gpgp(L.dptr(L.f64(locs)), L.ci(n), L.ci(d), L.ci(m), L.iptr(rs.state), L.iptr(out), C.byref(st))
    L.check(st)
    return out.reshape((n, m + 1), order="F")
