"""ctypes binding of libnngp_b200.so (the C ABI in include/nngp_b200.h).

This is the same binding an R maintainer would write with dyn.load()/.C(): every argument is a pointer, the last one is
`status`.  There is no CPU fallback here: if the shared library is missing or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

NA_INT = -2147483648
PKG_DIR = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(os.path.dirname(PKG_DIR), "lib", "libnngp_b200.so")   # built in-tree by build.py

COVFUN_IDS = {
    "exponential_isotropic": 0, "exponential_sphere": 1, "exponential_scaledim": 2, "exponential_spacetime": 3,
    "matern_isotropic": 4, "matern_sphere": 5, "matern_scaledim": 6, "matern_spacetime": 7,
}
SLOT_CURRENT, SLOT_PROPOSAL = 0, 1
RNG_SUPPLIED, RNG_PHILOX = 0, 1
LAYOUT_COLOR, LAYOUT_COLOR_MORTON, LAYOUT_MORTON = 1, 2, 3

# every symbol include/nngp_b200.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "nngp_version", "nngp_device_count", "nngp_last_error", "nngp_last_error_r", "nngp_host_greedy_coloring_adj", "nngp_host_set_num_threads", "nngp_host_find_ordered_nn", "nngp_host_greedy_coloring",
    "nngp_host_order_maxmin", "nngp_host_order_maxmin_gpgp", "nngp_host_find_ordered_nn_gpgp", "nngp_rng_set_seed", "nngp_rng_runif", "nngp_rng_rnorm", "nngp_rng_sample_int", "nngp_rng_rbeta",
    "nngp_ctx_create", "nngp_ctx_destroy", "nngp_comm_unique_id", "nngp_ctx_create_sharded", "nngp_shard_p2p_export", "nngp_shard_p2p_connect", "nngp_shard_sweep_begin", "nngp_shard_sweep_colour", "nngp_shard_halo_get", "nngp_shard_halo_put", "nngp_shard_sweep_end", "nngp_ctx_set_option", "nngp_ctx_info", "nngp_solve_timeline", "nngp_shard_timeline", "nngp_factor_build",
    "nngp_factor_get", "nngp_factor_accept", "nngp_factor_commit", "nngp_precision_diag", "nngp_field_set",
    "nngp_field_get", "nngp_obs_set", "nngp_loglik", "nngp_loglik_host", "nngp_spmv", "nngp_sptmv", "nngp_sptrsv",
    "nngp_gibbs_sweep", "nngp_sweep_loglik_host", "nngp_ancillary_propose", "nngp_ancillary_accept", "nngp_beta0_moments", "nngp_ssr",
    "nngp_field_init", "nngp_chain_run", "nngp_regressors_set", "nngp_chain_run_regressors", "nngp_records_summary", "nngp_predict_sample", "nngp_time_op", "nngp_launch_count", "nngp_host_alloc", "nngp_host_free",
    "nngp_host_spatial_blocks", "nngp_host_shard_plan_build", "nngp_host_shard_plan_get", "nngp_shard_connect_local", "nngp_shard_group_sweep",
    "nngp_shard_group_loglik", "nngp_shard_group_chain_run", "nngp_chains_run", "nngp_chains_run_regressors", "nngp_time_op_group", "nngp_fp64_peak",
]


class NNGPError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libnngp_b200 status {status}: {message}")
        self.status = status


_lib = None


def load():
    """dlopen the in-tree library; fails loudly (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise FileNotFoundError(
                f"{SO_PATH} is missing: build it with `python {os.path.join(PKG_DIR, 'build.py')}` "
                "(the NNGP hot path has no CPU fallback)")
        _lib = C.CDLL(SO_PATH)
        if os.environ.get("NNGP_QUIET") is None:   # one line so that a run's log shows which native library served it
            import sys
            print(f"[nngp_b200] loaded {os.path.realpath(SO_PATH)}", file=sys.stderr, flush=True)
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    load().nngp_last_error(buf, C.byref(C.c_int(1024)))
    return buf.value.decode(errors="replace")


def check(status: C.c_int) -> None:
    if status.value != 0:
        raise NNGPError(status.value, last_error())


def ci(v: int):
    return C.byref(C.c_int(int(v)))


def cd(v: float):
    return C.byref(C.c_double(float(v)))


def f64(a) -> np.ndarray:
    """column-major flat float64 copy/view, as R would hand the array to .C()"""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).ravel(order="F"))


def dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class PinnedArray:
    """float64 numpy array over page-locked host memory (nngp_host_alloc): vectors passed from it skip the staging copy."""

    def __init__(self, n: int):
        self._ptr = C.c_void_p(None)
        st = C.c_int(0)
        load().nngp_host_alloc(C.byref(C.c_double(8.0 * n)), C.byref(self._ptr), C.byref(st))
        check(st)
        self.array = np.ctypeslib.as_array(C.cast(self._ptr, C.POINTER(C.c_double)), shape=(n,))

    def free(self):
        if self._ptr:
            self.array = None
            st = C.c_int(0)
            load().nngp_host_free(C.byref(self._ptr), C.byref(st))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def launch_count() -> int:
    v = C.c_double(0.0)
    load().nngp_launch_count(C.byref(v))
    return int(v.value)


def device_count() -> int:
    n, st = C.c_int(0), C.c_int(0)
    load().nngp_device_count(C.byref(n), C.byref(st))
    return n.value if st.value == 0 else 0


def set_host_threads(n: int = 0) -> None:
    """OpenMP threads of the host set-up utilities (0 = all processors); torchrun exports OMP_NUM_THREADS=1"""
    st = C.c_int(0)
    load().nngp_host_set_num_threads(C.byref(C.c_int(int(n))), C.byref(st))
    check(st)
