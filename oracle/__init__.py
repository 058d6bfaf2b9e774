"""ctypes front-end of the CPU oracle (oracle/maxpro_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(maxpro_b200/) never imports this module.

Parity status: criteria / tie rules / annealer / ordering are pinned by the
reference's own vectors (tests/test_oracle.py); the random stream is
"parity unpinned" (see the header of maxpro_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

MAXPRO = 0
MAXIMIN = 1


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "maxpro_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "clean", "all"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, u64, i32, dbl = C.POINTER(C.c_double), C.c_uint64, C.c_int, C.c_double
        u64p = C.POINTER(C.c_uint64)
        L.oracle_maxpro_sum.restype = dbl
        L.oracle_maxpro_sum.argtypes = [dp, u64, u64]
        L.oracle_maxpro_criterion.restype = dbl
        L.oracle_maxpro_criterion.argtypes = [dp, u64, u64]
        L.oracle_maximin_criterion.restype = dbl
        L.oracle_maximin_criterion.argtypes = [dp, u64, u64]
        L.oracle_l2_distance.restype = dbl
        L.oracle_l2_distance.argtypes = [dp, dp, u64]
        L.oracle_score_batch.restype = None
        L.oracle_score_batch.argtypes = [dp, u64, u64, u64, i32, dp, u64p]
        L.oracle_philox4x32_10.restype = None
        L.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.oracle_generate_lhd.restype = i32
        L.oracle_generate_lhd.argtypes = [u64, u64, u64, dp]
        L.oracle_generate_batch.restype = i32
        L.oracle_generate_batch.argtypes = [u64, u64, u64, u64, u64, dp]
        L.oracle_search_range.restype = i32
        L.oracle_search_range.argtypes = [u64, u64, i32, u64, u64, u64, dp, u64p]
        L.oracle_build_lhd.restype = i32
        L.oracle_build_lhd.argtypes = [u64, u64, u64, i32, u64, dp, dp, u64p]
        L.oracle_anneal_lhd.restype = i32
        L.oracle_anneal_lhd.argtypes = [dp, u64, u64, u64, dbl, dbl, i32, i32, u64, i32, dp, dp]
        L.oracle_order_design.restype = i32
        L.oracle_order_design.argtypes = [dp, u64, u64, i32, i32, dp, u64p]
        L.oracle_order_design_incremental.restype = i32
        L.oracle_order_design_incremental.argtypes = [dp, u64, u64, i32, i32, dp, u64p]
        L.oracle_set_stream_version.restype = None
        L.oracle_set_stream_version.argtypes = [i32]
        L.oracle_get_stream_version.restype = i32
        L.oracle_num_threads.restype = i32
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [i32]
        _lib = L
    return _lib


def set_stream_version(v: int) -> None:
    """Design stream the oracle generates: 1 = LHD-Philox v1, 2 = LHD-mix v2 (default, as in the engine)."""
    lib().oracle_set_stream_version(int(v))


def get_stream_version() -> int:
    return int(lib().oracle_get_stream_version())


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _as2d(design) -> np.ndarray:
    a = np.ascontiguousarray(design, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("design must be 2-D (n x d)")
    return a


def maxpro_sum(design) -> float:
    a = _as2d(design)
    return lib().oracle_maxpro_sum(_dp(a), a.shape[0], a.shape[1])


def maxpro_criterion(design) -> float:
    a = _as2d(design)
    if a.shape[0] == 0:
        raise ValueError("Design cannot be empty")
    return lib().oracle_maxpro_criterion(_dp(a), a.shape[0], a.shape[1])


def maximin_criterion(design) -> float:
    a = _as2d(design)
    if a.shape[0] == 0:
        raise ValueError("Design cannot be empty")
    return lib().oracle_maximin_criterion(_dp(a), a.shape[0], a.shape[1])


def l2_distance(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape and a.size > 0
    return lib().oracle_l2_distance(_dp(a), _dp(b), a.size)


def score_batch(designs, metric: int):
    a = np.ascontiguousarray(designs, dtype=np.float64)
    b, n, d = a.shape
    scores = np.empty(b, dtype=np.float64)
    arg = C.c_uint64(0)
    lib().oracle_score_batch(_dp(a), b, n, d, metric, _dp(scores), C.byref(arg))
    return scores, int(arg.value)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def generate_lhd(n: int, d: int, stream: int) -> np.ndarray:
    out = np.empty((n, d), dtype=np.float64)
    rc = lib().oracle_generate_lhd(n, d, stream & (2**64 - 1), _dp(out))
    if rc:
        raise ValueError(f"oracle_generate_lhd failed rc={rc}")
    return out


def generate_batch(n: int, d: int, seed: int, lo: int, count: int) -> np.ndarray:
    """count consecutive candidates of the search seeded `seed`, from candidate `lo`: shape (count, n, d)."""
    out = np.empty((count, n, d), dtype=np.float64)
    if lib().oracle_generate_batch(n, d, seed & (2 ** 64 - 1), lo, count, _dp(out)):
        raise ValueError("oracle_generate_batch failed")
    return out


def search_range(n, d, metric, seed, lo, hi):
    s = C.c_double(0)
    i = C.c_uint64(0)
    rc = lib().oracle_search_range(n, d, metric, seed, lo, hi, C.byref(s), C.byref(i))
    if rc:
        raise ValueError(f"oracle_search_range failed rc={rc}")
    return s.value, int(i.value)


def build_lhd(n, d, iters, metric, seed):
    out = np.empty((n, d), dtype=np.float64)
    s = C.c_double(0)
    i = C.c_uint64(0)
    rc = lib().oracle_build_lhd(n, d, iters, -1 if metric is None else metric, seed, _dp(out), C.byref(s), C.byref(i))
    if rc:
        raise ValueError(f"oracle_build_lhd failed rc={rc}")
    return out, s.value, int(i.value)


def anneal_lhd(design, iters, initial_temp, cooling_rate, metric, minimize, seed, swap):
    a = _as2d(design)
    out = np.empty_like(a)
    best = C.c_double(0)
    rc = lib().oracle_anneal_lhd(_dp(a), a.shape[0], a.shape[1], iters, initial_temp, cooling_rate,
                                 metric, int(bool(minimize)), seed, int(bool(swap)), _dp(out), C.byref(best))
    if rc:
        raise ValueError(f"oracle_anneal_lhd failed rc={rc}")
    return out, best.value


def order_design(design, metric, minimize, incremental=False):
    a = _as2d(design)
    out = np.empty_like(a)
    perm = np.empty(a.shape[0], dtype=np.uint64)
    fn = lib().oracle_order_design_incremental if incremental else lib().oracle_order_design
    rc = fn(_dp(a), a.shape[0], a.shape[1], metric, int(bool(minimize)), _dp(out),
            perm.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc:
        raise ValueError(f"oracle_order_design failed rc={rc}")
    return out, perm


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    """Use n OpenMP threads (torchrun exports OMP_NUM_THREADS=1)."""
    lib().oracle_set_num_threads(int(n))
