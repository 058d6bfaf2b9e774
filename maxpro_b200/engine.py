"""NumPy-level API over the C ABI: arrays in, arrays out, metric as an enum.

The string-and-list surface of the reference's Python module lives in
maxpro_b200.maxpro; this module is what it (and bench.py, and the tests) call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import METRIC_MAXIMIN, METRIC_MAXPRO, METRIC_NONE, MaxproError, check, dptr, lib, u64ptr

U64 = (1 << 64) - 1


def _design(a) -> np.ndarray:
    x = np.ascontiguousarray(a, dtype=np.float64)
    if x.ndim != 2:
        raise ValueError("design must be 2-D (n x d)")
    return x


def generate_lhd(n: int, d: int, seed: int) -> np.ndarray:
    out = np.empty((n, d), dtype=np.float64)
    check(lib().maxpro_generate_lhd(n, d, seed & U64, dptr(out)))
    return out


def generate_batch(n: int, d: int, seed: int, lo: int, count: int) -> np.ndarray:
    out = np.empty((count, n, d), dtype=np.float64)
    check(lib().maxpro_generate_batch(n, d, seed & U64, lo, count, dptr(out)))
    return out


def maxpro_criterion(design) -> float:
    x = _design(design)
    out = C.c_double(0)
    check(lib().maxpro_criterion(dptr(x), x.shape[0], x.shape[1], C.byref(out)))
    return out.value


def maximin_criterion(design) -> float:
    x = _design(design)
    out = C.c_double(0)
    check(lib().maxpro_maximin_criterion(dptr(x), x.shape[0], x.shape[1], C.byref(out)))
    return out.value


def l2_distance(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    b = np.ascontiguousarray(b, dtype=np.float64).ravel()
    if a.size != b.size:
        raise ValueError("point_a and point_b must have the same length")  # maximin_utils.rs:27-31
    out = C.c_double(0)
    check(lib().maxpro_l2_distance(dptr(a), dptr(b), a.size, C.byref(out)))
    return out.value


def score_batch(designs, metric: int):
    """scores[b], argbest (later index on ties, src/lib.rs:85-97)."""
    x = np.ascontiguousarray(designs, dtype=np.float64)
    if x.ndim != 3:
        raise ValueError("designs must be 3-D (b x n x d)")
    b, n, d = x.shape
    scores = np.empty(b, dtype=np.float64)
    arg = C.c_uint64(0)
    check(lib().maxpro_score_batch(dptr(x), b, n, d, metric, dptr(scores), C.byref(arg)))
    return scores, int(arg.value)


def build_lhd(n: int, d: int, n_iterations: int, metric: int, seed: int, flags: int = 0):
    """(design, score, index) of the best of n_iterations random LHDs (src/lib.rs:30-100)."""
    out = np.empty((n, d), dtype=np.float64)
    score = C.c_double(0)
    idx = C.c_uint64(0)
    check(lib().maxpro_build_lhd_ex(n, d, n_iterations, metric, seed & U64, flags, dptr(out), C.byref(score), C.byref(idx)))
    return out, score.value, int(idx.value)


def search_range(n: int, d: int, metric: int, seed: int, lo: int, hi: int, flags: int = 0):
    score = C.c_double(0)
    idx = C.c_uint64(0)
    check(lib().maxpro_search_range_ex(n, d, metric, seed & U64, lo, hi, flags, C.byref(score), C.byref(idx)))
    return score.value, int(idx.value)


def reduce_best(metric: int, scores, indices):
    s = np.ascontiguousarray(scores, dtype=np.float64)
    i = np.ascontiguousarray(indices, dtype=np.uint64)
    bs = C.c_double(0)
    bi = C.c_uint64(0)
    check(lib().maxpro_reduce_best(metric, dptr(s), u64ptr(i), s.size, C.byref(bs), C.byref(bi)))
    return bs.value, int(bi.value)


def anneal_lhd(design, n_iterations: int, initial_temp: float, cooling_rate: float, metric: int,
               minimize: bool, seed: int, swap: bool, n_chains: int = 1):
    """(design, metric, chain) -- src/anneal.rs:56-143 run as n_chains independent chains."""
    x = _design(design)
    out = np.empty_like(x)
    m = C.c_double(0)
    ch = C.c_uint64(0)
    check(lib().maxpro_anneal_lhd(dptr(x), x.shape[0], x.shape[1], n_iterations, initial_temp, cooling_rate,
                                  metric, int(bool(minimize)), seed & U64, int(bool(swap)), n_chains,
                                  dptr(out), C.byref(m), C.byref(ch)))
    return out, m.value, int(ch.value)


def order_design(design, metric: int, minimize: bool):
    """(ordered design, perm) -- src/order.rs:34-105."""
    x = _design(design)
    out = np.empty_like(x)
    perm = np.empty(x.shape[0], dtype=np.uint64)
    check(lib().maxpro_order_design(dptr(x), x.shape[0], x.shape[1], metric, int(bool(minimize)), dptr(out), u64ptr(perm)))
    return out, perm


def fp64_peak_probe(repeats: int = 5):
    t = C.c_double(0)
    ms = C.c_double(0)
    check(lib().maxpro_fp64_peak_probe(repeats, C.byref(t), C.byref(ms)))
    return t.value, ms.value


def fp32_peak_probe(repeats: int = 5):
    t = C.c_double(0)
    ms = C.c_double(0)
    check(lib().maxpro_fp32_peak_probe(repeats, C.byref(t), C.byref(ms)))
    return t.value, ms.value
