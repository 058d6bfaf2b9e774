"""Drop-in for the reference's Python module `maxpro` (pyo3, src/lib.rs:166-176).

Same six functions, argument order, defaults and ValueError texts as the pyo3
wrappers; the work is done by libmaxpro_b200.so on a B200.

    import maxpro_b200.maxpro as maxpro
    lhd = maxpro.build_lhd(50, 3, 100000, "maxpro", 42)

Differences a caller can observe ("What is not identical"):
  * random draws come from the LHD-Philox v1 stream, not rand's StdRng, so a given
    seed yields a different (equally distributed) design than the Rust build;
  * anneal_lhd runs ANNEAL_CHAINS independent chains and returns the best (1 = the
    reference's single chain).
"""
from __future__ import annotations

import secrets

from . import engine
from ._lib import METRIC_MAXIMIN, METRIC_MAXPRO, MaxproError

# Chains per anneal_lhd call.  The reference runs one; more chains cost the same
# wall time on a B200 until the SMs are full.
ANNEAL_CHAINS = 1

_U64_MAX = (1 << 64) - 1


def _u64(name: str, v) -> int:
    # pyo3 extracts u64: negative or non-integer raises (OverflowError/TypeError there)
    if isinstance(v, bool) or not isinstance(v, int):
        raise TypeError(f"argument '{name}': '{type(v).__name__}' object cannot be interpreted as an integer")
    if v < 0 or v > _U64_MAX:
        raise OverflowError(f"argument '{name}': out of range for u64")
    return v


def _seed(seed) -> int:
    return secrets.randbits(64) if seed is None else _u64("seed", seed)  # rand::random::<u64>()


def _rows(design):
    rows = [list(map(float, r)) for r in design]
    return rows


def _tolist(a):
    return a.tolist()


def _wrap(fn, *a, **k):
    try:
        return fn(*a, **k)
    except MaxproError as e:
        if e.code == 1:
            raise ValueError(e.message) from None
        raise


def generate_lhd(n_samples, n_dim, seed=None):
    """src/lhd.rs:189-213"""
    n, d = _u64("n_samples", n_samples), _u64("n_dim", n_dim)
    if n == 0:
        raise ValueError("n_samples must be a positive, nonzero integer")
    if d == 0:
        raise ValueError("n_dim must be a positive, nonzero integer")
    return _tolist(_wrap(engine.generate_lhd, n, d, _seed(seed)))


def build_lhd(n_samples, n_dim, n_iterations, metric, seed=None):
    """src/lib.rs:114-163"""
    n, d, it = _u64("n_samples", n_samples), _u64("n_dim", n_dim), _u64("n_iterations", n_iterations)
    if n == 0:
        raise ValueError("n_samples must be positive and nonzero")
    if d == 0:
        raise ValueError("n_dim must be positive and nonzero")
    if it == 0:
        raise ValueError("n_iterations must be positive and nonzero")
    if not isinstance(metric, str):
        raise TypeError("argument 'metric': expected str")
    s = _seed(seed)
    if metric == "maxpro":  # case-sensitive here (lib.rs:147-149)
        m = METRIC_MAXPRO
    elif metric == "maximin":
        m = METRIC_MAXIMIN
    else:
        raise ValueError("Metric must be 'maxpro' or 'maximin'.")
    design, _, _ = _wrap(engine.build_lhd, n, d, it, m, s)
    return _tolist(design)


def maxpro_criterion(design):
    """src/maxpro_utils.rs:93-99"""
    rows = _rows(design)
    if len(rows) == 0:
        raise ValueError("Design cannot be empty")
    return _wrap(engine.maxpro_criterion, rows)


def maximin_criterion(design):
    """src/maximin_utils.rs:111-117"""
    rows = _rows(design)
    if len(rows) == 0:
        raise ValueError("Design cannot be empty")
    return _wrap(engine.maximin_criterion, rows)


def _metric_by_name(metric_name: str):
    if not isinstance(metric_name, str):
        raise TypeError("argument 'metric_name': expected str")
    low = metric_name.lower()  # anneal.rs:180, order.rs:168
    if low == "maxpro":
        return METRIC_MAXPRO
    if low == "maximin":
        return METRIC_MAXIMIN
    raise ValueError(f"Unknown metric: '{metric_name}'. Available metrics are 'maxpro' and 'maximin'.")


def anneal_lhd(design, n_iterations, initial_temp, cooling_rate, metric_name, minimize, swap, seed=None):
    """src/anneal.rs:161-204"""
    rows = _rows(design)
    it = _u64("n_iterations", n_iterations)
    if it == 0:
        raise ValueError("n_iterations must be positive and nonzero")
    if float(initial_temp) <= 0.0:
        raise ValueError("initial_temp must be positive")
    m = _metric_by_name(metric_name)
    if not isinstance(minimize, bool) or not isinstance(swap, bool):
        raise TypeError("argument 'minimize'/'swap': expected bool")  # strict pyo3 bool
    s = _seed(seed)
    if len(rows) == 0:
        return rows  # anneal.rs:69-71
    out, _, _ = _wrap(engine.anneal_lhd, rows, it, float(initial_temp), float(cooling_rate), m, minimize, s, swap,
                      ANNEAL_CHAINS)
    return _tolist(out)


def order_design(design, metric_name):
    """src/order.rs:161-179"""
    rows = _rows(design)
    if len(rows) == 0:
        return rows  # order.rs:163-165
    m = _metric_by_name(metric_name)
    if len(rows[0]) == 0:
        return rows  # order.rs:42-46
    out, _ = _wrap(engine.order_design, rows, m, m == METRIC_MAXPRO)  # order.rs:168-170
    return _tolist(out)


__all__ = ["generate_lhd", "build_lhd", "maxpro_criterion", "maximin_criterion", "anneal_lhd", "order_design"]
