"""Random (shape, design kind, temperature, direction, seed) runs of the maximin swap annealer: the screened step
(default) against the kernel with the exact row pass; (design, best, chain) must be identical.
usage: python tools/anneal_maximin_stress.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from maxpro_b200 import engine, METRIC_MAXIMIN, _lib
import oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(2026)
t_end = time.time() + budget
runs = bad = 0
while time.time() < t_end:
    n = int(rng.choice([2, 3, 5, 17, 33, 64, 100, 257, 500, 1000, int(rng.integers(2, 1200))]))
    d = int(rng.choice([1, 2, 3, 4, 5, 8, 13, 20, 33, int(rng.integers(1, 40))]))
    if n * d * 8 > 180_000:
        continue
    x = oracle.generate_lhd(n, d, int(rng.integers(1, 1 << 30)))
    kind = int(rng.integers(0, 6))
    if kind == 1:
        x = np.floor(x * 5.0) / 5.0
    elif kind == 2:
        x[:, 0] *= 10.0 ** float(rng.integers(1, 8))
    elif kind == 3:
        x = x * 2.0 - 7.0
    elif kind == 4 and d > 1:
        x[:, 1:] = 0.5 + (x[:, 1:] - 0.5) * 1e-5
    x = np.ascontiguousarray(x)
    steps = int(rng.integers(50, 1500))
    if rng.integers(0, 3) == 0:
        steps = int(rng.integers(20, 400))
    t0 = float(rng.choice([1e-300, 1e-3, 0.05, 1.0]))
    minimize = bool(rng.integers(0, 2))
    chains = int(rng.choice([1, 1, 3, 150, 600, 3000]))  # more chains than SMs: the launch picks smaller blocks
    seed = int(rng.integers(0, 1 << 40))
    _lib.debug_set("MAXPRO_ANNEAL_MAXIMIN_SCREEN", "1")
    a = engine.anneal_lhd(x, steps, t0, 0.99, METRIC_MAXIMIN, minimize, seed, True, n_chains=chains)
    _lib.debug_set("MAXPRO_ANNEAL_MAXIMIN_SCREEN", "0")
    b = engine.anneal_lhd(x, steps, t0, 0.99, METRIC_MAXIMIN, minimize, seed, True, n_chains=chains)
    _lib.debug_set("MAXPRO_ANNEAL_MAXIMIN_SCREEN", None)
    runs += 1
    if chains > 1 and steps <= 400:  # the winning chain alone (one row per thread) must end where it ended among many
        c = engine.anneal_lhd(x, steps, t0, 0.99, METRIC_MAXIMIN, minimize, seed + a[2], True, n_chains=1)
        if not (np.array_equal(a[0], c[0]) and a[1] == c[1]):
            bad += 1
            print(f"REPLAY MISMATCH n={n} d={d} kind={kind} steps={steps} t0={t0} minimize={minimize} chains={chains} seed={seed}", flush=True)
    if not (np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2]):
        bad += 1
        print(f"MISMATCH n={n} d={d} kind={kind} steps={steps} t0={t0} minimize={minimize} chains={chains} seed={seed}", flush=True)
print(f"{runs} runs, {bad} mismatches")
sys.exit(1 if bad else 0)
