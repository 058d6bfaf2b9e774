"""Many maximin chains with jitter moves (full re-score per step): threads per chain against chain-steps/s.
usage: python tools/anneal_many_chains_jit.py n d steps chains [threads ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import engine, METRIC_MAXIMIN, _lib
import oracle
n, d, steps, chains = (int(a) for a in sys.argv[1:5])
threads = [int(a) for a in sys.argv[5:]] or [0, 64, 128, 256]
x = oracle.generate_lhd(n, d, 42)
for th in threads:
    _lib.debug_set("MAXPRO_ANNEAL_THREADS", str(th) if th else None)
    engine.anneal_lhd(x, 10, 0.01, 0.99, METRIC_MAXIMIN, False, 43, False, n_chains=chains)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        out = engine.anneal_lhd(x, steps, 0.01, 0.99, METRIC_MAXIMIN, False, 43, False, n_chains=chains)
        best = min(best, time.perf_counter() - t0)
    print(f"n={n} d={d} jitter chains={chains} threads={th or 'default'}: {chains*steps/best:.3e} chain-steps/s  ({best*1e3:.1f} ms) best={out[1]!r}", flush=True)
