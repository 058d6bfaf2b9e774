"""Maximin swap annealer, screened step: threads per chain (MAXPRO_ANNEAL_THREADS) against us per step.
usage: python tools/anneal_maximin_threads.py n d steps [threads ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import engine, METRIC_MAXIMIN, _lib
import oracle
n, d, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
threads = [int(a) for a in sys.argv[4:]] or [128, 256, 384, 512]
x = oracle.generate_lhd(n, d, 42)
for th in threads:
    _lib.debug_set("MAXPRO_ANNEAL_THREADS", str(th))
    engine.anneal_lhd(x, 10, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=148)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        out = engine.anneal_lhd(x, steps, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=148)
        best = min(best, time.perf_counter() - t0)
    print(f"n={n} d={d} threads={th}: {best/steps*1e6:6.2f} us/step  best={out[1]!r}", flush=True)
