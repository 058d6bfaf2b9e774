"""Many maximin swap chains: screened against exact row pass, each at the block size the launch picks.
usage: python tools/anneal_many_chains_screen.py n d steps chains"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import engine, METRIC_MAXIMIN, _lib
import oracle
n, d, steps, chains = (int(a) for a in sys.argv[1:5])
x = oracle.generate_lhd(n, d, 42)
for tag, sw in (("screened", "1"), ("exact", "0")):
    _lib.debug_set("MAXPRO_ANNEAL_MAXIMIN_SCREEN", sw)
    engine.anneal_lhd(x, 10, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=chains)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        out = engine.anneal_lhd(x, steps, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=chains)
        best = min(best, time.perf_counter() - t0)
    print(f"n={n} d={d} chains={chains} {tag}: {chains*steps/best:.3e} chain-steps/s  best={out[1]!r}", flush=True)
