import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, engine
def timed(fn, *a):
    t0 = time.perf_counter(); out = fn(*a); return out, time.perf_counter() - t0
engine.search_range(50, 3, METRIC_MAXPRO, 1, 0, 1000)
for (n, d) in ((1000, 20), (50, 3), (200, 5)):
    x = engine.generate_lhd(n, d, 42)
    for chains, steps in ((148, 1), (148, 50), (148, 400), (1184, 50)):
        for swap in (True, False) if (n * d * 16 < 220000) else (True,):
            _, t = timed(engine.anneal_lhd, x, steps, 1.0, 0.99, METRIC_MAXPRO, True, 43, swap, chains)
            print(f"n={n} d={d} swap={swap} chains={chains} steps={steps}: {t*1e3:.2f} ms  -> {t/steps/ (chains/148)*1e6:.1f} us per step per SM-slot", flush=True)
