"""Maximin swap annealer: the screened step (MAXPRO_ANNEAL_MAXIMIN_SCREEN=1) against the exact row pass (=0).
Same chains through both kernels: (best metric, chain, design) must be identical; prints us per step and chain.
usage: python tools/anneal_maximin_ab.py [n d steps chains]..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from maxpro_b200 import engine, METRIC_MAXIMIN, _lib
import oracle

args = [int(a) for a in sys.argv[1:]] or [1000, 20, 1000, 148]
shapes = [args[i:i + 4] for i in range(0, len(args), 4)]
sms = 148
for n, d, steps, chains in shapes:
    x = oracle.generate_lhd(n, d, 42)
    res = {}
    for tag, sw in (("screened", "1"), ("exact", "0")):
        _lib.debug_set("MAXPRO_ANNEAL_MAXIMIN_SCREEN", sw)
        engine.anneal_lhd(x, 10, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=chains)
        best_t = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            out = engine.anneal_lhd(x, steps, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=chains)
            best_t = min(best_t, time.perf_counter() - t0)
        res[tag] = (out, best_t)
        waves = -(-chains // sms)
        print(f"n={n} d={d} steps={steps} chains={chains} {tag:9s}: {best_t*1e3:8.2f} ms  {best_t/steps/waves*1e6:6.2f} us/step/chain  "
              f"{chains*steps/best_t:.3e} chain-steps/s  best={out[1]!r} chain={out[2]}", flush=True)
    _lib.debug_set("MAXPRO_ANNEAL_MAXIMIN_SCREEN", None)
    a, b = res["screened"][0], res["exact"][0]
    same = np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2]
    print(f"  identical: {same}   speed-up {res['exact'][1]/res['screened'][1]:.2f}x", flush=True)
    assert same
