"""One maximin swap-annealing launch with the exact row pass (ncu target): python tools/anneal_maximin_prof_exact.py n d steps chains"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import engine, METRIC_MAXIMIN, _lib
import oracle
n, d, steps, chains = (int(a) for a in sys.argv[1:5])
_lib.debug_set("MAXPRO_ANNEAL_MAXIMIN_SCREEN", "0")
x = oracle.generate_lhd(n, d, 42)
print(engine.anneal_lhd(x, steps, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=chains)[1:])
