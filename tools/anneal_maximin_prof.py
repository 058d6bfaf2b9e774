"""One incremental maximin swap-annealing launch at the C4 shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import engine, METRIC_MAXIMIN
import oracle
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
x = oracle.generate_lhd(1000, 20, 42)
print(engine.anneal_lhd(x, steps, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=148)[1:])
