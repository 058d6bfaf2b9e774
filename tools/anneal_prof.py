import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXPRO, engine
x = engine.generate_lhd(1000, 20, 42)
t0 = time.perf_counter()
engine.anneal_lhd(x, 200, 1.0, 0.99, METRIC_MAXPRO, True, 43, True, 148)
print((time.perf_counter() - t0) * 1e3, "ms")
