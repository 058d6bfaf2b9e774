"""One annealing launch for ncu / timing: C4 shape, one chain per SM.
usage: python tools/anneal_prof.py [steps] [n] [d] [chains]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXPRO, engine
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 20
chains = int(sys.argv[4]) if len(sys.argv) > 4 else 148
x = engine.generate_lhd(n, d, 42)
engine.anneal_lhd(x, 2, 1.0, 0.99, METRIC_MAXPRO, True, 43, True, chains)
ts = []
for st in (steps // 4, steps):
    t0 = time.perf_counter()
    engine.anneal_lhd(x, st, 1.0, 0.99, METRIC_MAXPRO, True, 43, True, chains)
    ts.append(time.perf_counter() - t0)
print(f"n={n} d={d} chains={chains}: {steps//4} steps {ts[0]*1e3:.2f} ms, {steps} steps {ts[1]*1e3:.2f} ms, "
      f"marginal {(ts[1]-ts[0])/(steps-steps//4)*1e6:.3f} us/step", flush=True)
