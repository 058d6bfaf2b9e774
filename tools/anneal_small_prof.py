"""One chain of jitter (perturb) moves on a small design -- the CLI's last stage (main.rs:88-96) -- for the phase
profile of anneal_kernel (make EXTRA=-DANN_PROFILE=1) and for timing: usage anneal_small_prof.py [n d steps]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import engine, METRIC_MAXPRO, METRIC_MAXIMIN
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
d = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
x = engine.generate_lhd(n, d, 42)
for metric, mini, name in ((METRIC_MAXPRO, True, "maxpro"), (METRIC_MAXIMIN, False, "maximin")):
    for swap in (False, True):
        engine.anneal_lhd(x, 100, 1.0, 0.99, metric, mini, 44, swap, n_chains=1)
        t0 = time.perf_counter(); engine.anneal_lhd(x, steps, 1.0, 0.99, metric, mini, 44, swap, n_chains=1); t = time.perf_counter() - t0
        print(f"n={n} d={d} {name} {'swap' if swap else 'jitter'}: {t / steps * 1e6:.2f} us per step", flush=True)
