import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXPRO, _lib, engine
x = engine.generate_lhd(1000, 20, 42)
engine.anneal_lhd(x, 10, 1.0, 0.99, METRIC_MAXPRO, True, 43, True, 148)
for thr in (512, 384, 256, 128, 64):
    _lib.debug_set("MAXPRO_ANNEAL_THREADS", str(thr))
    ts = []
    for steps in (100, 500):
        t0 = time.perf_counter(); engine.anneal_lhd(x, steps, 1.0, 0.99, METRIC_MAXPRO, True, 43, True, 148); ts.append(time.perf_counter() - t0)
    print(f"threads={thr}: {(ts[1]-ts[0])/400*1e6:.2f} us/step (marginal), 500 steps {ts[1]*1e3:.2f} ms", flush=True)
