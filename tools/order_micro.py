import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
x = engine.generate_lhd(n, d, 42)
for metric, mn in ((METRIC_MAXPRO, True), (METRIC_MAXIMIN, False)):
    engine.order_design(x, metric, mn)
    t0 = time.perf_counter(); engine.order_design(x, metric, mn); t = time.perf_counter() - t0
    print(f"order n={n} d={d} metric={metric}: {t*1e3:.2f} ms ({t/n*1e6:.2f} us/step)")
