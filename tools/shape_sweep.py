"""Search throughput over design shapes (which kernel serves which shape, and how well).
usage: python tools/shape_sweep.py [maxpro|maximin]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, engine
metric = METRIC_MAXIMIN if len(sys.argv) > 1 and sys.argv[1] == "maximin" else METRIC_MAXPRO
engine.search_range(50, 3, metric, 1, 0, 1000)
shapes = [(10, 2), (20, 3), (30, 5), (50, 2), (50, 3), (50, 4), (50, 8), (64, 6), (80, 5), (100, 2), (100, 5), (100, 10),
          (128, 8), (150, 3), (200, 4), (200, 20), (300, 6), (500, 10), (1000, 5), (1000, 20), (2000, 10), (5000, 5)]
for n, d in shapes:
    flops = ((3 * d + 3) if metric == METRIC_MAXPRO else (3 * d + 2)) * n * (n - 1) // 2
    N = max(2000, min(1 << 26, int(2e12 / flops)))  # ~0.1-0.2 s of work at 10-20 TFLOP/s
    engine.search_range(n, d, metric, 42, 0, min(N, 1 << 16))  # same path as the timed call (two-stage from 32,768 candidates up)
    t0 = time.perf_counter()
    s, i = engine.search_range(n, d, metric, 42, 0, N)
    dt = time.perf_counter() - t0
    print(f"n={n:5d} d={d:3d} N={N:9d}: {N/dt:12.4g} cand/s  {N/dt*flops/1e12:6.2f} algorithmic TFLOP/s  ({dt*1e3:.1f} ms)", flush=True)
