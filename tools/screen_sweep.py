"""Two-stage (default) vs fp64-only search rate on the shapes the general screening kernel serves."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, engine
shapes = [(10, 2), (20, 3), (30, 5), (40, 4), (50, 4), (50, 5), (64, 2), (64, 3), (50, 8), (64, 6), (32, 7)]
if len(sys.argv) > 1 and sys.argv[1] == 'big':
    shapes = [(64, 4), (80, 2), (80, 5), (100, 2), (100, 3), (100, 4), (100, 5), (100, 8), (128, 3), (128, 6), (150, 3), (200, 2), (200, 4), (256, 1), (256, 3)]
if len(sys.argv) > 1 and sys.argv[1] == 'g4':
    shapes = [(100, 8), (100, 10), (128, 8), (150, 10), (200, 6), (200, 8), (256, 4), (256, 6), (90, 9), (64, 10), (170, 10)]
if len(sys.argv) > 1 and sys.argv[1] == 'odd':
    shapes = [(9, 2), (25, 3), (51, 3), (63, 4), (101, 2), (151, 3), (255, 2), (99, 5)]
if len(sys.argv) > 1 and sys.argv[1] == 'mp':
    shapes = [(80, 3), (100, 3), (128, 2), (128, 3), (150, 2), (150, 3), (200, 2), (200, 3), (256, 2), (256, 3)]
if len(sys.argv) > 1 and sys.argv[1] == 'grid':
    shapes = [(n, d) for d in (1, 2, 3, 4) for n in (8, 12, 16, 24, 32, 48, 64)]
for metric, name in ((METRIC_MAXPRO, "maxpro"), (METRIC_MAXIMIN, "maximin")):
    for n, d in shapes:
        flops = ((3 * d + 3) if metric == METRIC_MAXPRO else (3 * d + 2)) * n * (n - 1) // 2
        N = max(1 << 16, min(1 << 26, int(2e12 / flops)))
        out = []
        for flags in (0, 2):
            engine.search_range(n, d, metric, 42, 0, 1 << 16, flags)
            t0 = time.perf_counter(); r = engine.search_range(n, d, metric, 42, 0, N, flags); dt = time.perf_counter() - t0
            out.append((N / dt, N / dt * flops / 1e12, r))
        same = out[0][2][1] == out[1][2][1] and abs(out[0][2][0] - out[1][2][0]) <= 1e-12 * abs(out[1][2][0])  # CTA kernel: 1-ulp run-to-run variation
        print(f"{name:7s} n={n:3d} d={d}: default {out[0][0]:10.4g} cand/s {out[0][1]:6.2f} TF | fp64-only {out[1][0]:10.4g} cand/s {out[1][1]:6.2f} TF | x{out[0][0]/out[1][0]:.2f} same winner {same}", flush=True)
