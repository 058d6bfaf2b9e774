"""Mid-size designs: the warp-per-candidate kernel (default) against the CTA-per-candidate kernels (MAXPRO_NO_WARP=1),
fp64 search (MAXPRO_FLAG_FP64_ONLY semantics do not matter here: shapes outside the screening kernels).
usage: python tools/warp_vs_cta.py [maxpro|maximin]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, engine, _lib
metric = METRIC_MAXIMIN if len(sys.argv) > 1 and sys.argv[1] == "maximin" else METRIC_MAXPRO
engine.search_range(50, 3, metric, 1, 0, 1000)
shapes = [(200, 20), (300, 6), (260, 8), (400, 5), (300, 12), (160, 30), (350, 10), (120, 16), (600, 3), (257, 5)]
for n, d in shapes:
    flops = ((3 * d + 3) if metric == METRIC_MAXPRO else (3 * d + 2)) * n * (n - 1) // 2
    N = max(2000, min(1 << 26, int(2e12 / flops)))
    row = []
    for tag, sw in (("warp", None), ("cta", "1")):
        _lib.debug_set("MAXPRO_NO_WARP", sw)
        engine.search_range(n, d, metric, 42, 0, min(N, 1 << 16))
        t0 = time.perf_counter()
        s, i = engine.search_range(n, d, metric, 42, 0, N)
        dt = time.perf_counter() - t0
        row.append((tag, N / dt * flops / 1e12, i))
    _lib.debug_set("MAXPRO_NO_WARP", None)
    print(f"n={n:5d} d={d:3d}: " + "  ".join(f"{t} {v:6.2f} TFLOP/s" for t, v, _ in row) + f"   same winner: {row[0][2] == row[1][2]}", flush=True)
