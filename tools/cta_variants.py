"""CTA-per-candidate search: double-buffered kernel against the single-buffer kernel at its block sizes and layouts.
usage: python tools/cta_variants.py [n d]..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXPRO, engine, _lib
args = [int(a) for a in sys.argv[1:]] or [500, 10, 300, 6, 200, 20, 1000, 5, 120, 16, 350, 10]
engine.search_range(50, 3, METRIC_MAXPRO, 1, 0, 1000)
_lib.debug_set("MAXPRO_NO_WARP", "1")
for n, d in zip(args[0::2], args[1::2]):
    flops = (3 * d + 3) * n * (n - 1) // 2
    N = max(2000, min(1 << 26, int(2e12 / flops)))
    out = []
    variants = [("pipe", {"MAXPRO_CTA_PIPE": "1", "MAXPRO_CTA_PIPE_COMPACT": "1"}), ("pipe staged", {"MAXPRO_CTA_PIPE": "1", "MAXPRO_CTA_PIPE_COMPACT": "0"})]
    for th in ("128", "256", "512"):
        for cp in ("0", "1"):
            variants.append((f"single t{th} c{cp}", {"MAXPRO_CTA_PIPE": "0", "MAXPRO_CTA_THREADS": th, "MAXPRO_CTA_COMPACT": cp}))
    for tag, sw in variants:
        for k in ("MAXPRO_CTA_PIPE", "MAXPRO_CTA_PIPE_COMPACT", "MAXPRO_CTA_THREADS", "MAXPRO_CTA_COMPACT"):
            _lib.debug_set(k, sw.get(k))
        try:
            engine.search_range(n, d, METRIC_MAXPRO, 42, 0, min(N, 1 << 14))
            t0 = time.perf_counter()
            s, i = engine.search_range(n, d, METRIC_MAXPRO, 42, 0, N)
            dt = time.perf_counter() - t0
            out.append(f"{tag} {N / dt * flops / 1e12:5.2f}")
        except Exception as e:
            out.append(f"{tag} n/a")
    print(f"n={n} d={d}: " + " | ".join(out), flush=True)
