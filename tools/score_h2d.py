"""maxpro_score_batch from pageable host memory: the staged multi-threaded copy (default for large inputs) against the plain one.
usage: python tools/score_h2d.py [designs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from maxpro_b200 import engine, METRIC_MAXPRO, _lib
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
designs = np.random.default_rng(7).random((nb, 50, 3))
engine.score_batch(designs[:1024], METRIC_MAXPRO)
for tag, sw in (("staged", None), ("plain", "0"), ("staged", None)):
    _lib.debug_set("MAXPRO_H2D_STAGED", sw)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter(); sc, arg = engine.score_batch(designs, METRIC_MAXPRO); best = min(best, time.perf_counter() - t0)
    print(f"{tag:7s}: {nb/best:.3e} designs/s  {designs.nbytes/best/1e9:6.2f} GB/s host->device  ({best*1e3:.1f} ms, argbest {arg})", flush=True)
print("host threads:", os.cpu_count())
