"""Random shapes, seeds and ranges through the two-stage search and through the fp64-only search: the (score, index)
pairs must be identical.  usage: python tools/screen_stress.py [seconds]"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, engine
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(20241020)
t0 = time.time(); runs = 0; screened = 0; cands = 0; bad = 0
while time.time() - t0 < budget:
    d = rng.randint(1, 10)
    n = rng.randint(8, min(256, 1500 // d))
    metric = rng.choice((METRIC_MAXPRO, METRIC_MAXIMIN))
    seed = rng.getrandbits(64); lo = rng.getrandbits(rng.randint(1, 60))
    count = rng.randint(1 << 15, 1 << rng.randint(16, 21))
    try:
        a = engine.search_range(n, d, metric, seed, lo, lo + count, 1)   # two-stage wherever the shape has a screening kernel
        b = engine.search_range(n, d, metric, seed, lo, lo + count, 2)   # fp64 only
    except Exception as e:
        bad += 1
        print("ERROR", n, d, metric, seed, lo, count, str(e)[-120:], flush=True)
        continue
    runs += 1; cands += count
    # identical bits wherever the fp64 kernel of the shape is deterministic (all but the dynamically scheduled CTA kernel,
    # whose MaxPro sums vary in the last bit from run to run)
    if a[1] != b[1] or abs(a[0] - b[0]) > 1e-13 * abs(b[0]) or (a[0] != b[0] and n * d <= 908):
        bad += 1
        print("MISMATCH", n, d, metric, seed, lo, count, a, b, flush=True)
print(f"{runs} random (shape, metric, seed, range) runs, {cands:.3g} candidates each way, {bad} mismatches in {time.time() - t0:.0f} s")
