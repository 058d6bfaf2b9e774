"""Shapes at the hand-over points between the search kernels: every one must launch (and agree with its neighbours' path
on a re-score of the winner).  usage: python tools/boundary_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, engine
bad = 0
def probe(n, d):
    global bad
    for metric in (METRIC_MAXPRO, METRIC_MAXIMIN):
        for flags, N in ((2, 300), (0, 40000 if n * d <= 1000 else 300)):
            try:
                s, i = engine.search_range(n, d, metric, 5, 0, N, flags)
                x = engine.generate_lhd(n, d, 5 + i)
                c = engine.maxpro_criterion(x) if metric == METRIC_MAXPRO else engine.maximin_criterion(x)
                if not math.isclose(c, s, rel_tol=1e-9):
                    bad += 1; print("MISMATCH", n, d, metric, flags, s, c, flush=True)
            except Exception as e:
                bad += 1; print("ERROR", n, d, metric, flags, str(e)[:160], flush=True)
for d in (1, 2, 3, 4, 5, 8):
    # thread-group kernel limit (n d 8 B x 64 candidates ~ 227 KB), general screening kernel limits (n d = 454, 908)
    for target in (448, 454, 908, 1500, 2900):
        n0 = target // d
        for n in range(max(2, n0 - 2), n0 + 3):
            probe(n, d)
for d in (5, 10, 20):
    # CTA kernel: double-buffered up to n d 18 B ~ 111 KB, single-buffered up to n d 10 B ~ 227 KB, chunked beyond
    for target in (111 * 1024 // 18, 227 * 1024 // 10):
        n0 = target // d
        for n in range(n0 - 2, n0 + 3):
            probe(n, d)
# designs whose size divides the 227 KB budget exactly (a design per CTA that fills it to the byte): 232,448 = 2^10 * 227
for n, d in ((151, 1), (151, 2), (151, 3), (227, 1), (227, 2), (227, 4), (227, 8), (113, 2), (113, 4)):
    probe(n, d)
print("boundary sweep done, failures:", bad)
