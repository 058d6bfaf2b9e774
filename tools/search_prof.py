"""One search launch at a given shape (ncu target): python tools/search_prof.py n d candidates"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import engine, METRIC_MAXPRO
n, d, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
engine.search_range(n, d, METRIC_MAXPRO, 1, 0, 2000)
print(engine.search_range(n, d, METRIC_MAXPRO, 42, 0, N))
