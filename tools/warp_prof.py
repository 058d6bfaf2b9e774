import os, sys
sys.path.insert(0, "/root/repo")
from maxpro_b200 import METRIC_MAXPRO, engine
engine.search_range(100, 5, METRIC_MAXPRO, 42, 0, 20000)
print(engine.search_range(100, 5, METRIC_MAXPRO, 42, 0, 2000000))
