import os, sys
sys.path.insert(0, '/root/repo')
from maxpro_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libmaxpro_b200_amx.so")
from maxpro_b200 import engine, METRIC_MAXIMIN
import oracle
x = oracle.generate_lhd(1000, 20, 42)
print(engine.anneal_lhd(x, 1000, 1.0, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=148)[1:])
