"""Phase clocks of the maximin swap annealer's step (library built with -DAMX_PROFILE=1:
tools/exp_build.sh amx anneal_maximin.cu -DAMX_PROFILE=1).  usage: anneal_maximin_phases.py [n d steps [exact]]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libmaxpro_b200_amx.so")
from maxpro_b200 import engine, METRIC_MAXIMIN
import oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
_lib.debug_set("MAXPRO_ANNEAL_MAXIMIN_SCREEN", "0" if len(sys.argv) > 4 else "1")
x = oracle.generate_lhd(n, d, 42)
print(engine.anneal_lhd(x, steps, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=148)[1:])
