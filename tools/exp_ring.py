import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import _lib
_lib.debug_set("MAXPRO_CYCLIC_RING", sys.argv[1])
sys.argv = ["x", "-"] + sys.argv[2:]
exec(open(os.path.join(os.path.dirname(__file__), "exp_run.py")).read())
