"""Stress of the mbarrier ordering kernel: random shapes, both cluster sizes, every result against the kernel with
cluster barriers (order_cluster.cu).  usage: order_stress.py [seconds]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, _lib, engine

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 45.0
rng = np.random.default_rng(7)
t_end = time.time() + budget
runs = bad = 0
while time.time() < t_end:
    n = int(rng.integers(64, 6000)); d = int(rng.integers(1, 24))
    x = engine.generate_lhd(n, d, int(rng.integers(0, 1 << 30)))
    if rng.random() < 0.3:  # a grid: exact ties everywhere
        x = np.round(x * 8) / 8
    metric, mini = ((METRIC_MAXPRO, True), (METRIC_MAXIMIN, False), (METRIC_MAXIMIN, True), (METRIC_MAXPRO, False))[int(rng.integers(0, 4))]
    _lib.debug_set("MAXPRO_ORDER_CLUSTER_SYNC", "1")
    _, ref = engine.order_design(x, metric, mini)
    _lib.debug_set("MAXPRO_ORDER_CLUSTER_SYNC", None)
    for ctas in ("8", "16"):
        _lib.debug_set("MAXPRO_ORDER_CLUSTER_CTAS", ctas)
        for _ in range(3):
            _, p = engine.order_design(x, metric, mini)
            runs += 1
            if p.tolist() != ref.tolist():
                bad += 1
                print(f"MISMATCH n={n} d={d} metric={metric} minimize={mini} ctas={ctas}", flush=True)
    _lib.debug_set("MAXPRO_ORDER_CLUSTER_CTAS", None)
print(f"order stress: {runs} runs, {bad} mismatches")
