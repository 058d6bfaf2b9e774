"""A/B of the cluster ordering kernels: order_cluster_async.cu (mbarriers, st.async, remote loads) against
order_cluster.cu (two cluster barriers per step)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, _lib, engine

shapes = [(5000, 8), (500, 5), (1000, 20), (10000, 2), (2000, 4), (3000, 10), (20000, 3)]
variants = [("sync", {"MAXPRO_ORDER_CLUSTER_SYNC": "1"})]
variants.append(("async8", {"MAXPRO_ORDER_CLUSTER_CTAS": "8"}))
variants.append(("async16", {"MAXPRO_ORDER_CLUSTER_CTAS": "16"}))
variants.append(("async", {}))
for n, d in shapes:
    x = engine.generate_lhd(n, d, 42)
    for metric, mn, name in ((METRIC_MAXPRO, True, "maxpro"), (METRIC_MAXIMIN, False, "maximin")):
        ref = None
        out = []
        for label, env in variants:
            for k, v in env.items(): _lib.debug_set(k, v)
            engine.order_design(x, metric, mn)
            best = 1e9
            for _ in range(4):
                t0 = time.perf_counter(); _, p = engine.order_design(x, metric, mn); best = min(best, time.perf_counter() - t0)
            for k in env: _lib.debug_set(k, None)
            if ref is None: ref = p.tolist()
            out.append(f"{label}: {best/n*1e6:.2f}{'' if p.tolist() == ref else ' MISMATCH'}")
        print(f"n={n} d={d} {name} us/step | " + " | ".join(out), flush=True)
