"""BASELINE config C4 at its full chain count: n=1000, d=20, 65,536 chains on one GPU, swap moves."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import engine, METRIC_MAXPRO, METRIC_MAXIMIN
n, d, chains = 1000, 20, 65536
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
x = engine.generate_lhd(n, d, 42)
for met, mini, name, t0_, cool in ((METRIC_MAXPRO, True, "maxpro", 1.0, 0.99), (METRIC_MAXIMIN, False, "maximin", 0.01, 0.99)):
    engine.anneal_lhd(x, 2, t0_, cool, met, mini, 43, True, n_chains=148)
    t0 = time.perf_counter()
    out, best, chain = engine.anneal_lhd(x, steps, t0_, cool, met, mini, 43, True, n_chains=chains)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "C4-full-" + name, "n": n, "d": d, "chains": chains, "steps": steps, "wall_s": dt,
                      "chain_steps_per_s": chains * steps / dt, "best_metric": best, "best_chain": chain}), flush=True)
