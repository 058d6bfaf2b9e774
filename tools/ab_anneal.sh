set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "anneal" > $O/ab_anneal_tests.log 2>&1; tail -25 $O/ab_anneal_tests.log
python - <<'PY'
import os, sys, time
sys.path.insert(0, '.')
import numpy as np
from maxpro_b200 import engine, METRIC_MAXIMIN
import oracle
for (n, d, chains, steps) in [(1000, 20, 148, 2000), (1000, 20, 1184, 1000), (200, 4, 148, 4000), (50, 3, 148, 20000)]:
    x = oracle.generate_lhd(n, d, 42)
    for inc in ("0", "1"):
        if inc == "0": os.environ["MAXPRO_NO_ANNEAL_MAXIMIN_INC"] = "1"
        else: os.environ.pop("MAXPRO_NO_ANNEAL_MAXIMIN_INC", None)
        st = steps if inc == "1" else max(20, steps // 50)
        engine.anneal_lhd(x, 10, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=chains)
        t0 = time.perf_counter()
        got, best, chain = engine.anneal_lhd(x, st, 0.01, 0.99, METRIC_MAXIMIN, False, 43, True, n_chains=chains)
        dt = time.perf_counter() - t0
        print(f"n={n} d={d} chains={chains} steps={st} incremental={inc}: {dt*1e3:.1f} ms, {chains*st/dt:.4g} chain-steps/s, {dt/st/((chains+147)//148)*1e6:.2f} us/step/chain-slot, best {best}", flush=True)
PY
