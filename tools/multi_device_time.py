import sys, time; sys.path.insert(0, '/root/repo')
from maxpro_b200 import _lib, engine, METRIC_MAXPRO
for devs in ([0], [0, 1], [0], [0, 1]):
    _lib.set_devices(devs)
    t0 = time.perf_counter(); r = engine.build_lhd(500, 10, 4_000_000, METRIC_MAXPRO, 42); dt = time.perf_counter() - t0
    print(devs, f"{dt:.3f} s", r[1], r[2], flush=True)
    t0 = time.perf_counter(); r = engine.build_lhd(50, 3, 1 << 28, METRIC_MAXPRO, 42); dt = time.perf_counter() - t0
    print(devs, f"{dt:.3f} s", r[1], r[2], flush=True)
