"""GPU micro-measurements used while tuning (not part of the product or the tests).
usage: python tools/gpu_micro.py [search|ncu]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, _lib, engine

L = _lib.lib()
dev = torch.device("cuda", 0)


def dev_search(n, d, metric, total, reps=3, flags=0):
    ws_bytes = C.c_size_t(0)
    _lib.check(L.maxpro_search_workspace_bytes(n, d, C.byref(ws_bytes)))
    ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
    res = torch.zeros(2, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()
    best = 1e30
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        _lib.check(L.maxpro_search_range_device(n, d, metric, 42, r * total, (r + 1) * total, flags, res.data_ptr(),
                                                ws.data_ptr(), ws_bytes.value, st.cuda_stream))
        e1.record(st)
        torch.cuda.synchronize()
        if r > 0:
            best = min(best, e0.elapsed_time(e1))
    return best


mode = sys.argv[1] if len(sys.argv) > 1 else "search"
if mode == "ncu":
    total = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 192 * 16
    print(engine.search_range(50, 3, METRIC_MAXPRO, 42, 0, total))
    sys.exit(0)

if mode == "screen":
    total = 1 << 26
    for metric in (METRIC_MAXPRO, METRIC_MAXIMIN):
        for (n, d) in ((50, 3), (50, 2)):
            ms0 = dev_search(n, d, metric, total, flags=2)   # MAXPRO_FLAG_FP64_ONLY
            ms1 = dev_search(n, d, metric, total, flags=1)   # two-stage
            print(f"n={n} d={d} metric={metric}: exact {total/ms0/1e6:.3f} Gcand/s | fp32 screen + fp64 rescore {total/ms1/1e6:.3f} Gcand/s", flush=True)
    sys.exit(0)

if mode == "sweep":
    total = 1 << 25
    _lib.debug_set("MAXPRO_NO_CYCLIC", "0")
    for ring in (0, 2, 3):
        _lib.debug_set("MAXPRO_CYCLIC_RING", str(ring))
        for metric in (METRIC_MAXPRO, METRIC_MAXIMIN):
            for (n, d) in ((50, 3), (50, 2)):
                ms = dev_search(n, d, metric, total)
                print(f"cyclic ring={ring} n={n} d={d} metric={metric}: {ms:.2f} ms = {total/ms/1e6:.3f} Gcand/s", flush=True)
    _lib.debug_set("MAXPRO_NO_CYCLIC", "1")
    for lock in (0,):
        for G, cands in ((1, 192), (2, 192)):
            _lib.debug_set("MAXPRO_SMALL_GROUP", str(G))
            _lib.debug_set("MAXPRO_SMALL_CANDS", str(cands))
            _lib.debug_set("MAXPRO_SMALL_LOCK", str(lock))
            ms = dev_search(50, 3, METRIC_MAXPRO, total)
            print(f"lock={lock} G={G} cands/CTA={cands} threads={G*cands}: {ms:.2f} ms = {total/ms/1e6:.3f} Gcand/s", flush=True)
    sys.exit(0)

print("fp64 probe TFLOP/s:", engine.fp64_peak_probe(5))
for total in (1 << 22, 1 << 24, 1 << 26):
    ms = dev_search(50, 3, METRIC_MAXPRO, total)
    t0 = time.perf_counter(); engine.search_range(50, 3, METRIC_MAXPRO, 42, 0, total); host = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); engine.build_lhd(50, 3, total, METRIC_MAXPRO, 42); host2 = (time.perf_counter() - t0) * 1e3
    print(f"maxpro n50 d3 N={total}: device {ms:.2f} ms = {total/ms/1e6:.1f} Gcand/s | host search_range {host:.2f} ms | build_lhd {host2:.2f} ms")
for (n, d, metric, total) in [(50, 3, METRIC_MAXIMIN, 1 << 24), (50, 2, METRIC_MAXPRO, 1 << 24), (50, 8, METRIC_MAXPRO, 1 << 22),
                              (25, 3, METRIC_MAXPRO, 1 << 24), (100, 3, METRIC_MAXPRO, 1 << 22)]:
    ms = dev_search(n, d, metric, total)
    print(f"n={n} d={d} metric={metric} N={total}: {ms:.2f} ms = {total/ms/1e6:.1f} Gcand/s")
