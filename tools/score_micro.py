import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, _lib, engine
L = _lib.lib(); dev = torch.device("cuda", 0)
"""K2 on supplied designs (maxpro_score_batch_device, designs resident in HBM): rates per shape.
usage: python tools/score_micro.py [n d]   (one shape: the ncu target)"""
SHAPES = ((50, 3, 1 << 20), (50, 2, 1 << 20), (100, 5, 1 << 18), (500, 10, 1 << 13), (1000, 20, 1 << 11), (5000, 8, 64))
if len(sys.argv) == 3:
    SHAPES = tuple(s for s in SHAPES if s[:2] == (int(sys.argv[1]), int(sys.argv[2])))
for (n, d, b) in SHAPES:
    x = torch.rand(b, n, d, dtype=torch.float64, device=dev)
    scores = torch.empty(b, dtype=torch.float64, device=dev)
    res = torch.zeros(2, dtype=torch.int64, device=dev)
    wsb = C.c_size_t(0); _lib.check(L.maxpro_score_workspace_bytes(b, n, d, C.byref(wsb)))
    ws = torch.empty(wsb.value, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    for metric in (METRIC_MAXPRO, METRIC_MAXIMIN):
        best = 1e30
        for r in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            _lib.check(L.maxpro_score_batch_device(x.data_ptr(), b, n, d, metric, scores.data_ptr(), res.data_ptr(), ws.data_ptr(), wsb.value, st.cuda_stream))
            e1.record(st); torch.cuda.synchronize()
            if r: best = min(best, e0.elapsed_time(e1))
        pairs = n * (n - 1) / 2
        flops = ((3 * d + 3) if metric == METRIC_MAXPRO else (3 * d + 2)) * pairs
        print(f"score_batch n={n} d={d} b={b} metric={metric}: {best:.3f} ms  {b/best/1e3:.3f} Mdesigns/s  {b*flops/best/1e9:.2f} alg TFLOP/s  {b*n*d*8/best/1e6:.1f} GB/s read", flush=True)
