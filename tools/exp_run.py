"""Time the headline search with a variant library: python tools/exp_run.py <tag|-> [n d metric flags log2N]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import _lib
tag = sys.argv[1] if len(sys.argv) > 1 else "-"
if tag != "-":
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), f"libmaxpro_b200_{tag}.so")
import torch
n, d, metric, flags, lg = (int(x) for x in (sys.argv[2:7] + ["50", "3", "0", "0", "27"][len(sys.argv) - 2:]))
L = _lib.lib(); dev = torch.device("cuda", 0)
ws_bytes = C.c_size_t(0); _lib.check(L.maxpro_search_workspace_bytes(n, d, C.byref(ws_bytes)))
ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev); res = torch.zeros(2, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream(); total = 1 << lg; best = 1e30
for r in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    _lib.check(L.maxpro_search_range_device(n, d, metric, 42, r * total, (r + 1) * total, flags, res.data_ptr(), ws.data_ptr(), ws_bytes.value, st.cuda_stream))
    e1.record(st); torch.cuda.synchronize()
    if r > 0: best = min(best, e0.elapsed_time(e1))
print(f"{tag:>8} n={n} d={d} metric={metric} flags={flags} N=2^{lg}: {best:.2f} ms = {total/best/1e6:.4f} Gcand/s", flush=True)
