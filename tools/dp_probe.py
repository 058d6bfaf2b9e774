"""FP64 pipe characterisation on the GPU box (tuning aid)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxpro_b200 import _lib
L = _lib.lib()
L.maxpro_debug_dp_probe.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_double)]
iters = 1 << 15
GHZ = 1.965
names = {1: "fma only", 5: "fma + 1 IMAD", 6: "fma + IMAD + ~3 ALU", 7: "fma + 2 IMAD + ~3 ALU"}
for op in (1, 5, 6, 7):
    for warps_per_sm in (4, 8, 16):
        ilp = 8
        threads = 32 * warps_per_sm
        ms = C.c_double(0)
        L.maxpro_debug_dp_probe(148, threads, ilp, op, iters, C.byref(ms))
        cyc = ms.value * 1e-3 * GHZ * 1e9 / (iters * ilp)
        per_smsp = cyc / (warps_per_sm / 4)
        print(f"{names[op]:24s} warps/SM={warps_per_sm:2d} ilp={ilp}: {ms.value:8.3f} ms  cycles per (DFMA + extras) per SMSP = {per_smsp:5.2f}", flush=True)
