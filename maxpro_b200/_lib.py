"""ctypes binding of libmaxpro_b200.so (include/maxpro_b200.h).

This is the stub a maintainer of the reference's Python package would write to
swap the pyo3 extension (src/lib.rs:166-176) for the C-ABI engine.  There is no
CPU fallback: if the shared library is missing the import fails loudly, and if
no B200 is present every compute call raises MaxproError(MAXPRO_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmaxpro_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_OOM, ERR_UNSUPPORTED = range(6)
METRIC_NONE, METRIC_MAXPRO, METRIC_MAXIMIN = -1, 0, 1

# every symbol include/maxpro_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "maxpro_last_error", "maxpro_version", "maxpro_device_count", "maxpro_set_device",
    "maxpro_get_device", "maxpro_kernel_launch_count", "maxpro_generate_lhd",
    "maxpro_generate_batch", "maxpro_criterion", "maxpro_maximin_criterion",
    "maxpro_l2_distance", "maxpro_score_batch", "maxpro_build_lhd", "maxpro_build_lhd_ex",
    "maxpro_search_range", "maxpro_reduce_best", "maxpro_anneal_lhd", "maxpro_order_design",
    "maxpro_search_workspace_bytes", "maxpro_search_range_device",
    "maxpro_score_workspace_bytes", "maxpro_score_batch_device", "maxpro_fp64_peak_probe", "maxpro_fp32_peak_probe",
    "maxpro_release_device_cache", "maxpro_set_device_cache_limit", "maxpro_set_devices", "maxpro_get_devices",
    "maxpro_anneal_scratch_bytes", "maxpro_debug_set", "maxpro_debug_pow", "maxpro_set_stream_version",
    "maxpro_get_stream_version", "maxpro_search_range_ex", "maxpro_debug_screen", "maxpro_debug_screen_threshold",
]


class MaxproError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[maxpro_b200 status {code}] {message}")
        self.code = code
        self.message = message


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C maxpro_b200/csrc`. maxpro_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    dp, u64, i32, dbl, u32 = C.POINTER(C.c_double), C.c_uint64, C.c_int, C.c_double, C.c_uint32
    u64p, vp, szp = C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_size_t)
    sig = {
        "maxpro_last_error": (C.c_char_p, []),
        "maxpro_version": (C.c_char_p, []),
        "maxpro_device_count": (i32, [C.POINTER(i32)]),
        "maxpro_set_device": (i32, [i32]),
        "maxpro_get_device": (i32, [C.POINTER(i32)]),
        "maxpro_kernel_launch_count": (u64, []),
        "maxpro_release_device_cache": (i32, [u64p]),
        "maxpro_set_device_cache_limit": (i32, [u64]),
        "maxpro_set_devices": (i32, [C.POINTER(i32), i32]),
        "maxpro_get_devices": (i32, [C.POINTER(i32), i32, C.POINTER(i32)]),
        "maxpro_anneal_scratch_bytes": (i32, [u64, u64, i32, i32, u64, szp]),
        "maxpro_set_stream_version": (i32, [i32]),
        "maxpro_get_stream_version": (i32, [C.POINTER(i32)]),
        "maxpro_search_range_ex": (i32, [u64, u64, i32, u64, u64, u64, u32, dp, u64p]),
        "maxpro_debug_screen": (i32, [u64, u64, u64, u64, u64, u64p, u64, u64p, dp]),
        "maxpro_debug_screen_threshold": (i32, [u64, u64, dbl, dp]),
        "maxpro_debug_set": (i32, [C.c_char_p, C.c_char_p]),
        "maxpro_debug_pow": (i32, [dp, dp, u64, dp]),
        "maxpro_generate_lhd": (i32, [u64, u64, u64, dp]),
        "maxpro_generate_batch": (i32, [u64, u64, u64, u64, u64, dp]),
        "maxpro_criterion": (i32, [dp, u64, u64, dp]),
        "maxpro_maximin_criterion": (i32, [dp, u64, u64, dp]),
        "maxpro_l2_distance": (i32, [dp, dp, u64, dp]),
        "maxpro_score_batch": (i32, [dp, u64, u64, u64, i32, dp, u64p]),
        "maxpro_build_lhd": (i32, [u64, u64, u64, i32, u64, dp, dp, u64p]),
        "maxpro_build_lhd_ex": (i32, [u64, u64, u64, i32, u64, u32, dp, dp, u64p]),
        "maxpro_search_range": (i32, [u64, u64, i32, u64, u64, u64, dp, u64p]),
        "maxpro_reduce_best": (i32, [i32, dp, u64p, u64, dp, u64p]),
        "maxpro_anneal_lhd": (i32, [dp, u64, u64, u64, dbl, dbl, i32, i32, u64, i32, u64, dp, dp, u64p]),
        "maxpro_order_design": (i32, [dp, u64, u64, i32, i32, dp, u64p]),
        "maxpro_search_workspace_bytes": (i32, [u64, u64, szp]),
        "maxpro_search_range_device": (i32, [u64, u64, i32, u64, u64, u64, u32, vp, vp, C.c_size_t, vp]),
        "maxpro_score_workspace_bytes": (i32, [u64, u64, u64, szp]),
        "maxpro_score_batch_device": (i32, [vp, u64, u64, u64, i32, vp, vp, vp, C.c_size_t, vp]),
        "maxpro_fp64_peak_probe": (i32, [i32, dp, dp]),
        "maxpro_fp32_peak_probe": (i32, [i32, dp, dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != OK:
        raise MaxproError(rc, lib().maxpro_last_error().decode("utf-8", "replace"))


def dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def u64ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def device_count() -> int:
    c = C.c_int(0)
    check(lib().maxpro_device_count(C.byref(c)))
    return c.value


def set_device(device: int) -> None:
    check(lib().maxpro_set_device(device))


def set_devices(devices=None) -> None:
    """Devices the build_lhd / search_range / anneal_lhd calls of this thread spread over (None or [] = all)."""
    devs = list(devices or [])
    arr = (C.c_int * max(len(devs), 1))(*devs)
    check(lib().maxpro_set_devices(arr, len(devs)))


def get_devices() -> list:
    n = C.c_int(0)
    arr = (C.c_int * 64)()
    check(lib().maxpro_get_devices(arr, 64, C.byref(n)))
    return [arr[i] for i in range(min(n.value, 64))]


def set_stream_version(version: int) -> None:
    """Design stream: 2 = LHD-mix v2 (default), 1 = LHD-Philox v1."""
    check(lib().maxpro_set_stream_version(version))


def get_stream_version() -> int:
    v = C.c_int(0)
    check(lib().maxpro_get_stream_version(C.byref(v)))
    return v.value


def debug_set(name: str, value) -> None:
    """Test / tooling switch; value None clears it."""
    check(lib().maxpro_debug_set(name.encode(), None if value is None else str(value).encode()))


def debug_screen(n: int, d: int, seed: int, lo: int, hi: int, want_sums: bool = False, capacity: int = 1 << 20):
    """Stage 1 of the two-stage search alone: (survivor indices, count stage 1 produced, fp32 sums or None)."""
    surv = np.zeros(capacity, dtype=np.uint64)
    cnt = C.c_uint64(0)
    sums = np.zeros(hi - lo, dtype=np.float64) if want_sums else None
    check(lib().maxpro_debug_screen(n, d, seed & (2 ** 64 - 1), lo, hi, u64ptr(surv), capacity, C.byref(cnt),
                                    dptr(sums) if want_sums else None))
    return surv[:min(cnt.value, capacity)], int(cnt.value), sums


def debug_screen_threshold(n: int, d: int, s: float) -> float:
    t = C.c_double(0)
    check(lib().maxpro_debug_screen_threshold(n, d, s, C.byref(t)))
    return t.value


def debug_pow(x, y) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    check(lib().maxpro_debug_pow(dptr(x), dptr(y), x.size, dptr(out)))
    return out


def anneal_scratch_bytes(n: int, d: int, metric: int, swap: bool, n_chains: int) -> int:
    b = C.c_size_t(0)
    check(lib().maxpro_anneal_scratch_bytes(n, d, metric, 1 if swap else 0, n_chains, C.byref(b)))
    return int(b.value)


def set_device_cache_limit(nbytes: int) -> None:
    check(lib().maxpro_set_device_cache_limit(nbytes))


def kernel_launch_count() -> int:
    return int(lib().maxpro_kernel_launch_count())


def release_device_cache() -> int:
    """Give the device blocks the library kept from finished calls of this thread back to the driver; returns the bytes."""
    b = C.c_uint64(0)
    check(lib().maxpro_release_device_cache(C.byref(b)))
    return int(b.value)
