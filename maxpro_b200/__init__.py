"""maxpro_b200 -- B200-native engine for the data-parallel hot path of mrhheffernan/maxpro.

Layout:
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/maxpro_b200.h)
  _lib.py    ctypes binding of libmaxpro_b200.so
  engine.py  NumPy-level calls over the C ABI
  maxpro.py  drop-in for the reference's pyo3 module `maxpro`
  shard.py   multi-GPU sharding of the search + min-loc over torch.distributed
  host/      C++ mirror of the reference's Rust API and its CLI
"""
from ._lib import (ERR_CUDA, ERR_INVALID, ERR_NO_DEVICE, ERR_OOM, ERR_UNSUPPORTED, METRIC_MAXIMIN,
                   METRIC_MAXPRO, METRIC_NONE, MaxproError)

__all__ = ["MaxproError", "METRIC_MAXPRO", "METRIC_MAXIMIN", "METRIC_NONE", "ERR_INVALID", "ERR_CUDA",
           "ERR_NO_DEVICE", "ERR_OOM", "ERR_UNSUPPORTED"]
