"""Multi-GPU sharding of the random search.

Candidates are independent (src/lib.rs:65-75: candidate i depends on seed+i only),
so rank r of W scores the contiguous GLOBAL index block [r*N/W, (r+1)*N/W) with no
data-path collective.  The one exchange is the final min-loc: every rank
contributes its 16-byte (score, index) record, one all-gather moves W*16 bytes
(NCCL over NVLink on GPUs, gloo in the CPU tests) and each rank folds the records
with the reference's rule (strictly better score wins, equal score keeps the
later index -- src/lib.rs:85-97).  NCCL has no MINLOC reduction op, hence
all-gather + local fold rather than all-reduce.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np

from . import engine


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of global candidate indices owned by `rank`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return total * rank // world, total * (rank + 1) // world


def fold_records(metric: int, scores, indices) -> Tuple[float, int]:
    """min-loc / max-loc with the later-index tie rule over per-shard winners."""
    return engine.reduce_best(metric, scores, indices)


def allgather_minloc(score: float, index: int, metric: int, device=None, group=None) -> Tuple[float, int]:
    """All ranks get the global winner.  16 bytes per rank on the wire."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rec = torch.tensor([np.float64(score).view(np.int64).item(), np.uint64(index).view(np.int64).item()],
                       dtype=torch.int64, device=device)
    out = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, rec, group=group)
    flat = out.cpu().numpy().reshape(world, 2)
    scores = flat[:, 0].copy().view(np.float64)
    indices = flat[:, 1].copy().view(np.uint64)
    return fold_records(metric, scores, indices)


def build_lhd_sharded(n: int, d: int, n_iterations: int, metric: int, seed: int,
                      search_fn: Callable = None, device=None, group=None):
    """build_lhd (src/lib.rs:30-100) over all ranks of the default process group.

    Returns (score, global index) on every rank; regenerate the design with
    engine.generate_lhd(n, d, seed + index).  The result is identical for any
    world size because indices are global.
    """
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n_iterations, rank, world)
    fn = search_fn or engine.search_range
    score, index = fn(n, d, metric, seed, lo, hi)
    return allgather_minloc(score, index, metric, device=device, group=group)
