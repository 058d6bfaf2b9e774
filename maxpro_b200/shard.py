"""Multi-GPU sharding of the random search and of the annealing chains.

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


def anneal_lhd_sharded(design, n_iterations: int, initial_temp: float, cooling_rate: float, metric: int,
                       minimize: bool, seed: int, swap: bool, n_chains: int,
                       anneal_fn: Callable = None, device=None, group=None):
    """n_chains independent chains of anneal_lhd (src/anneal.rs:56-143) over all ranks.

    Chain c draws from stream seed + c whatever the world size (rank r runs the contiguous block
    [r*C/W, (r+1)*C/W) by passing seed + lo as its seed), so the result is identical for any
    world size.  One all-gather of (metric, global chain) per rank -- strictly better metric wins,
    the lowest chain index on ties, as on one GPU -- and one broadcast of the winning design
    (n*d doubles) from the rank that owns it.  Returns (design, metric, chain) on every rank.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n_chains, rank, world)
    fn = anneal_fn or engine.anneal_lhd
    x = np.ascontiguousarray(design, dtype=np.float64)
    if hi > lo:
        out, best, chain = fn(x, n_iterations, initial_temp, cooling_rate, metric, minimize,
                              (seed + lo) & 0xFFFFFFFFFFFFFFFF, swap, hi - lo)
        chain += lo
    else:  # more ranks than chains: this rank contributes the identity
        out, best, chain = x.copy(), (np.inf if minimize else -np.inf), 2 ** 63 - 1
    rec = torch.tensor([np.float64(best).view(np.int64).item(), int(chain)], dtype=torch.int64, device=device)
    allrec = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allrec, rec, group=group)
    flat = allrec.cpu().numpy().reshape(world, 2)
    metrics = flat[:, 0].copy().view(np.float64)
    chains = flat[:, 1]
    owner = 0
    for r in range(1, world):
        better = metrics[r] < metrics[owner] if minimize else metrics[r] > metrics[owner]
        if better or (metrics[r] == metrics[owner] and chains[r] < chains[owner]):
            owner = r
    t = torch.from_numpy(np.ascontiguousarray(out)).to(device) if device is not None else torch.from_numpy(np.ascontiguousarray(out))
    dist.broadcast(t, src=owner if group is None else dist.get_global_rank(group, owner), group=group)
    return t.cpu().numpy(), float(metrics[owner]), int(chains[owner])
