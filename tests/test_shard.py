"""CPU, world_size 2 over gloo: the N>1 paths -- the search (contiguous global shards, one 16-byte
all-gather, fold with the reference tie rule) and the annealing chains (contiguous chain blocks,
one all-gather, one broadcast of the winning design).  The per-rank scorer here is the CPU
oracle (tests may use it); on GPUs the same code runs with engine.search_range over NCCL."""
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from maxpro_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        res = {}
        for metric in (oracle.MAXPRO, oracle.MAXIMIN):
            res[metric] = shard.build_lhd_sharded(10, 3, 501, metric, 42, search_fn=oracle.search_range)
        # an all-tie case: n == 1 scores 0.0 / +inf everywhere -> the LAST global index must win
        tie = shard.allgather_minloc(0.0, 100 + rank, oracle.MAXPRO)
        # an empty shard contributes the identity
        lo, hi = shard.shard_range(1, rank, world)
        s, i = oracle.search_range(10, 3, oracle.MAXPRO, 42, lo, hi)
        res["tiny"] = shard.allgather_minloc(s, i, oracle.MAXPRO)
        res["tie"] = tie
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def _anneal_worker(rank, world, port, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        x = oracle.generate_lhd(12, 2, 3)

        def chains_on_cpu(design, iters, t0, cooling, metric, minimize, seed, swap, n_chains):
            # what engine.anneal_lhd does on a GPU: chain c uses stream seed + c, best chain wins, lowest index on ties
            runs = [oracle.anneal_lhd(design, iters, t0, cooling, metric, minimize, seed + c, swap) for c in range(n_chains)]
            best = 0
            for c in range(1, n_chains):
                if (runs[c][1] < runs[best][1]) if minimize else (runs[c][1] > runs[best][1]):
                    best = c
            return runs[best][0], runs[best][1], best

        res = {}
        for chains in (5, 1):  # 1 chain on 2 ranks: one rank contributes the identity
            res[chains] = shard.anneal_lhd_sharded(x, 150, 1.0, 0.98, oracle.MAXPRO, True, 77, True, chains,
                                                   anneal_fn=chains_on_cpu)
        res["ref5"] = chains_on_cpu(x, 150, 1.0, 0.98, oracle.MAXPRO, True, 77, True, 5)
        res["ref1"] = chains_on_cpu(x, 150, 1.0, 0.98, oracle.MAXPRO, True, 77, True, 1)
        q.put((rank, {k: (np.asarray(v[0]).tolist(), float(v[1]), int(v[2])) for k, v in res.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_anneal_world2_gloo():
    """Chains split over two ranks give the single-process answer: same best chain, same design."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_anneal_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for r in range(world):
        assert got[r][5] == got[r]["ref5"]
        assert got[r][1] == got[r]["ref1"]
    assert got[0][5] == got[1][5]


def test_shard_range_partitions():
    for total in (0, 1, 7, 1000, 2 ** 40 + 3):
        for world in (1, 2, 3, 8):
            blocks = [shard.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))


@pytest.mark.timeout(120)
def test_sharded_search_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for metric in (oracle.MAXPRO, oracle.MAXIMIN):
        single = oracle.search_range(10, 3, metric, 42, 0, 501)
        assert got[0][metric] == got[1][metric] == single  # same winner at any world size
    assert got[0]["tie"] == got[1]["tie"] == (0.0, 101)
    assert got[0]["tiny"] == got[1]["tiny"] == oracle.search_range(10, 3, oracle.MAXPRO, 42, 0, 1)
