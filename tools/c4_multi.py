"""BASELINE config C4 sharded over the GPUs of one box: n=1000, d=20, 65,536 chains PER GPU, swap moves;
chain c draws from stream seed + c whatever the world size, one all-gather of (metric, chain) and one
broadcast of the winning design at the end.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c4_multi.py [steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, _lib, engine, shard

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
_lib.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n, d, per_gpu = 1000, 20, 65536
x = engine.generate_lhd(n, d, 42)
for met, mini, name, t0_ in ((METRIC_MAXPRO, True, "maxpro", 1.0), (METRIC_MAXIMIN, False, "maximin", 0.01)):
    shard.anneal_lhd_sharded(x, 2, t0_, 0.99, met, mini, 43, True, 148 * world, device=dev)  # warm-up
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, best, chain = shard.anneal_lhd_sharded(x, steps, t0_, 0.99, met, mini, 43, True, per_gpu * world, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the winning chain, re-run alone on this rank, must give the same design (chains are independent)
    solo, solo_best, _ = engine.anneal_lhd(x, steps, t0_, 0.99, met, mini, (43 + chain) & (2 ** 64 - 1), True, n_chains=1)
    same = bool(np.array_equal(solo, out)) and solo_best == best
    if rank == 0:
        print(json.dumps({"config": "C4-sharded-" + name, "n": n, "d": d, "n_gpus": world, "chains": per_gpu * world, "steps": steps,
                          "wall_s": t.item(), "chain_steps_per_s": per_gpu * world * steps / t.item(), "best_metric": best,
                          "best_chain": chain, "winner_reproduced_by_single_chain": same}), flush=True)
dist.destroy_process_group()
