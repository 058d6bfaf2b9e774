"""BASELINE config C3 as written: MaxPro random search n=500, d=10, 1e7 candidates sharded over the GPUs of
one box (index blocks, no data-path collective) with one 16-byte-per-rank NCCL all-gather for the min-loc.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c3_multi.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from maxpro_b200 import METRIC_MAXPRO, _lib, engine, shard

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
_lib.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
n, d, N, seed = 500, 10, 10_000_000, 42
shard.build_lhd_sharded(n, d, 20000 * world, METRIC_MAXPRO, seed, device=dev)  # warm-up (module load, NCCL channels)
times = []
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    score, index = shard.build_lhd_sharded(n, d, N, METRIC_MAXPRO, seed, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    times.append(time.perf_counter() - t0)
t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    flops = (3 * d + 3) * n * (n - 1) // 2 + 2
    design = engine.generate_lhd(n, d, seed + index)
    print(json.dumps({"config": "C3-sharded", "n": n, "d": d, "N": N, "n_gpus": world, "wall_s": t.item(), "candidates_per_s": N / t.item(),
                      "algorithmic_tflops": N / t.item() * flops / 1e12, "score": score, "index": index,
                      "winner_rescored": engine.maxpro_criterion(design)}), flush=True)
dist.destroy_process_group()
