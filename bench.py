#!/usr/bin/env python
"""bench.py -- headline measurement of the maxpro hot path on B200.

Metric (BASELINE.json): MaxPro candidate LHDs scored per second at n=50, d=3, per node,
plus the fraction of the FP64 (non-tensor) pipe peak the criterion kernel reaches.

A "step" is one pass of the fused generate-and-score search (src/lib.rs:65-98) over
CANDIDATES_PER_GPU fresh candidates per GPU (weak scaling: every rank scores its own
contiguous block of GLOBAL candidate indices; no data-path collective).

  value  device-resident throughput of the engine's DEFAULT search path: maxpro_search_range_device
         on the launching stream, timed with CUDA events, max over ranks.  For this shape the
         default is the exact two-stage search -- every candidate generated and scored in fp32 with a
         proven error bound, every candidate that can still win re-scored in fp64 -- which returns
         the winner, score and tie-break of the fp64-only search (tests/test_gpu_parity.py).  The
         fp64-only kernel (MAXPRO_FLAG_FP64_ONLY) is timed the same way and reported as `fp64_only`
         with its own FP64 roofline.  The search has no HBM inputs at all -- candidates are
         generated in shared memory and only 16 bytes per CTA (plus 8 per survivor) leave the chip.
  e2e    the same metric through the reference-facing host call (maxpro_search_range /
         maxpro_build_lhd: scalars in, winner design out to host memory), including the
         16-byte-per-rank min-loc all-gather over NCCL when N > 1.
  roofline      algorithmic flops ((3d+3)*n(n-1)/2 + 2 per candidate) over the step
                time, against the non-tensor FMA peak of the pipe the dominant kernel runs on,
                measured in this run by an independent-FMA probe (FP32 for the screening kernel,
                FP64 for `fp64_only`).
  cpu_baseline  the C/OpenMP oracle port of the reference algorithm on this box's host cores,
                bounded sample (rank 0, N == 1 only).

`--impl reference` times that CPU port alone (the Rust reference cannot be built here: no
cargo/rustc in the image -- DESIGN.md), all host threads, same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, N_DIMS, SEED = 50, 3, 42
CANDIDATES_PER_GPU = 1 << 28
FLOPS_PER_CANDIDATE = (3 * N_DIMS + 3) * N_POINTS * (N_POINTS - 1) // 2 + 2  # 14,702
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12                          # 37.2, datasheet-level
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12                         # 74.4
FLAG_DEFAULT, FLAG_FP32_SCREEN, FLAG_FP64_ONLY = 0, 1, 2                     # include/maxpro_b200.h
METRIC_NAME = "maxpro_candidate_lhds_scored_per_sec_n50_d3"
UNIT = "candidates/s"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, lines):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, power, reasons = [], [], [], set()
        for ln in lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_info():
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"nproc": os.cpu_count(), "cpu_model": model}


def cpu_rate(seconds_target: float, lo0: int, stream_version: int = 2):
    """Oracle port of build_lhd's map+reduce on the host cores; bounded sample."""
    import oracle
    oracle.set_stream_version(stream_version)
    oracle.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    probe = 100_000
    oracle.search_range(N_POINTS, N_DIMS, oracle.MAXPRO, SEED, lo0, lo0 + probe)
    dt = time.perf_counter() - t0
    sample = max(probe, int(probe / max(dt, 1e-6) * seconds_target))
    t0 = time.perf_counter()
    oracle.search_range(N_POINTS, N_DIMS, oracle.MAXPRO, SEED, lo0 + probe, lo0 + probe + sample)
    dt = time.perf_counter() - t0
    return sample / dt, sample, dt, oracle.num_threads()


def run_reference(args, rank, world, emit):
    """--impl reference: the reference algorithm on the host CPU, rank 0 only."""
    if rank != 0:
        return
    import oracle
    oracle.set_stream_version(args.stream_version)
    oracle.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
    cores = oracle.num_threads()
    probe = 200_000
    t0 = time.perf_counter()
    oracle.search_range(N_POINTS, N_DIMS, oracle.MAXPRO, SEED, 0, probe)
    rate0 = probe / (time.perf_counter() - t0)
    budget_s = 120.0  # whole run (warm-up + steps) stays within a couple of minutes
    per_step = max(50_000, int(rate0 * budget_s / max(1, args.steps + args.warmup)))
    per_step = min(per_step, int(rate0 * 20))
    lo = 0
    for _ in range(args.warmup):
        oracle.search_range(N_POINTS, N_DIMS, oracle.MAXPRO, SEED, lo, lo + per_step); lo += per_step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.search_range(N_POINTS, N_DIMS, oracle.MAXPRO, SEED, lo, lo + per_step); lo += per_step
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} candidates/step x {args.steps} steps (n=50, d=3, MaxPro, seed 42), OpenMP over candidates"
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "maxpro_search_n50_d3 (CPU port of src/lib.rs:65-98; Rust toolchain absent, see DESIGN.md)",
                   "candidates_per_step": per_step, "seed": SEED, "host": host_info()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)



def newest_profile_traffic(kernel="search_screen"):
    """dram__bytes_read.sum + dram__bytes_write.sum of the headline kernel from the newest
    profiles/*_<kernel>_metrics.csv (one launch under `ncu --set full`); None if no capture."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_{kernel}_metrics.csv")), key=os.path.basename)  # r1k < r2a < ...: round, then pass
    for path in reversed(files):
        got = {}
        try:
            for row in csv.reader(open(path)):
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    if key in row:
                        i = row.index(key)
                        unit, val = row[i + 1].lower(), float(row[i + 2].replace(",", ""))
                        got[key] = val * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        except (OSError, ValueError, IndexError):
            continue
        if len(got) == 2:
            return int(got["dram__bytes_read.sum"] + got["dram__bytes_write.sum"]), os.path.basename(path)
    return None, None


# ---------------------------------------------------------------- BASELINE.json configs C1-C5
# Short in-run measurements of the five configs BASELINE.json names, each with its own parity check and its own
# roofline (algorithmic flops of SURVEY.md 8d over the measured time, against the FP64 peak probed in this run).
CONFIG_NAMES = ("C1", "C2", "C3", "C4", "C5")


def _best_of(fn, reps=3):
    best, out = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return out, best


def run_configs(which, rank, world, dev, peak_tflops, peak32_tflops, max_over_ranks, barrier):
    """Returns {name: record}.  Sequential paths (C1 CLI, C5 ordering) and the millisecond-sized C2 run on rank 0
    (replicas only); C3 shards candidate-index blocks and C4 chains across the ranks."""
    import numpy as np

    import oracle
    from maxpro_b200 import METRIC_MAXIMIN, METRIC_MAXPRO, engine, shard

    out = {}

    def roof(rate, flops_per_unit, note=None, bound="fp64"):
        # bound = the non-tensor FMA pipe the dominant kernel runs on: fp32 for the two-stage search (screening kernel)
        a = rate * flops_per_unit / 1e12
        peak = (peak32_tflops if bound == "fp32" else peak_tflops) * world
        r = {"bound": bound, "achieved": a, "peak": peak, "unit": "TFLOP/s",
             "frac": a / peak, "flops_per_unit": flops_per_unit, "traffic": None}
        if note:
            r["note"] = note
        return r

    if "C1" in which and rank == 0:
        # cargo run --release -- --iterations 100000 --samples 50 --ndims 2 --metric max-pro  (src/main.rs)
        n, d = 50, 2
        cli = os.path.join(ROOT, "maxpro_b200", "bin", "maxpro")
        argv = [cli, "--iterations", "100000", "--samples", "50", "--ndims", "2", "--metric", "max-pro", "--seed", "42"]
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(dev.index))
        r, wall = _best_of(lambda: subprocess.run(argv, capture_output=True, text=True, env=env), 3)
        lines = r.stdout.strip().splitlines()
        metrics = {}
        for ln in lines:
            for key in ("Original metric", "Swapped metric", "Annealed metric"):
                if ln.startswith(key + ":"):
                    metrics[key] = float(ln.split(":", 1)[1])
        # what any process pays before its first kernel: a bare cudaFree(0) in a fresh process (tools/micro/ctx.cu, built by
        # __graft_entry__.build()) -- the floor of the CLI's wall time on this box
        ctx_s = None
        ctx_bin = os.path.join(ROOT, "tools", "micro", "ctx")
        if os.path.exists(ctx_bin):
            try:
                vals = []
                for _ in range(2):
                    pr = subprocess.run([ctx_bin], capture_output=True, text=True, env=env)
                    vals.append(float(pr.stdout.split(":")[1].split()[0]) * 1e-3)
                ctx_s = min(vals)
            except Exception:  # noqa: BLE001 -- the probe is informational
                ctx_s = None
        N = 1 << 26
        (_, s, i), t = _best_of(lambda: engine.build_lhd(n, d, N, METRIC_MAXPRO, 42), 2)
        flops = (3 * d + 3) * n * (n - 1) // 2 + 2
        ok = (r.returncode == 0 and len(metrics) == 3 and
              metrics["Annealed metric"] <= metrics["Swapped metric"] <= metrics["Original metric"])
        out["C1"] = {"workload": "CLI --iterations 100000 --samples 50 --ndims 2 --metric max-pro --seed 42 "
                                 "(search + 20,000 swap + 80,000 jitter annealing steps, process start included)",
                     "value": wall, "unit": "s (CLI wall, best of 3)", "cli_rc": r.returncode, "cli_metrics": metrics,
                     "cuda_context_s": ctx_s,  # a bare cudaFree(0) in a fresh process on this box: the floor of any GPU CLI call
                     "check": {"annealed <= swapped <= original": ok},
                     "search_rate": {"value": N / t, "unit": "candidates/s", "candidates": N},
                     "roofline": roof(N / t, flops, "n=50, d=2 MaxPro search at 2^26 candidates through build_lhd "
                                                             "(two-stage exact search: fp32 screening kernel dominant)", "fp32")}

    if "C2" in which and rank == 0:
        n, d, N = 50, 3, 1_000_000
        (s, i), t = _best_of(lambda: engine.search_range(n, d, METRIC_MAXIMIN, 42, 0, N))
        oracle.set_num_threads(os.cpu_count() or 1)
        s_ref, i_ref = oracle.search_range(n, d, oracle.MAXIMIN, 42, 0, N)
        Nb = 1 << 27
        _, tb = _best_of(lambda: engine.search_range(n, d, METRIC_MAXIMIN, 42, N, N + Nb), 2)
        flops = (3 * d + 2) * n * (n - 1) // 2
        out["C2"] = {"workload": "maximin random search --iterations 1000000 --samples 50 --ndims 3 --seed 42",
                     "value": t * 1e3, "unit": "ms (host call, best of 3)", "winner": {"score": s, "index": i},
                     "check": {"winner == CPU oracle (bit-exact score, same index)": (s, i) == (s_ref, i_ref)},
                     "steady_rate": {"value": Nb / tb, "unit": "candidates/s", "candidates": Nb},
                     "roofline": roof(Nb / tb, flops, "steady rate at 2^27 candidates (two-stage exact search: fp32 screening "
                                                              "kernel dominant; the returned distance is the fp64 kernel's, bit-exact)", "fp32")}

    if "C3" in which:
        n, d, N = 500, 10, 10_000_000
        lo, hi = rank * N // world, (rank + 1) * N // world
        engine.search_range(n, d, METRIC_MAXPRO, 42, 0, 2000)  # module load, workspace
        barrier()
        t0 = time.perf_counter()
        s, i = engine.search_range(n, d, METRIC_MAXPRO, 42, lo, hi)
        if world > 1:
            s, i = shard.allgather_minloc(s, i, METRIC_MAXPRO, device=dev)  # 16 B per rank over NCCL
        barrier()
        t = max_over_ranks(time.perf_counter() - t0)
        flops = (3 * d + 3) * n * (n - 1) // 2 + 2
        if rank == 0:
            Nc = 3000
            oracle.set_num_threads(os.cpu_count() or 1)
            s_ref, i_ref = oracle.search_range(n, d, oracle.MAXPRO, 42, 0, Nc)
            s_gpu, i_gpu = engine.search_range(n, d, METRIC_MAXPRO, 42, 0, Nc)
            out["C3"] = {"workload": f"MaxPro random search n=500 d=10, 1e7 candidates in index blocks over {world} GPU(s), "
                                     "16-byte min-loc all-gather",
                         "value": N / t, "unit": "candidates/s", "seconds": t, "winner": {"score": s, "index": i},
                         "check": {f"first {Nc} candidates: winner index == CPU oracle, score within 1e-9":
                                   bool(i_gpu == i_ref and abs(s_gpu - s_ref) <= 1e-9 * s_ref)},
                         "roofline": roof(N / t, flops)}

    if "C4" in which:
        n, d, chains, steps = 1000, 20, 65536, 1000
        x = engine.generate_lhd(n, d, 42)
        rec = {"workload": f"anneal_lhd swap moves n=1000 d=20, {chains} chains per GPU x {steps} steps, T0=1, cooling 0.99, "
                           f"start = generate_lhd(1000, 20, 42), chain c of rank r draws from stream 43 + r*{chains} + c",
               "unit": "chain-steps/s"}
        for name, metric, minimize, per_pair in (("maximin", METRIC_MAXIMIN, False, 3 * d + 2), ("maxpro", METRIC_MAXPRO, True, 3 * d + 3)):
            engine.anneal_lhd(x, 10, 1.0, 0.99, metric, minimize, 43, True, 148)
            barrier()
            t0 = time.perf_counter()
            best_design, best, chain = engine.anneal_lhd(x, steps, 1.0, 0.99, metric, minimize, 43 + rank * chains, True, chains)
            barrier()
            t = max_over_ranks(time.perf_counter() - t0)
            if rank == 0:
                start = oracle.maximin_criterion(x) if metric == METRIC_MAXIMIN else oracle.maxpro_criterion(x)
                rescored = oracle.maximin_criterion(best_design) if metric == METRIC_MAXIMIN else oracle.maxpro_criterion(best_design)
                never_worse = best >= start if not minimize else best <= start
                rate = chains * world * steps / t
                rec[name] = {"value": rate, "seconds": t, "start_metric": start, "best_metric": best, "best_chain": chain,
                             "check": {"never worse than the input (anneal.rs:92-94,142)": bool(never_worse),
                                       "returned design re-scored by the CPU oracle": bool(abs(rescored - best) <= 1e-9 * abs(best))},
                             "roofline": roof(rate, 4 * (n - 2) * per_pair,
                                              "incremental step: the 2(n-2) pairs of the two swapped rows, old and new"
                                              + ("; the kernel screens these pairs on an 8-bit copy of the design (IDP.4A) and evaluates "
                                                 "in fp64 only the ~14 per step that can attain a minimum, so the FP64 pipe is ~10 % busy (ncu) -- "
                                                 "the figure is the exact step's work per second, not pipe utilisation" if metric == METRIC_MAXIMIN else ""))}
        if rank == 0:
            rec["value"] = rec["maximin"]["value"]
            rec["roofline"] = rec["maximin"]["roofline"]
            rec["note"] = ("headline of this config = maximin (trajectory bit-exact against the oracle in the tests). "
                           "MaxPro at d=20 is numerically degenerate in the reference itself: every product of 20 squared "
                           "differences is far below eps=1e-12, so psi = 3.981 for every design and accept/reject is rounding "
                           "noise; its rate is reported as throughput only")
            out["C4"] = rec

    if "C5" in which and rank == 0:
        n, d = 5000, 8
        x = engine.generate_lhd(n, d, 42)
        (ordered, perm), t = _best_of(lambda: engine.order_design(x, METRIC_MAXPRO, True))
        _, perm_ref = oracle.order_design(x, oracle.MAXPRO, True, True)
        flops = (3 * d + 3) * n * (n - 1) // 2
        out["C5"] = {"workload": "order_design(generate_lhd(5000, 8, seed 42), maxpro): greedy ordering, 4,999 dependent argmin steps",
                     "value": t * 1e3, "unit": "ms (host call, best of 3)",
                     "check": {"permutation == incremental CPU oracle (src/order.rs:34-105 tie rules)": perm.tolist() == perm_ref.tolist()},
                     "roofline": roof(1.0 / t, flops, "latency-bound: one cluster of CTAs, n-1 sequential steps; "
                                                      "the reference's own O(n^4 d) formulation would be 7e14 flops")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--candidates", type=int, default=CANDIDATES_PER_GPU, help="candidates per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="headline", choices=["headline", "all"] + list(CONFIG_NAMES),
                    help="headline (default): the BASELINE.json metric plus a short `configs` block with C1-C5; "
                         "C1..C5: that config alone; all: C1-C5 without the headline")
    ap.add_argument("--no-configs", action="store_true", help="headline without the C1-C5 block")
    ap.add_argument("--stream-version", type=int, default=2, choices=[1, 2],
                    help="design stream: 2 = LHD-mix v2 (default), 1 = LHD-Philox v1")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly ONE line (the JSON).  Libraries write there too (NCCL prints its
    # version banner on fd 1), so everything else goes to stderr and the line is written to the
    # saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    from maxpro_b200 import METRIC_MAXPRO, _lib, engine, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: maxpro_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    _lib.set_device(local_rank)
    _lib.set_stream_version(args.stream_version)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.config != "headline":
        # one BASELINE.json config (or all five) instead of the headline metric
        which = set(CONFIG_NAMES) if args.config == "all" else {args.config}
        probed, _ = engine.fp64_peak_probe(10)
        peak = probed if probed >= 0.95 * NOMINAL_FP64_TFLOPS else NOMINAL_FP64_TFLOPS
        t0 = time.perf_counter()
        probed32, _ = engine.fp32_peak_probe(10)
        peak32 = probed32 if probed32 >= 0.95 * NOMINAL_FP32_TFLOPS else NOMINAL_FP32_TFLOPS
        recs = run_configs(which, rank, world, dev, peak, peak32, max_over_ranks, barrier)
        wall = time.perf_counter() - t0
        if rank == 0:
            first = recs[sorted(recs)[0]]
            emit({"metric": "maxpro_baseline_config_" + args.config, "value": first["value"], "unit": first["unit"],
                  "n_gpus": world, "steps": 1, "warmup": 0, "ms_per_step": wall * 1e3,
                  "higher_is_better": not first["unit"].startswith(("s ", "ms ")), "scaling": "weak", "vs_baseline": None,
                  "dtype": "f64", "data": "synthetic", "config": {"workload": first["workload"]},
                  "roofline": first["roofline"], "configs": recs, "gpu_launches": int(_lib.kernel_launch_count())})
        if world > 1:
            dist.destroy_process_group()
        return

    L = _lib.lib()
    per_gpu = args.candidates
    ws_bytes = C.c_size_t(0)
    _lib.check(L.maxpro_search_workspace_bytes(N_POINTS, N_DIMS, C.byref(ws_bytes)))
    ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
    result = torch.zeros(2, dtype=torch.int64, device=dev)  # {f64 score bits, u64 index}
    stream = torch.cuda.current_stream()

    def block(step: int):
        lo = (step * world + rank) * per_gpu  # global candidate indices, disjoint per (step, rank)
        return lo, lo + per_gpu

    def device_step(step: int, flags: int = 0):
        lo, hi = block(step)
        _lib.check(L.maxpro_search_range_device(N_POINTS, N_DIMS, METRIC_MAXPRO, SEED, lo, hi, flags,
                                                result.data_ptr(), ws.data_ptr(), ws_bytes.value,
                                                stream.cuda_stream))

    # ---------------- value: device-resident, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~1 s to come up: start it before the warm-up steps
    for s in range(args.warmup):
        device_step(s)
    barrier()
    sampler.lines.clear()  # keep only samples taken inside the timed region
    launches0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record(stream)
    for s in range(args.steps):
        device_step(args.warmup + s)
        marks[s].record(stream)
    e1.record(stream)
    barrier()
    if rank == 0:
        prev, per = e0, []
        for m in marks:
            per.append(prev.elapsed_time(m)); prev = m
        print("device per-step ms: " + " ".join(f"{v:.1f}" for v in per), file=sys.stderr)
    launches = _lib.kernel_launch_count() - launches0
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    timed_lines = list(sampler.lines)
    # stop polling before the host-facing leg: NVML queries and cudaMalloc/cudaFree contend for
    # driver locks, which showed up as 100+ ms stalls in e2e steps
    sampler.stop()
    value = per_gpu * world * args.steps / (dev_ms * 1e-3)
    per_gpu_rate = per_gpu * args.steps / (dev_ms * 1e-3)
    kernel_ms = dev_ms / args.steps  # one criterion-kernel launch per step (+ a 1-thread counter reset)
    achieved_tflops = per_gpu_rate * FLOPS_PER_CANDIDATE / 1e12
    rec = result.cpu().numpy()
    last_score, last_index = float(rec[:1].view(np.float64)[0]), int(rec[1:].view(np.uint64)[0])

    # Pipe peaks of this GPU (roofline denominators), probed right after the timed region so the clocks are the
    # ones the search ran at.  A probe that lands below the datasheet-level figure by more than 5% (clock ramp,
    # a neighbour's power draw) is not used: the nominal figure is the more conservative denominator then.
    probed_tflops, _ = engine.fp64_peak_probe(10)
    peak_tflops = probed_tflops if probed_tflops >= 0.95 * NOMINAL_FP64_TFLOPS else NOMINAL_FP64_TFLOPS
    probed32_tflops, _ = engine.fp32_peak_probe(10)
    peak32_tflops = probed32_tflops if probed32_tflops >= 0.95 * NOMINAL_FP32_TFLOPS else NOMINAL_FP32_TFLOPS

    # ---------------- fp64_only: every candidate scored in fp64 (the kernel the FP64 roofline describes)
    def read_result():
        r = result.cpu().numpy()
        return float(r[:1].view(np.float64)[0]), int(r[1:].view(np.uint64)[0])
    device_step(0, flags=FLAG_DEFAULT)
    torch.cuda.synchronize()
    default_winner = read_result()
    device_step(0, flags=FLAG_FP64_ONLY)
    torch.cuda.synchronize()
    exact_winner = read_result()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for s in range(args.steps):
        device_step(args.warmup + s, flags=FLAG_FP64_ONLY)
    s1.record(stream)
    barrier()
    fp64_ms = max_over_ranks(s0.elapsed_time(s1))
    fp64_rate = per_gpu * args.steps / (fp64_ms * 1e-3)
    fp64_only = {"value": fp64_rate * world, "unit": UNIT, "ms_per_step": fp64_ms / args.steps,
                 "same_winner_as_default_path": default_winner == exact_winner,
                 "roofline": {"bound": "fp64", "achieved": fp64_rate * FLOPS_PER_CANDIDATE / 1e12, "peak": peak_tflops,
                              "unit": "TFLOP/s", "frac": fp64_rate * FLOPS_PER_CANDIDATE / 1e12 / peak_tflops,
                              "kernel": "search_cyclic_kernel<50,3,192,true,false,2>", "peak_probed": probed_tflops,
                              "traffic": newest_profile_traffic("search_cyclic")[0],
                              "note": "MAXPRO_FLAG_FP64_ONLY; the kernel issues ~8.3 FP64-pipe instructions per pair for 12 "
                                      "algorithmic flops, see profiles/ for pipe utilisation"}}

    # ---------------- e2e: the host-facing call, winner design back in host memory
    base = args.warmup + args.steps
    def host_step(step: int):
        lo, hi = block(step)
        if world == 1:
            design, score, index = engine.build_lhd(N_POINTS, N_DIMS, per_gpu, METRIC_MAXPRO, SEED + lo)
            return score, index + lo
        score, index = engine.search_range(N_POINTS, N_DIMS, METRIC_MAXPRO, SEED, lo, hi)
        score, index = shard.allgather_minloc(score, index, METRIC_MAXPRO, device=dev)  # 16 B per rank over NCCL
        if rank == 0:
            engine.generate_lhd(N_POINTS, N_DIMS, SEED + index)  # winner rebuilt from (seed, index)
        return score, index
    for s in range(min(args.warmup, 2)):
        host_step(base + s)
    barrier()
    t0 = time.perf_counter()
    step_ms = []
    for s in range(args.steps):
        t1 = time.perf_counter()
        host_step(base + 2 + s)
        step_ms.append((time.perf_counter() - t1) * 1e3)
    barrier()
    if rank == 0:
        print("e2e per-step ms: " + " ".join(f"{v:.1f}" for v in step_ms), file=sys.stderr)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = per_gpu * world * args.steps / e2e_s
    clocks = None
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary(timed_lines)  # clocks seen during the device-timed region

    # ---------------- K2 through its host-buffer entry point (maxpro_score_batch: the one call of the path that takes
    # input buffers -- host designs in, host scores out; copies inside the timed region)
    score_e2e = None
    if rank == 0:
        nb = 1 << 17
        designs = np.random.default_rng(7).random((nb, N_POINTS, N_DIMS))
        engine.score_batch(designs[:1024], METRIC_MAXPRO)
        best_t = None
        for _ in range(3):
            t1 = time.perf_counter()
            sc, arg = engine.score_batch(designs, METRIC_MAXPRO)
            dt = time.perf_counter() - t1
            best_t = dt if best_t is None else min(best_t, dt)
        score_e2e = {"value": nb / best_t, "unit": "supplied designs scored/s (n=50, d=3, MaxPro)", "designs": nb,
                     "h2d_bytes_per_call": int(designs.nbytes), "d2h_bytes_per_call": nb * 8 + 16,
                     "h2d_gb_per_s": designs.nbytes / best_t / 1e9,
                     "note": "maxpro_score_batch with pageable host buffers, best of 3: bound by the host-to-device copy (staged through "
                             "pinned buffers by several host threads, ~2x a plain cudaMemcpy), not by the kernel (1.07e9 designs/s "
                             "device-resident, profiles/)"}

    traffic, traffic_src = newest_profile_traffic()
    line = {
        "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 screen + f64 re-score (fp64_only: f64)", "data": "synthetic",
        "config": {"workload": "maxpro_search_n50_d3 (fused LHD generation + MaxPro criterion + on-chip argmin)",
                   "n": N_POINTS, "d": N_DIMS, "candidates_per_gpu_per_step": per_gpu, "seed": SEED,
                   "sharding": f"candidate-index blocks x{world}, min-loc all-gather (16 B/rank)",
                   "l2": "n/a: the search reads no HBM inputs (candidates are generated in shared memory); "
                         "each step scores fresh candidate indices, nothing is cached between steps",
                   "last_winner": {"score": last_score, "index": last_index}},
        "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": peak32_tflops, "unit": "TFLOP/s",
                     "frac": achieved_tflops / peak32_tflops if peak32_tflops else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch under `ncu --set full`, read from the
                     # newest profiles/*_search_screen_metrics.csv (instruction/constant fetch plus the survivor
                     # list; it does not grow with the candidate count).  Algorithmic HBM bytes per candidate: 0.
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": ("FFMA probe measured in this run" if peak32_tflops == probed32_tflops else
                                     "nominal (the in-run FFMA probe came out lower and was discarded)")
                                    + "; MEASURED_PEAKS.json has no non-tensor FP32/FP64 figure; "
                                    f"nominal 148 SM x 128 lanes x 2 x 1.965 GHz = {NOMINAL_FP32_TFLOPS:.1f}",
                     "peak_probed": probed32_tflops,
                     "flops_per_candidate": FLOPS_PER_CANDIDATE, "kernel": "search_screen_kernel<50,3,384,2>",
                     "kernel_ms_per_launch": kernel_ms, "candidates_per_launch": per_gpu,
                     "note": "algorithmic flops of the criterion (12 per pair) over the whole step -- the screening kernel "
                             "plus the fp64 re-score of its survivors and a 1-thread counter reset, so the kernel's own "
                             "fraction is slightly higher; the screening kernel issues 8 fp32 lane-operations per pair. "
                             "`fp64_only.roofline` is the FP64-pipe figure BASELINE.json's metric asks for"},
        "path": "two-stage exact search (engine default for this shape): fp32 screening with a proven bound + fp64 re-score "
                "of every candidate that can still win; winner, score and tie-break identical to the fp64-only search",
        "fp64_only": fp64_only,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 16 + N_POINTS * N_DIMS * 8,
                "note": "reference API takes scalars (n, d, iterations, metric, seed): no input buffers exist; "
                        "the winner record and the regenerated design are copied to host every step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "score_batch_e2e": score_e2e,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, sample, dt, cores = cpu_rate(15.0, 1 << 40, args.stream_version)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{sample} candidates (n=50, d=3, MaxPro) in {dt:.1f} s, OpenMP over candidates, "
                                          f"{host_info()['cpu_model']}"}
    if not args.no_configs:
        # the headline line must survive a failing config: one config at a time, an error string in its place
        recs = {}
        for name in CONFIG_NAMES:
            try:
                recs.update(run_configs({name}, rank, world, dev, peak_tflops, peak32_tflops, max_over_ranks, barrier))
            except Exception as exc:  # noqa: BLE001 -- reported, not swallowed
                if world > 1:
                    raise  # a rank that skips a collective would hang the others: fail loudly under torchrun
                recs[name] = {"error": f"{type(exc).__name__}: {exc}"}
        line["configs"] = recs
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
