"""Pinned-memory D2H bandwidth of the box, from 1 up to WORLD_SIZE concurrent ranks: the ceiling under bench.py's e2e number.

One rank:   python tools/d2h_probe.py
All ranks:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/d2h_probe.py

Each rank owns one GPU and copies a 64 MiB buffer (one step's observations at 4096 x 128 x 128) device -> pinned host memory,
20 times back to back; for k = 1, 2, 4, ... the first k ranks copy CONCURRENTLY (the others idle) and the aggregate is
k x bytes / max-over-ranks time.  Variants: `plain` (what TactileVecEnv does), `pinned-cpu` (the rank's host thread and its
pinned buffer bound to a disjoint core set before the allocation, so first-touch places the pages next to those cores), `chunks16`
(16 chunk copies instead of one), and cudaHostAlloc'd write-combined memory.  Prints one JSON line per (variant, k) on rank 0 and
the host topology nvidia-smi reports.
"""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 64 * 1024 * 1024
d = torch.empty(N, dtype=torch.uint8, device=dev)


def bind_cores():
    """a disjoint slice of the cores this process may run on"""
    cores = sorted(os.sched_getaffinity(0))
    per = max(1, len(cores) // max(1, world))
    mine = cores[rank * per:(rank + 1) * per] or cores
    os.sched_setaffinity(0, mine)
    return mine


def run(variant, k):
    if variant == "pinned-cpu":
        bind_cores()
    h = torch.empty(N, dtype=torch.uint8).pin_memory()
    h.fill_(1)        # first touch under the current affinity
    chunks = 16 if variant == "chunks16" else 1
    step = N // chunks

    def copy():
        for c in range(chunks):
            h[c * step:(c + 1) * step].copy_(d[c * step:(c + 1) * step], non_blocking=True)
    active = rank < k
    for _ in range(3):
        if active:
            copy()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    if active:
        for _ in range(20):
            copy()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) if active else 0.0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        agg = k * 20 * N / (float(t) * 1e-3) / 1e9
        print(json.dumps({"variant": variant, "ranks": k, "aggregate_GBps": round(agg, 1), "per_rank_GBps": round(agg / k, 1),
                          "ms_per_64MiB": round(float(t) / 20, 3)}), flush=True)
    del h


if rank == 0:
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[-2500:], flush=True)
        print("cores visible:", len(os.sched_getaffinity(0)), flush=True)
    except Exception as e:  # noqa: BLE001
        print("topo unavailable:", e)
ks = [k for k in (1, 2, 4, 8) if k <= world]
for variant in ("plain", "chunks16", "pinned-cpu"):     # pinned-cpu last: it narrows the affinity for good
    for k in ks:
        run(variant, k)
if world > 1:
    dist.destroy_process_group()
