#!/usr/bin/env python3
"""Concurrent D2H bandwidth of all ranks into pinned host memory, with and without binding each rank
to its GPU's NUMA node (what bounds bench.py's e2e arm at N > 1).
    python -m torch.distributed.run --nproc-per-node N tools/d2h_diag.py"""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)


def run(tag):
    host = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    for _ in range(2):
        host.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        host.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = 8 * (1 << 30) / dt / 1e9
    t = torch.tensor([gbs], device=dev)
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        vals = [float(x) for x in allv]
    else:
        vals = [gbs]
    if rank == 0:
        print(f"{tag}: per-rank GB/s {[round(v, 1) for v in vals]} sum {sum(vals):.1f}", file=sys.stderr, flush=True)
    del host


run("unbound")
node = bench.bind_to_gpu_numa_node(lr)
print(f"rank {rank}: numa node {node}, cpus {len(os.sched_getaffinity(0))}", file=sys.stderr, flush=True)
run("bound  ")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
