"""Aggregate host<->device copy bandwidth with one process per GPU, all at once (torchrun): the platform ceiling of the
end-to-end path (every step uploads the shard's genotypes and downloads its results).
usage: python -m torch.distributed.run --nproc-per-node N profiles/r2_pcie_aggregate.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
up = torch.empty(192 << 20, dtype=torch.uint8).pin_memory()
down = torch.empty(48 << 20, dtype=torch.uint8).pin_memory()
d_up, d_down = torch.empty_like(up, device="cuda"), torch.empty_like(down, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for mode in ("h2d", "d2h", "both"):
    for it in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_up.copy_(up, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    down.copy_(d_down, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    nbytes = 20 * ((up.numel() if mode != "d2h" else 0) + (down.numel() if mode != "h2d" else 0))
    t = torch.tensor([nbytes / dt / 1e9], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    res[mode] = {"per_rank_gbs": nbytes / dt / 1e9, "aggregate_gbs": float(t[0])}
if rank == 0:
    print(json.dumps({"n_gpus": world, "copy": "192 MiB up / 48 MiB down per rank per iteration, pinned host memory", **res}))
if world > 1:
    dist.destroy_process_group()
