#!/usr/bin/env python
"""Raw NCCL all-reduce time of the hash-table gradient payload (1.46 GB fp32), whole and in per-level slabs.
torchrun --nproc-per-node N tools/bench_allreduce.py"""
import os
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 365792384
x = torch.zeros(n, device="cuda")
for label, slabs in (("whole", 1), ("11 slabs", 11), ("22 slabs", 22)):
    bounds = [n * i // slabs for i in range(slabs + 1)]
    for it in range(3):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for a, b in zip(bounds[:-1], bounds[1:]):
            dist.all_reduce(x[a:b], op=dist.ReduceOp.AVG)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    w = dist.get_world_size()
    if dist.get_rank() == 0:
        print(f"{label}: {ms:.2f} ms  algbw {n * 4 / ms / 1e6:.0f} GB/s  busbw {n * 4 / ms / 1e6 * 2 * (w - 1) / w:.0f} GB/s", flush=True)
dist.destroy_process_group()
