"""Launched by tests/test_gpu_multi.py through torch.distributed.run on >= 2 GPUs: PeerTableReducer (copy-engine exchange
over NVLink peer memory) against NCCL's all-reduce(mean) of the same buffers, uneven slabs, several steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from mli_nerf_b200.dist import PeerTableReducer  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = 8 * 1_000_003  # not a multiple of the shard granularity
slabs = [(0, 8 * 1000), (8 * 1000, 8 * 300_000), (8 * 300_000, 8 * 300_001), (8 * 300_001, n)]
red = PeerTableReducer(n, torch.device("cuda", local), max_slabs=len(slabs))
g = torch.Generator(device="cuda").manual_seed(1234 + rank)
worst = 0.0
for step in range(4):
    red.buf.zero_()
    x = torch.randn(n, device="cuda", generator=g) * (1.0 + rank)
    want = x.clone()
    dist.all_reduce(want, op=dist.ReduceOp.AVG)
    for a, b in slabs:  # the "scatter" of each slab, followed by its exchange
        red.buf[a:b].copy_(x[a:b])
        red.reduce_slab(a, b)
    red.finish()
    err = float((red.buf - want).abs().max())
    worst = max(worst, err)
    assert err <= 1e-6 * float(want.abs().max()) + 1e-7, (step, rank, err)
# all ranks hold the identical (bitwise) result: each shard is summed on exactly one rank
chk = red.buf.double().sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
assert all(float(c) == float(allc[0]) for c in allc)
red.close()
if rank == 0:
    print(f"PEER_ALLREDUCE_OK world={world} max_err={worst:.3e}", flush=True)
dist.destroy_process_group()
