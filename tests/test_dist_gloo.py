"""Multi-rank host logic on CPU (gloo, world_size 2): gradient all-reduce(mean) semantics of GradReducer and the ray
partition helper.  The GPU runs use the same class over NCCL."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mli_nerf_b200.dist import GradReducer, shard_rays


def _worker(rank, world, port_no, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.ModuleDict(dict(table=torch.nn.Embedding(1 << 19, 8), head=torch.nn.Linear(7, 3),
                                     frozen=torch.nn.Linear(3, 3)))
    g = torch.Generator().manual_seed(100 + rank)
    model["table"].weight.grad = torch.randn(model["table"].weight.shape, generator=g)
    model["head"].weight.grad = torch.randn(3, 7, generator=g)
    model["head"].bias.grad = torch.randn(3, generator=g)
    n = model["table"].weight.numel()
    red = GradReducer(model, world, level_slices=[(0, n // 3), (n // 3, n)], side_stream=False)
    red.allreduce_grads()
    torch.save({k: p.grad for k, p in model.named_parameters() if p.grad is not None}, out + f".{rank}")
    dist.destroy_process_group()


def test_grad_allreduce_mean_world2(tmp_path):
    world, out = 2, str(tmp_path / "g")
    mp.spawn(_worker, args=(world, 29731, out), nprocs=world, join=True)
    res = [torch.load(out + f".{r}") for r in range(world)]
    assert set(res[0]) == {"table.weight", "head.weight", "head.bias"}  # parameters without .grad are skipped
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    exp_table = sum(torch.randn(1 << 19, 8, generator=g) for g in gens) / world
    exp_w = sum(torch.randn(3, 7, generator=g) for g in gens) / world
    for r in range(world):
        assert torch.allclose(res[r]["table.weight"], exp_table, atol=1e-6)
        assert torch.allclose(res[r]["head.weight"], exp_w, atol=1e-6)
        assert torch.equal(res[r]["head.bias"], res[0]["head.bias"])


def test_shard_rays_partition():
    for n, w in ((640000, 8), (97200, 4), (10, 3), (5, 8)):
        spans = [shard_rays(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
