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


def _worker_hook(rank, world, port_no, out):
    """Overlap path: slabs of the big gradient go through the engine hook during 'backward'; the final call must only
    reduce what is left (and must not average the table a second time)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = torch.nn.ModuleDict(dict(table=torch.nn.Embedding(1 << 19, 8), head=torch.nn.Linear(7, 3)))
    g = torch.Generator().manual_seed(200 + rank)
    tg = torch.randn(model["table"].weight.numel(), generator=g)
    red = GradReducer(model, world, side_stream=False)

    class Eng:
        table_grad_hook = None
    eng = Eng()
    red.attach(eng)
    n = tg.numel()
    for a, b in ((0, n // 4), (n // 4, n // 2), (n // 2, n)):
        eng.table_grad_hook(tg, a, b)
    model["table"].weight.grad = tg.view_as(model["table"].weight)
    model["head"].weight.grad = torch.randn(3, 7, generator=g)
    red.allreduce_grads()
    torch.save({k: p.grad for k, p in model.named_parameters() if p.grad is not None}, out + f".{rank}")
    # a .grad that does not alias the buffer the hook exchanged (accumulation into an existing gradient) is refused
    for a, b in ((0, n // 2), (n // 2, n)):
        eng.table_grad_hook(tg, a, b)
    model["table"].weight.grad = tg.clone().view_as(model["table"].weight)
    try:
        red.allreduce_grads()
        refused = False
    except RuntimeError:
        refused = True
    assert refused
    dist.destroy_process_group()


def test_grad_allreduce_overlap_hook_world2(tmp_path):
    world, out = 2, str(tmp_path / "h")
    mp.spawn(_worker_hook, args=(world, 29741, out), nprocs=world, join=True)
    res = [torch.load(out + f".{r}") for r in range(world)]
    gens = [torch.Generator().manual_seed(200 + r) for r in range(world)]
    tabs = [torch.randn((1 << 19) * 8, generator=g) for g in gens]
    heads = [torch.randn(3, 7, generator=g) for g in gens]
    for r in range(world):
        assert torch.allclose(res[r]["table.weight"].reshape(-1), sum(tabs) / world, atol=1e-6)
        assert torch.allclose(res[r]["head.weight"], sum(heads) / world, atol=1e-6)


def test_shard_rays_partition():
    for n, w in ((640000, 8), (97200, 4), (10, 3), (5, 8)):
        spans = [shard_rays(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))


def test_peer_shard_bounds_partition_every_slab():
    from mli_nerf_b200.dist import shard_bounds
    for world in (2, 3, 4, 8):
        for a, b in ((0, 8), (0, 8 * 1000), (8 * 1000, 8 * 300_000), (64, 64), (8, 8 * 1_000_003)):
            parts = [shard_bounds(a, b, r, world) for r in range(world)]
            assert parts[0][0] == a and parts[-1][1] == b
            assert all(x[1] == y[0] for x, y in zip(parts, parts[1:]))
            assert all((p0 - a) % 4 == 0 for p0, _ in parts)
            slot = ((b - a + world - 1) // world + 3) // 4 * 4
            assert all(0 <= p1 - p0 <= slot for p0, p1 in parts)


def _worker_gather(rank, world, port_no, out):
    """Ray-sharded inference: every rank holds the outputs of its contiguous ray range; Model._gather_rays rebuilds the
    full per-ray tensors on every rank (ragged last shard, a bool tensor, a rank-4 per-sample tensor)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mli_nerf_b200.model import Model
    n, B = 1001, 2
    r0, r1 = shard_rays(n, rank, world)
    idx = torch.arange(r0, r1, dtype=torch.float32)
    local = dict(rgb=(idx[None, :, None] * torch.tensor([1.0, 2.0, 3.0])).expand(B, -1, -1).contiguous(),
                 outside=(idx.long() % 3 == 0)[None, :, None].expand(B, -1, -1).contiguous(),
                 weights=idx[None, :, None, None].expand(B, -1, 4, 1).contiguous())
    full = Model._gather_rays(local, n, world)
    torch.save(full, out + f".{rank}")
    dist.destroy_process_group()


def test_sharded_inference_gather_world3(tmp_path):
    world, out = 3, str(tmp_path / "i")
    mp.spawn(_worker_gather, args=(world, 29751, out), nprocs=world, join=True)
    idx = torch.arange(1001, dtype=torch.float32)
    for r in range(world):
        full = torch.load(out + f".{r}")
        assert full["rgb"].shape == (2, 1001, 3) and torch.equal(full["rgb"][1, :, 2], 3 * idx)
        assert full["outside"].dtype == torch.bool and torch.equal(full["outside"][0, :, 0], idx.long() % 3 == 0)
        assert full["weights"].shape == (2, 1001, 4, 1) and torch.equal(full["weights"][0, :, 3, 0], idx)


def _worker_rs(rank, world, port_no, out):
    """reduce_scatter mode: after the exchange every rank holds the mean gradient of ITS shard of every slab (and the
    all-reduced small gradients); the dense table .grad is left alone."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = torch.nn.ModuleDict(dict(table=torch.nn.Embedding(1 << 19, 8), head=torch.nn.Linear(7, 3)))
    g = torch.Generator().manual_seed(300 + rank)
    n = model["table"].weight.numel()
    tg = torch.randn(n, generator=g)
    red = GradReducer(model, world, side_stream=False, table_mode="reduce_scatter")
    slabs = [(0, n // 4), (n // 4, n // 2), (n // 2, n)]

    class Eng:
        table_grad_hook = None
        device = "cpu"

        def n_table_params(self):
            return n

        def level_groups(self):
            return [(0, 0, a, b) for a, b in slabs]
    eng = Eng()
    red.attach(eng)
    for step in range(2):  # the shard list is rebuilt every step
        for a, b in slabs:
            eng.table_grad_hook(tg, a, b)
        model["table"].weight.grad = tg.view_as(model["table"].weight)
        model["head"].weight.grad = torch.randn(3, 7, generator=torch.Generator().manual_seed(400 + rank))
        red.exchange_grads()
    shards = [(a, b, t.clone()) for a, b, t in red._last_shards]
    torch.save(dict(shards=shards, head=model["head"].weight.grad, local=tg), out + f".{rank}")
    dist.destroy_process_group()


def test_grad_reduce_scatter_world2(tmp_path):
    world, out = 2, str(tmp_path / "rs")
    mp.spawn(_worker_rs, args=(world, 29761, out), nprocs=world, join=True)
    res = [torch.load(out + f".{r}") for r in range(world)]
    mean = sum(r["local"] for r in res) / world
    covered = torch.zeros_like(mean, dtype=torch.bool)
    for r in range(world):
        assert len(res[r]["shards"]) == 3
        for a, b, t in res[r]["shards"]:
            assert torch.allclose(t, mean[a:b], atol=1e-6)
            assert not bool(covered[a:b].any())
            covered[a:b] = True
        assert torch.equal(res[r]["head"], res[0]["head"])
    assert bool(covered.all())  # the ranks' shards partition the table
