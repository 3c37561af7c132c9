"""Launched by tests/test_gpu_multi.py through torch.distributed.run on >= 2 GPUs: gradients of the ray-sharded fused train
step after GradReducer (table gradient over NVLink peer memory, MLP gradients in an NCCL bucket) against the NCCL
all-reduce(mean) of the gradients every rank computes on its own."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench  # noqa: E402
from mli_nerf_b200 import config  # noqa: E402
from mli_nerf_b200.dist import GradReducer  # noqa: E402
from mli_nerf_b200.losses import loss_cfg_from_trainer  # noqa: E402
from mli_nerf_b200.model import Model  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
R = 256
cfg = config.experiment("syn_hotdog_b", dict_size=16, rand_rays=R)
cfg.model.mli_precision = "bf16"
cfg.model.render.stratified = False
torch.manual_seed(0)  # same weights on every rank
model = Model(cfg.model, cfg.data).cuda().train()
model.progress = 0.5
# The reference's geometric init zeroes the hash-encoding columns of SDF layer 0 (mlp.py:71-84): the table gradient would
# be exactly zero and every check of its exchange vacuous.  Give those columns (and the table) some weight, identically on
# every rank.
with torch.no_grad():
    g0 = torch.Generator(device="cuda").manual_seed(1234)
    w0 = model.neural_sdf.mlp.linears[0].weight_v
    w0[:, 3:] = torch.randn(w0[:, 3:].shape, device="cuda", generator=g0) * 0.3
    model.neural_sdf.mlp.linears[0].weight_g.copy_(w0.norm(dim=1, keepdim=True))
    tab0 = model.neural_sdf.tcnn_encoding.params
    tab0.copy_((torch.rand(tab0.shape, device="cuda", generator=g0) * 2 - 1) * 5e-3)
lcfg = loss_cfg_from_trainer(cfg.trainer)
data = {k: v.cuda() for k, v in bench.synthetic_batch(R, 100 + rank).items()}  # different rays per rank

model.fused_train_step(data, lcfg)
want = {}
for n, p in model.named_parameters():
    g = p.grad.detach().clone()
    dist.all_reduce(g, op=dist.ReduceOp.AVG)
    want[n] = g
    p.grad = None
assert float(want["neural_sdf.tcnn_encoding.params"].abs().max()) > 0, "table gradient is zero: the exchange checks would be vacuous"

table_mode = os.environ.get("MLI_TABLE_EXCHANGE", "allreduce")
TAB = "neural_sdf.tcnn_encoding.params"
for mode in (os.environ.get("MLI_TABLE_ALLREDUCE", "peer"),):
    reducer = GradReducer(model, world, table_mode=table_mode)
    reducer.attach(model.engine)
    for it in range(3):  # the persistent buffer is zeroed and refilled every step
        model.fused_train_step(data, lcfg, after_backward=reducer.exchange_grads)
    torch.cuda.synchronize()
    tab = dict(model.named_parameters())[TAB]
    for n, p in model.named_parameters():
        if n == TAB and table_mode == "reduce_scatter":
            continue  # only this rank's shards are reduced in this mode (checked below)
        err = float((p.grad - want[n]).norm() / (want[n].norm() + 1e-30))
        assert err < 2e-5, (mode, n, err)
    if table_mode == "allreduce":
        if reducer.peer is not None:
            assert tab.grad.data_ptr() == reducer.peer.buf.data_ptr()
        chk = tab.grad.double().sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(float(c) == float(allc[0]) for c in allc)  # bitwise identical replicas
    else:
        shards = reducer._last_shards
        assert len(shards) == len(model.engine.level_groups())
        wt = want[TAB].view(-1)
        own = 0
        for a, b, g in shards:
            err = float((g - wt[a:b]).norm() / (wt[a:b].norm() + 1e-30))
            assert err < 1e-4, (mode, a, b, err)  # float atomics: the scatter's summation order differs run to run
            own += b - a
        assert own * world == wt.numel()
        # rank-owned AdamW shard + parameter all-gather == dense AdamW on the mean gradient, identical replicas
        ref_p = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
        for q, (n, p) in zip(ref_p, model.named_parameters()):
            q.grad = want[n].clone().view_as(q)
        ref_opt = torch.optim.AdamW(ref_p, lr=1e-3, weight_decay=1e-2)
        opt = reducer.make_optimizer(lr=1e-3, weight_decay=1e-2)
        before = [p.detach().clone() for p in model.parameters()]
        opt.step()
        ref_opt.step()
        torch.cuda.synchronize()
        for q, p0, (n, p) in zip(ref_p, before, model.named_parameters()):
            # first AdamW step: the update is lr * g / (|g| + eps) -- where |g| ~ eps = 1e-8 the rounding noise of the
            # gradient (float atomics, reduction order) moves single entries by a fraction of lr, so the UPDATE VECTORS are
            # compared in relative L2 (and no entry may be off by more than the full step lr)
            du, dq = p.detach() - p0, q.detach() - p0
            err = float((du - dq).norm() / (dq.norm() + 1e-30))
            assert err < 1e-2 and float((du - dq).abs().max()) <= 2.1e-3, (n, err, float((du - dq).abs().max()))
        chk = tab.detach().double().sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(float(c) == float(allc[0]) for c in allc)  # bitwise identical parameter replicas
        pre = [(k, float(st[0].abs().max()), float(st[1].max())) for k, st in sorted(opt.state.items())]
        m, v = opt.gather_state()
        if rank == 0:
            print("moments per owned shard (range, max |m|, max v):", pre, flush=True)
        stats = (tuple(m.shape), tuple(tab.shape), float(m.abs().max()), float(v.min()), float(v.max()),
                 int(torch.isnan(m).sum()), int(torch.isnan(v).sum()))
        assert m.shape == tab.shape and stats[2] > 0 and stats[3] >= 0 and stats[5] == 0 and stats[6] == 0, stats
        # dense moments == torch's after the same single step: exp_avg = 0.1 g, exp_avg_sq = 0.001 g^2
        gm = want[TAB].view(-1)
        assert float((m.view(-1) - 0.1 * gm).norm()) <= 1e-4 * float((0.1 * gm).norm()), "gathered exp_avg"
    reducer.close()
if rank == 0:
    print(f"TRAIN_STEP_EXCHANGE_OK world={world} table={table_mode}", flush=True)
dist.destroy_process_group()
