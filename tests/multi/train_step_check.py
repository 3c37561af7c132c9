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
lcfg = loss_cfg_from_trainer(cfg.trainer)
data = {k: v.cuda() for k, v in bench.synthetic_batch(R, 100 + rank).items()}  # different rays per rank

model.fused_train_step(data, lcfg)
want = {}
for n, p in model.named_parameters():
    g = p.grad.detach().clone()
    dist.all_reduce(g, op=dist.ReduceOp.AVG)
    want[n] = g
    p.grad = None

for mode in (os.environ.get("MLI_TABLE_ALLREDUCE", "peer"),):
    reducer = GradReducer(model, world)
    reducer.attach(model.engine)
    for it in range(3):  # the persistent buffer is zeroed and refilled every step
        model.fused_train_step(data, lcfg, after_backward=reducer.allreduce_grads)
    torch.cuda.synchronize()
    for n, p in model.named_parameters():
        err = float((p.grad - want[n]).norm() / (want[n].norm() + 1e-30))
        assert err < 2e-5, (mode, n, err)
    tab = dict(model.named_parameters())["neural_sdf.tcnn_encoding.params"]
    if reducer.peer is not None:
        assert tab.grad.data_ptr() == reducer.peer.buf.data_ptr()
    chk = tab.grad.double().sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    assert all(float(c) == float(allc[0]) for c in allc)  # bitwise identical replicas
    reducer.close()
if rank == 0:
    print(f"TRAIN_STEP_ALLREDUCE_OK world={world}", flush=True)
dist.destroy_process_group()
