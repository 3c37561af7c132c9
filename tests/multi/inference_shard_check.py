"""Launched by tests/test_gpu_multi.py through torch.distributed.run on >= 2 GPUs: Model.inference with the frame's rays
partitioned over the ranks (BASELINE.json configs[4]) against the same frame rendered by one rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench  # noqa: E402
from mli_nerf_b200 import config  # noqa: E402
from mli_nerf_b200.model import Model  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
cfg = config.experiment("syn_hotdog_b", dict_size=16)
cfg.model.mli_precision = "bf16"
cfg.data.val.image_size = [37, 53]  # 1961 rays: ragged shards and ragged chunks
cfg.model.render.rand_rays_val = 300
torch.manual_seed(0)  # same weights on every rank
model = Model(cfg.model, cfg.data).cuda()
b = bench.synthetic_batch(8, 5)
view = dict(pose=b["pose"].cuda(), intr=torch.tensor([[[60.0, 0, 26.5], [0, 60.0, 18.5], [0, 0, 1]]]).cuda(),
            pose_light=b["pose_light"].cuda(), idx=torch.zeros(1, dtype=torch.long))
full = model.inference(view, per_sample=False)
shard = model.inference(view, per_sample=False, shard=True)
assert set(full) == set(shard)
for k in full:
    assert full[k].shape == shard[k].shape and full[k].dtype == shard[k].dtype, k
    # the same rays go through the same kernels in different launch shapes: identical up to the tile a ray lands in
    a, c = full[k].float(), shard[k].float()
    assert float((a - c).abs().max()) <= 1e-5 * (1.0 + float(a.abs().max())), (k, float((a - c).abs().max()))
if rank == 0:
    print(f"INFERENCE_SHARD_OK world={world}", flush=True)
dist.destroy_process_group()
