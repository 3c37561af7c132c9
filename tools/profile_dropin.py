#!/usr/bin/env python
"""tools/profile_dropin.py -- timing breakdown of the autograd drop-in path (model(data) -> torch losses -> backward()) at the bench workload."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from mli_nerf_b200 import _lib  # noqa: E402
from mli_nerf_b200.model import Model  # noqa: E402

cfg = bench.workload_cfg("syn_hotdog_b", "bf16")
torch.manual_seed(0)
model = Model(cfg.model, cfg.data).cuda().train()
model.progress = 0.5
dev = [{k: v.cuda() for k, v in bench.workload_batch("syn_hotdog_b", i, 0, 1).items()} for i in range(4)]


def dropin(i):
    for p in model.parameters():
        p.grad = None
    o = model(dev[i % len(dev)])
    bench.trainer_losses_torch(cfg.trainer, o, dev[i % len(dev)]).backward()


for i in range(3):
    dropin(i)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    for i in range(5):
        dropin(i)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"rep {rep}: host enqueue {1e3 * (t1 - t0) / 5:.2f} ms/step, total {1e3 * (t2 - t0) / 5:.2f} ms/step")
_lib.profile_begin(None)
for i in range(5):
    dropin(i)
prof = _lib.profile_end()
tot = 0.0
for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k:36s} {c / 5:5.1f} calls {1e3 * ms / 5:9.1f} us/step")
    tot += ms
print("sum of all entries", round(1e3 * sum(v[1] for v in prof.values()) / 5, 1), "us/step")
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as pr:
    for i in range(3):
        dropin(i)
    torch.cuda.synchronize()
print(pr.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=50))
print(torch.cuda.memory_summary(abbreviated=True)[:1800])
