#!/usr/bin/env python
"""Render (inference) throughput of the hot path: one 800x800 view (BASELINE configs[4] shape: 800x800 x 67 views are
640 000 rays each, chunked by rand_rays_val = 20 000 like the reference), eval outputs, bf16 mode.
Prints rays/s with and without materialising the per-sample debug tensors.    python tools/bench_render.py [--views 3]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mli_nerf_b200 import config  # noqa: E402
from mli_nerf_b200.model import Model  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=3)
    ap.add_argument("--size", type=int, default=800)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    cfg = config.experiment("syn_hotdog_b", dict_size=22)
    cfg.model.mli_precision = a.precision
    cfg.data.val.image_size = [a.size, a.size]
    torch.manual_seed(0)
    model = Model(cfg.model, cfg.data).cuda().eval()
    model.progress = 1.0
    out = {}
    for per_sample in (True, False):
        times = []
        for v in range(a.views + 1):
            b = bench.synthetic_batch(8, 50 + v, H=a.size, W=a.size)
            data = {k: b[k].cuda() for k in ("pose", "intr", "pose_light")}
            data["intr"][:, 0, 0] = data["intr"][:, 1, 1] = 711.0 * a.size / 512
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o = model.inference(data, per_sample=per_sample)
            e1.record()
            torch.cuda.synchronize()
            if v:  # first view = warm-up
                times.append(e0.elapsed_time(e1))
            del o
        ms = sum(times) / len(times)
        out["per_sample" if per_sample else "maps_only"] = {"ms_per_view": ms, "rays_per_s": a.size * a.size / ms * 1e3}
    print(json.dumps({"metric": "render_rays_per_sec", "image": [a.size, a.size], "chunk": model.rand_rays_val,
                      "precision": a.precision, **out}))


if __name__ == "__main__":
    main()
