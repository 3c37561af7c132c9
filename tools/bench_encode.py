#!/usr/bin/env python
"""Micro-benchmark of the hash-grid kernels on the bench workload's REAL sample distribution: one eager training step
of `bench.py`'s workload is run, the arguments of its stencil encode (`mli_encode_rays_tcl`, taps > 0) and of its table
scatter (`mli_encode_rays_bwd_tcl`) are captured, and those two calls are replayed under CUDA events -- the whole call,
and the scatter level by level.  Both forward kernels (`MLI_ENCODE_VARIANT` = 1 thread per plane, 2 corner
caching) are timed and their outputs compared bit for bit.

    python tools/bench_encode.py [--workload syn_hotdog_b] [--iters 10] [--levels]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from mli_nerf_b200 import _lib  # noqa: E402
from mli_nerf_b200._lib import call  # noqa: E402
from mli_nerf_b200.losses import loss_cfg_from_trainer  # noqa: E402
from mli_nerf_b200.model import Model  # noqa: E402


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="syn_hotdog_b")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--levels", action="store_true", help="time the scatter level by level")
    args = ap.parse_args()
    cfg = bench.workload_cfg(args.workload, "bf16")
    torch.manual_seed(0)
    model = Model(cfg.model, cfg.data).cuda().train()
    model.progress = 0.5
    lcfg = loss_cfg_from_trainer(cfg.trainer)
    batch = {k: v.cuda() for k, v in bench.workload_batch(args.workload, 0, 0, 1).items()}
    for _ in range(2):
        model.fused_train_step(batch, lcfg, use_graph=False)
    captured = {}

    def grab(name, a):
        if name == "mli_encode_rays_tcl" and a[8] > 0:
            captured["fwd"] = list(a)
        if name == "mli_encode_rays_bwd_tcl":
            captured["bwd"] = list(a)
        return None

    _lib.profile_begin(grab)
    model.fused_train_step(batch, lcfg, use_graph=False)
    _lib.profile_end()
    f, b = captured["fwd"], captured["bwd"]
    # the engine's buffers are pooled: keep private copies of everything the replay reads or writes
    f = [x.clone() if isinstance(x, torch.Tensor) and i != 1 else x for i, x in enumerate(f)]
    b = [x.clone() if isinstance(x, torch.Tensor) else x for x in b]
    R, n, taps = f[6], f[7], f[8]
    print(f"workload {args.workload}: R={R} n={n} taps={taps} tap_eps={f[9]:.3e}")
    out_bytes = f[12].numel() * 2
    outs = {}
    for variant in ("1", "2"):
        os.environ["MLI_ENCODE_VARIANT"] = variant
        f[12].zero_()
        us = timed(lambda: call("mli_encode_rays_tcl", *f), args.iters)
        outs[variant] = f[12].clone()
        print(f"encode fwd (stencil launch) variant {variant}: {us:8.1f} us   ({out_bytes / 1e6:.0f} MB written -> "
              f"{out_bytes / us / 1e3:.0f} GB/s of output alone)")
    same = torch.equal(outs["1"].view(torch.int16), outs["2"].view(torch.int16))
    print("variants bit-exact:", same)
    if not same:
        d = (outs["1"].float() - outs["2"].float()).abs()
        print("  max abs diff", float(d.max()), "differing elements", int((d != 0).sum()))
    b[11].copy_((torch.randn(b[11].shape, device="cuda") * 1e-3).to(b[11].dtype))  # geometric init: the real dX is exactly 0
    tgs = {}
    for variant in ("1", "2"):
        os.environ["MLI_ENCODE_VARIANT"] = variant
        us_all = timed(lambda: call("mli_encode_rays_bwd_tcl", *b), args.iters)
        b[13].zero_()
        call("mli_encode_rays_bwd_tcl", *b)
        tgs[variant] = b[13].clone()
        print(f"scatter bwd (all levels) variant {variant}: {us_all:8.1f} us")
    scale = float(tgs["1"].abs().max())
    err = float((tgs["1"] - tgs["2"]).abs().max())
    print(f"scatter variant 2 vs 1: max abs diff {err:.3e} at scale {scale:.3e}")
    del tgs
    if args.levels:
        L = b[15]
        tot = 0.0
        for lv in range(L):
            bb = list(b)
            bb[14], bb[15] = lv, lv + 1
            us = timed(lambda: call("mli_encode_rays_bwd_tcl", *bb), args.iters)
            tot += us
            lvl = model.engine.grid.level[lv]
            print(f"  level {lv:2d} res {lvl.res:5d} size {lvl.size:8d} hashed {lvl.hashed}: {us:7.1f} us")
        print(f"  sum of single-level launches {tot:.1f} us")


if __name__ == "__main__":
    main()
