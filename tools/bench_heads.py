#!/usr/bin/env python
"""Micro-benchmark of the fused head-stack kernels (csrc/heads_fused.cu) at the bench workload's shape
(M = 2048 rays x 128 samples, three heads, random operands): CUDA-event time per launch and TFLOP/s of the executed
shapes, next to the layer-by-layer tensor-core launches they replace.  Short enough to point `ncu --set full` at:

    python tools/bench_heads.py [--rays 2048] [--iters 10] [--only fwd|fwd_nostore|bwd|layered_fwd|layered_bwd]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mli_nerf_b200 import _lib, config  # noqa: E402
from mli_nerf_b200._lib import ACT_RELU, ACT_SIGMOID, call  # noqa: E402
from mli_nerf_b200.engine import HID, KH_PAD  # noqa: E402
from mli_nerf_b200.model import Model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.manual_seed(0)
    cfg = config.experiment("syn_hotdog_b", dict_size=14, rand_rays=args.rays)
    cfg.model.mli_precision = "bf16"
    model = Model(cfg.model, cfg.data).cuda()
    eng = model.engine
    eng.fuse_heads = True
    p = dict(model.named_parameters())
    with torch.no_grad():
        W = eng.pack_weights(p)
    T, nh = W["T"], eng.nh
    M = args.rays * 128
    XH = (torch.randn(M // 128, KH_PAD // 8, 128, 8, device="cuda") * 0.5).to(torch.bfloat16)
    A = [eng._tcl(M, nh * 32) for _ in range(4)]
    Am = [eng._mask(M, nh * HID) for _ in range(4)]
    dZs = [eng._tcl(M, nh * 32) for _ in range(4)]
    S = eng._f(M, eng.lds)
    dS = torch.randn(M, eng.lds, device="cuda") * 1e-3
    j0s, njs, j = [], [], 0
    for h in eng.heads:
        j0s.append(j)
        njs.append(h[2])
        j += h[2]

    def fwd(store):
        call("mli_tc_heads_fwd", XH, M, nh, KH_PAD, int(store), KH_PAD // 8, T["Wh0"], T["Whl128"][0], T["Whl128"][1],
             T["Whl128"][2], W["bh"][0], W["bh"][1], W["bh"][2], W["bh"][3], W["Wout"], W["bout"], j0s, njs, ACT_SIGMOID,
             eng.act_mask, A[0] if store else None, A[1] if store else None, A[2] if store else None,
             A[3] if store else None, Am[0] if store else None, Am[1] if store else None, Am[2] if store else None,
             Am[3] if store else None, S, eng.lds)

    def bwd():
        call("mli_tc_heads_bwd", dS, eng.lds, M, nh, T["Whlt128"][2], T["Whlt128"][1], T["Whlt128"][0], W["Wout"], j0s, njs,
             Am[0], Am[1], Am[2], Am[3], dZs[0], dZs[1], dZs[2], dZs[3])

    def layered_fwd():
        eng._tc_linear(XH, 0, 0, T["Wh0"], 0, KH_PAD, nh * HID, 256, W["bh"][0], 0, None, 0, 0, ACT_RELU, A[0], False, 0, 0, 0,
                       M, 1, 0, mask=Am[0])
        for l in range(2):
            eng._tc_linear(A[l], 0, 32, T["Whl"][l], HID * HID, HID, HID, 256, W["bh"][l + 1], HID, None, 0, 0, ACT_RELU,
                           A[l + 1], False, 0, 32, 0, M, nh, 0, mask=Am[l + 1], mask_bchunks=8)
        call("mli_tc_linear_dot", A[2], nh * 32, 0, 32, T["Whl"][2], HID * HID, HID, W["bh"][3], HID, A[3], nh * 32, 0, 32, M,
             nh, W["Wout"], W["bout"], j0s, njs, ACT_SIGMOID, eng.act_mask, S, eng.lds, Am[3], Am[3].shape[1], 0, 8)

    def layered_bwd():
        dZ = dZs[3]
        call("mli_tc_rowdot_bwd_data", dS, eng.lds, A[3], nh * 32, M, W["Wout"], eng.col_off, eng.J, HID, ACT_RELU, dZ, Am[3])
        for l in (2, 1, 0):
            eng._tc_linear(dZ, 0, 32, T["Whlt"][l], HID * HID, HID, HID, 256, None, 0, None, 0, 32, ACT_RELU, dZs[l], False, 0,
                           32, 0, M, nh, 1, mask=Am[l], mask_bchunks=8)
            dZ = dZs[l]

    fl_f = 2.0 * M * nh * HID * (KH_PAD + 3 * HID)
    fl_b = 2.0 * M * nh * HID * 3 * HID
    cases = [("fwd", lambda: fwd(True), fl_f), ("fwd_nostore", lambda: fwd(False), fl_f), ("bwd", bwd, fl_b),
             ("layered_fwd", layered_fwd, fl_f), ("layered_bwd", layered_bwd, fl_b)]
    for name, fn, flops in cases:
        if args.only and name != args.only:
            continue
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / args.iters
        print(f"{name:12s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s (executed shapes, M = {M})", flush=True)


if __name__ == "__main__":
    main()
