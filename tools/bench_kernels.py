#!/usr/bin/env python
"""Micro-benchmarks of single C-ABI entry points at the bench workload's shapes (CUDA events, L2 flushed between
iterations by rotating through buffers larger than L2).  Usage on a B200: python tools/bench_kernels.py [name ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mli_nerf_b200 import _lib  # noqa: E402

M = 2048 * 128
dev = "cuda"


def tcl(rows, chunks, tile=128):
    return torch.randn((rows + tile - 1) // tile, chunks, tile, 8, device=dev).mul_(0.1).to(torch.bfloat16)


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3  # us


def bench_linear(K, N, BN, batch, epi, act, aux, out_f32=False, nbuf=3, label="", mask=False):
    A = [tcl(M, batch * K // 8) for _ in range(nbuf)]
    W = tcl(batch * N, K // 8, BN)
    bias = torch.zeros(batch * N, device=dev)
    AUX = [tcl(M, batch * N // 8).abs_() for _ in range(nbuf)] if aux else [None] * nbuf
    if out_f32:
        O = [torch.empty(M, batch * N, device=dev) for _ in range(nbuf)]
    else:
        O = [tcl(M, batch * N // 8) for _ in range(nbuf)]

    MK = [torch.randint(-2 ** 31, 2 ** 31 - 1, ((M + 127) // 128, batch * N // 32, 128), dtype=torch.int32, device=dev)
          for _ in range(nbuf)] if mask else [None] * nbuf

    def fn(i):
        j = i % nbuf
        _lib.call("mli_tc_linear", A[j], batch * K // 8, 0, K // 8, W, N * K, K, N, BN, bias if epi == 0 else None, N,
                  AUX[j], batch * N // 8 if aux else 0, 0, N // 8, act, O[j], int(out_f32), 0 if out_f32 else batch * N // 8,
                  0, N // 8, batch * N if out_f32 else 0, 0, N, M, batch, epi, MK[j], batch * N // 32 if mask else 0, 0,
                  N // 32 if mask else 0)
    us = timeit(fn)
    by = M * batch * (K * 2 + N * (4 if out_f32 else 2) + (N * 2 if aux else 0) + (N // 8 if mask else 0))
    fl = 2.0 * M * batch * K * N
    print(f"tc_linear {label:28s} K={K} N={N} BN={BN} b={batch} epi={epi}: {us:8.1f} us  {by / us / 1e6:6.2f} TB/s  "
          f"{fl / us / 1e6:7.1f} TFLOP/s  stages={os.environ.get('MLI_NT_STAGES', 'max')}")


def bench_wgrad(rows, cols, batch, with_db, nbuf=3, label=""):
    L = [tcl(M, batch * rows // 8) for _ in range(nbuf)]
    R = [tcl(M, batch * cols // 8) for _ in range(nbuf)]
    out = torch.empty(batch, rows, cols, device=dev)
    db = torch.empty(batch * rows, device=dev) if with_db else None
    ws = torch.empty(_lib.load().mli_tc_wgrad_ws_bytes(M, rows, cols, batch), dtype=torch.uint8, device=dev)

    def fn(i):
        j = i % nbuf
        _lib.call("mli_tc_wgrad", L[j], batch * rows // 8, 0, rows // 8, R[j], batch * cols // 8, 0, cols // 8, M, rows, cols,
                  batch, out, cols, rows * cols, 0, db, rows, ws)
    us = timeit(fn)
    by = M * batch * (rows + cols) * 2
    print(f"tc_wgrad  {label:28s} rows={rows} cols={cols} b={batch} db={with_db}: {us:8.1f} us  {by / us / 1e6:6.2f} TB/s  "
          f"{2.0 * M * batch * rows * cols / us / 1e6:7.1f} TFLOP/s")


def bench_encode(log2_T, taps=4, R=2048, n=128):
    import math
    grid = _lib.make_grid(16, 8, log2_T, 32, math.exp((math.log(2048) - math.log(32)) / 15))
    n_par = int(grid.n_entries) * 8
    table = (torch.rand(n_par, device=dev) * 2 - 1) * 1e-2
    g = torch.Generator(device="cpu").manual_seed(0)
    center = 3.0 * torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    ray = torch.nn.functional.normalize(0.3 * torch.randn(R, 3, generator=g) - center, dim=-1)
    dists = (2.0 + 2.0 * torch.rand(R, n, generator=g).sort(dim=1).values).to(dev)
    center, ray = center.to(dev), ray.to(dev)
    P, Mq = 1 + taps, R * n
    X = torch.empty(P * Mq // 128, 36, 128, 8, dtype=torch.bfloat16, device=dev)
    eps = 1.0 / 2048 / math.sqrt(3)
    us_f = timeit(lambda i: _lib.call("mli_encode_rays_tcl", grid, table, center, ray, dists, n, R, n, taps, eps, -2.0, 2.0,
                                      X, 36, 18))
    dX = tcl(P * Mq, 16)
    tg = torch.zeros(n_par, device=dev)
    us_b = timeit(lambda i: _lib.call("mli_encode_rays_bwd_tcl", grid, center, ray, dists, n, R, n, taps, eps, -2.0, 2.0,
                                      dX, 16, tg, 0, 16))
    print(f"encode T=2^{log2_T} ({n_par * 4 / 1e6:.0f} MB table): fwd {us_f:8.1f} us   bwd {us_b:8.1f} us")


def main():
    _lib.load()
    which = set(sys.argv[1:])
    if not which or "encode" in which:
        for t in (14, 17, 19, 22):
            bench_encode(t)
    if "mask" in which:
        for st in ("4", "5", "6"):
            os.environ["MLI_NT_STAGES"] = st
            bench_linear(256, 256, 256, 3, 0, 1, False, label="fwd relu (no mask)")
            bench_linear(256, 256, 256, 3, 0, 1, False, label="fwd relu + mask out", mask=True)
            bench_linear(256, 256, 256, 3, 1, 1, True, label="dgrad relu' from aux")
            bench_linear(256, 256, 256, 3, 1, 1, False, label="dgrad relu' from mask", mask=True)
        os.environ.pop("MLI_NT_STAGES", None)
    if not which or "linear" in which:
        for st in ("4", "6", None):
            if st is None:
                os.environ.pop("MLI_NT_STAGES", None)
            else:
                os.environ["MLI_NT_STAGES"] = st
            bench_linear(256, 256, 256, 3, 0, 1, False, label="head layer fwd (relu)")
            bench_linear(256, 256, 256, 3, 1, 1, True, label="head layer dgrad (relu')")
            bench_linear(256, 256, 128, 3, 0, 1, False, label="head layer fwd BN=128")
            bench_linear(256, 256, 128, 3, 1, 1, True, label="head layer dgrad BN=128")
        bench_linear(304, 768, 256, 1, 0, 1, False, label="head layer 0 fwd")
        bench_linear(304, 768, 128, 1, 0, 1, False, label="head layer 0 fwd BN=128")
        bench_linear(256, 256, 256, 1, 0, 2, False, label="sdf layer 1 fwd (softplus)")
        bench_linear(768, 256, 256, 1, 1, 2, True, label="feature dgrad (non-persist)")
        bench_linear(768, 256, 64, 1, 1, 2, True, label="feature dgrad BN=64")
    if not which or "wgrad" in which:
        bench_wgrad(256, 256, 3, True, label="head layer wgrad + db")
        bench_wgrad(256, 256, 3, False, label="head layer wgrad")
        bench_wgrad(768, 256, 1, True, label="head layer 0 wgrad")


if __name__ == "__main__":
    main()
