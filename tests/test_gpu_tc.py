"""GPU tests of the tcgen05 (bf16 tensor-core) dense-layer kernels against a plain PyTorch fp32 reference of the same
op evaluated on bf16-rounded operands (so only accumulation order and the bf16 output rounding differ)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from mli_nerf_b200 import _lib
    _lib.load()
    assert torch.cuda.is_available() and _lib.device_ok()
    return _lib


def bf(x):
    return x.to(torch.bfloat16).float()


def to_tcl_host(x, tile_rows=128):
    """[M, C] fp32 -> bf16 TCL tensor (host-side definition of the layout)."""
    M, C = x.shape
    Mp = (M + tile_rows - 1) // tile_rows * tile_rows
    xp = torch.zeros(Mp, C)
    xp[:M] = x
    return xp.view(Mp // tile_rows, tile_rows, C // 8, 8).permute(0, 2, 1, 3).contiguous().to(torch.bfloat16)


def from_tcl_host(t, M):
    nt, nc, tr, _ = t.shape
    return t.float().permute(0, 2, 1, 3).reshape(nt * tr, nc * 8)[:M]


def test_tcl_roundtrip(lib):
    torch.manual_seed(0)
    for (M, C, tile) in ((300, 48, 128), (256, 144, 128), (768, 304, 256), (48, 768, 48)):
        x = torch.randn(M, C)
        ref = to_tcl_host(x, tile)
        dst = torch.empty_like(ref, device="cuda")
        lib.call("mli_tc_to_tcl", x.cuda(), C, M, C, dst, tile, C // 8, 0, C // 8)
        assert torch.equal(dst.cpu(), ref), (M, C, tile)
        if tile == 128:
            back = torch.empty(M, C, device="cuda")
            lib.call("mli_tc_from_tcl", dst, C // 8, 0, C // 8, M, back, C)
            assert torch.equal(back.cpu(), bf(x))


@pytest.mark.parametrize("M,N,K,BN,act,batch", [(128, 64, 16, 64, 0, 1), (128, 256, 64, 256, 0, 1), (1000, 256, 144, 256, 2, 1),
                                                (640, 768, 304, 256, 1, 1), (512, 256, 256, 256, 1, 3), (300, 48, 768, 48, 0, 1),
                                                (256, 128, 256, 128, 3, 1)])
def test_tc_linear_forward(lib, M, N, K, BN, act, batch):
    torch.manual_seed(1)
    X = torch.randn(M, batch * K) * 0.5
    W = torch.randn(batch, N, K) / math.sqrt(K)
    b = torch.randn(batch, N) * 0.1
    A = to_tcl_host(X).cuda()
    Wt = torch.stack([to_tcl_host(W[i], BN) for i in range(batch)]).cuda()
    f = {0: lambda v: v, 1: torch.relu, 2: lambda v: torch.nn.functional.softplus(v, beta=100), 3: torch.sigmoid}[act]
    ref = f(torch.einsum("mbk,bnk->mbn", bf(X).view(M, batch, K), bf(W)) + b)
    # bf16 TCL output
    out = torch.zeros((M + 127) // 128, batch * N // 8, 128, 8, dtype=torch.bfloat16, device="cuda")
    lib.call("mli_tc_linear", A, batch * K // 8, 0, K // 8, Wt, N * K, K, N, BN, b.cuda(), N, None, 0, 0, 0, act, out, 0,
             batch * N // 8, 0, N // 8, 0, 0, 0, M, batch, 0)
    got = from_tcl_host(out.cpu(), M).view(M, batch, N)
    assert torch.allclose(got, ref, rtol=1e-2, atol=1e-2), float((got - ref).abs().max())
    # fp32 row-major output: only fp32 accumulation-order differences remain
    out32 = torch.zeros(M, batch * N, device="cuda")
    lib.call("mli_tc_linear", A, batch * K // 8, 0, K // 8, Wt, N * K, K, N, BN, b.cuda(), N, None, 0, 0, 0, act, out32, 1,
             0, 0, 0, batch * N, 0, N, M, batch, 0)
    assert torch.allclose(out32.cpu().view(M, batch, N), ref, rtol=1e-4, atol=1e-4), float((out32.cpu().view(M, batch, N) - ref).abs().max())


def test_tc_linear_dgrad_epilogue(lib):
    torch.manual_seed(2)
    M, N_out, K_in = 384, 768, 256
    dZ, Wt = torch.randn(M, N_out), torch.randn(K_in, N_out) / 16   # Wt = W^T: [K_in, N_out]
    Yp = torch.randn(M, K_in).abs() * 0.01
    out = torch.zeros(3, K_in // 8, 128, 8, dtype=torch.bfloat16, device="cuda")
    lib.call("mli_tc_linear", to_tcl_host(dZ).cuda(), N_out // 8, 0, 0, to_tcl_host(Wt, 256).cuda(), 0, N_out, K_in, 256, None,
             0, to_tcl_host(Yp).cuda(), K_in // 8, 0, 0, 2, out, 0, K_in // 8, 0, 0, 0, 0, 0, M, 1, 1)
    y = bf(Yp)
    ref = (bf(dZ) @ bf(Wt).t()) * torch.where(y > 0.2, torch.ones_like(y), -torch.expm1(-100 * y))
    got = from_tcl_host(out.cpu(), M)
    assert torch.allclose(got, ref, rtol=1e-2, atol=2e-2), float((got - ref).abs().max())


@pytest.mark.parametrize("M,rows,cols,batch", [(128, 128, 16, 1), (1024, 256, 144, 1), (2048, 256, 256, 3), (640, 768, 48, 1),
                                               (4096, 256, 16, 1)])
def test_tc_wgrad(lib, M, rows, cols, batch):
    torch.manual_seed(3)
    dZ = torch.randn(M, batch * rows)
    X = torch.randn(M, batch * cols)
    out = torch.zeros(batch, rows, cols, device="cuda")
    ws = torch.empty(lib.load().mli_tc_wgrad_ws_bytes(M, rows, cols, batch), dtype=torch.uint8, device="cuda")
    lib.call("mli_tc_wgrad", to_tcl_host(dZ).cuda(), batch * rows // 8, 0, rows // 8, to_tcl_host(X).cuda(), batch * cols // 8,
             0, cols // 8, M, rows, cols, batch, out, cols, rows * cols, 0, ws)
    ref = torch.einsum("mbr,mbc->brc", bf(dZ).view(M, batch, rows), bf(X).view(M, batch, cols))
    err = float((out.cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-4, err
    outT = torch.zeros(batch, cols, rows, device="cuda")
    lib.call("mli_tc_wgrad", to_tcl_host(dZ).cuda(), batch * rows // 8, 0, rows // 8, to_tcl_host(X).cuda(), batch * cols // 8,
             0, cols // 8, M, rows, cols, batch, outT, rows, rows * cols, 1, ws)
    assert torch.equal(outT.cpu().transpose(1, 2), out.cpu())


def test_tc_colsum(lib):
    torch.manual_seed(4)
    M, C = 3000, 768
    X = torch.randn(M, C)
    out = torch.zeros(256, device="cuda")
    ws = torch.empty(lib.load().mli_tc_colsum_ws_bytes(M, 32), dtype=torch.uint8, device="cuda")
    lib.call("mli_tc_colsum", to_tcl_host(X).cuda(), C // 8, 32, 32, M, out, ws)
    ref = bf(X)[:, 256:512].sum(0)
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-3)
