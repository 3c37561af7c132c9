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
             batch * N // 8, 0, N // 8, 0, 0, 0, M, batch, 0, None, 0, 0, 0)
    got = from_tcl_host(out.cpu(), M).view(M, batch, N)
    assert torch.allclose(got, ref, rtol=1e-2, atol=1e-2), float((got - ref).abs().max())
    # fp32 row-major output: only fp32 accumulation-order differences remain
    out32 = torch.zeros(M, batch * N, device="cuda")
    lib.call("mli_tc_linear", A, batch * K // 8, 0, K // 8, Wt, N * K, K, N, BN, b.cuda(), N, None, 0, 0, 0, act, out32, 1,
             0, 0, 0, batch * N, 0, N, M, batch, 0, None, 0, 0, 0)
    assert torch.allclose(out32.cpu().view(M, batch, N), ref, rtol=1e-4, atol=1e-4), float((out32.cpu().view(M, batch, N) - ref).abs().max())


def test_tc_linear_dgrad_epilogue(lib):
    torch.manual_seed(2)
    M, N_out, K_in = 384, 768, 256
    dZ, Wt = torch.randn(M, N_out), torch.randn(K_in, N_out) / 16   # Wt = W^T: [K_in, N_out]
    Yp = torch.randn(M, K_in).abs() * 0.01
    out = torch.zeros(3, K_in // 8, 128, 8, dtype=torch.bfloat16, device="cuda")
    lib.call("mli_tc_linear", to_tcl_host(dZ).cuda(), N_out // 8, 0, 0, to_tcl_host(Wt, 256).cuda(), 0, N_out, K_in, 256, None,
             0, to_tcl_host(Yp).cuda(), K_in // 8, 0, 0, 2, out, 0, K_in // 8, 0, 0, 0, 0, 0, M, 1, 1, None, 0, 0, 0)
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
             0, cols // 8, M, rows, cols, batch, out, cols, rows * cols, 0, None, 0, ws)
    ref = torch.einsum("mbr,mbc->brc", bf(dZ).view(M, batch, rows), bf(X).view(M, batch, cols))
    err = float((out.cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-4, err
    outT = torch.zeros(batch, cols, rows, device="cuda")
    db = torch.zeros(batch * rows, device="cuda")
    lib.call("mli_tc_wgrad", to_tcl_host(dZ).cuda(), batch * rows // 8, 0, rows // 8, to_tcl_host(X).cuda(), batch * cols // 8,
             0, cols // 8, M, rows, cols, batch, outT, rows, rows * cols, 1, db, rows, ws)
    assert torch.equal(outT.cpu().transpose(1, 2), out.cpu())
    # fused column sums of L (bias gradient) from the same pass
    ref_db = bf(dZ).view(M, batch, rows).sum(0)
    assert float((db.cpu().view(batch, rows) - ref_db).abs().max()) < 1e-4 * float(ref_db.abs().max()) + 1e-3


def test_tc_colsum(lib):
    torch.manual_seed(4)
    M, C = 3000, 768
    X = torch.randn(M, C)
    out = torch.zeros(256, device="cuda")
    ws = torch.empty(lib.load().mli_tc_colsum_ws_bytes(M, 32), dtype=torch.uint8, device="cuda")
    lib.call("mli_tc_colsum", to_tcl_host(X).cuda(), C // 8, 32, 32, M, out, ws)
    ref = bf(X)[:, 256:512].sum(0)
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-3)


# ---------------------------------------------------------------------------------------------------------------
# SDF trunk on the tensor cores: split-bf16 operands, delta basis (gemm_tcgen05.cu EPI_SDF_*, sdf_trunk.cu, hashgrid.cu)
# ---------------------------------------------------------------------------------------------------------------
def split_tcl_host(x, tile_rows=128):
    """[M, C] fp32 -> bf16 TCL with 2*C/8 chunks: [bf16(x) | bf16(x - bf16(x))]."""
    hi = bf(x)
    return torch.cat([to_tcl_host(hi, tile_rows), to_tcl_host(x - hi, tile_rows)], dim=1)


def tcl32_host_to_rows(t, M):
    """fp32 TCL32 [tiles][C/4][128][4] -> [M, C]."""
    nt, nc, tr, _ = t.shape
    return t.permute(0, 2, 1, 3).reshape(nt * tr, nc * 4)[:M]


def rows_to_tcl32_host(x):
    M, C = x.shape
    return x.view(M // 128, 128, C // 4, 4).permute(0, 2, 1, 3).contiguous()


def test_to_tcl_split(lib):
    torch.manual_seed(5)
    x = torch.randn(256, 144)
    dst = torch.zeros(1, 36, 256, 8, dtype=torch.bfloat16, device="cuda")
    lib.call("mli_tc_to_tcl_split", x.cuda(), 144, 256, 144, dst, 256, 36, 0, 18, 18)
    assert torch.equal(dst.cpu(), split_tcl_host(x, 256))
    rec = from_tcl_host(dst.cpu()[:, :18], 256) + from_tcl_host(dst.cpu()[:, 18:], 256)
    assert float((rec - x).abs().max()) < 2e-5 * float(x.abs().max())


def _trunk_case(M, seed=6):
    torch.manual_seed(seed)
    K = 144
    W0 = torch.randn(256, K) * 0.15
    b0 = torch.randn(256) * 0.05
    w_sdf = torch.randn(256) * 0.1 + 0.1
    b_sdf = torch.tensor([-0.5])
    x0 = torch.randn(M, K) * 0.3
    x0[:, 131:] = 0
    return K, W0, b0, w_sdf, b_sdf, x0


def _softplus100(z):
    return torch.nn.functional.softplus(z, beta=100)


def test_sdf_trunk_fwd_center_and_sdf_only(lib):
    M = 640
    K, W0, b0, w_sdf, b_sdf, x0 = _trunk_case(M)
    Xs, Ws = split_tcl_host(x0).cuda(), split_tcl_host(W0, 256).cuda()
    s0 = torch.zeros(M // 128, 64, 128, 4, device="cuda")
    h0 = torch.zeros(M // 128, 32, 128, 8, dtype=torch.bfloat16, device="cuda")
    sdf = torch.zeros(M, device="cuda")
    lib.call("mli_tc_sdf_trunk_fwd", Xs, 36, K, Ws, b0.cuda(), w_sdf.cuda(), b_sdf.cuda(), M, 0, M, s0, h0, sdf)
    z = x0.double() @ W0.double().t() + b0.double()
    h = _softplus100(z)
    ref_sdf = h @ w_sdf.double() + b_sdf.double()
    # split-bf16 (3 products) carries ~16 mantissa bits per operand: |z| ~ 1 -> errors of a few 1e-5
    assert float((sdf.cpu().double() - ref_sdf).abs().max()) < 2e-4, float((sdf.cpu().double() - ref_sdf).abs().max())
    got_s0 = tcl32_host_to_rows(s0.cpu(), M).double()
    assert float((got_s0 - torch.sigmoid(100 * z)).abs().max()) < 5e-3   # sigmoid(100 z): slope 25 x |dz| ~ 1e-4
    got_h = from_tcl_host(h0.cpu(), M).double()
    assert torch.allclose(got_h, h, rtol=1e-2, atol=1e-3)
    sdf2 = torch.zeros(M, device="cuda")
    lib.call("mli_tc_sdf_trunk_fwd", Xs, 36, K, Ws, b0.cuda(), w_sdf.cuda(), b_sdf.cuda(), M, 2, 0, None, None, sdf2)
    assert torch.equal(sdf2.cpu(), sdf.cpu())
    # ragged row count (sampling queries): rows beyond M are not written
    Mr = 300
    sdf3 = torch.full((384,), 7.0, device="cuda")
    lib.call("mli_tc_sdf_trunk_fwd", Xs, 36, K, Ws, b0.cuda(), w_sdf.cuda(), b_sdf.cuda(), Mr, 2, 0, None, None, sdf3)
    assert torch.equal(sdf3.cpu()[:Mr], sdf.cpu()[:Mr]) and bool((sdf3.cpu()[Mr:] == 7.0).all())


def test_sdf_trunk_fwd_taps_delta(lib):
    """d_i = sdf(x0 + dx_i) - sdf(x0) from deltas: must match an fp64 evaluation of the ABSOLUTE formula to ~1e-5
    relative -- the property the 4-tap stencil needs and bf16 *values* cannot give."""
    M, taps = 256, 4
    K, W0, b0, w_sdf, b_sdf, x0 = _trunk_case(M, seed=7)
    dx = torch.randn(taps, M, K) * 3e-4
    dx[:, :, 131:] = 0
    Ws = split_tcl_host(W0, 256).cuda()
    X = torch.cat([split_tcl_host(x0)] + [split_tcl_host(dx[i]) for i in range(taps)], dim=0).cuda()
    s0 = torch.zeros(M // 128, 64, 128, 4, device="cuda")
    h0 = torch.zeros(M // 128, 32, 128, 8, dtype=torch.bfloat16, device="cuda")
    dz = torch.zeros(taps * M // 128, 32, 128, 8, dtype=torch.bfloat16, device="cuda")
    sdf = torch.zeros((1 + taps) * M, device="cuda")
    args = (36, K, Ws, b0.cuda(), w_sdf.cuda(), b_sdf.cuda())
    lib.call("mli_tc_sdf_trunk_fwd", X, *args, M, 0, M, s0, h0, sdf)
    lib.call("mli_tc_sdf_trunk_fwd", X[M // 128:], *args, taps * M, 1, M, s0, dz, sdf[M:])
    # reference on the operands the kernel actually sees (hi + lo), in float64, absolute formulation
    def rec(t):
        return (bf(t) + bf(t - bf(t))).double()
    Wd = rec(W0)
    z0 = rec(x0) @ Wd.t() + b0.double()
    f0 = _softplus100(z0) @ w_sdf.double()
    got = sdf.cpu().double()[M:].view(taps, M)
    for i in range(taps):
        zi = z0 + rec(dx[i]) @ Wd.t()
        ref = _softplus100(zi) @ w_sdf.double() - f0
        err = float((got[i] - ref).abs().max() / ref.abs().max())
        assert err < 2e-4, (i, err)
        dzi = from_tcl_host(dz.cpu()[i * M // 128:(i + 1) * M // 128], M).double()
        assert torch.allclose(dzi, zi - z0, rtol=1e-2, atol=1e-5)
    # fused single-launch variant (sigma0 resident in TMEM, tap accumulators in two 128-column halves): same numbers
    s0f, h0f, dzf = torch.zeros_like(s0), torch.zeros_like(h0), torch.zeros_like(dz)
    sdff = torch.zeros_like(sdf)
    lib.call("mli_tc_sdf_trunk_fused", X, *args, M, taps, s0f, h0f, dzf, sdff)
    assert torch.equal(s0f.cpu(), s0.cpu()) and torch.equal(h0f.cpu(), h0.cpu()) and torch.equal(dzf.cpu(), dz.cpu())
    assert torch.allclose(sdff.cpu(), sdf.cpu(), rtol=1e-6, atol=1e-9)
    sdfe = torch.zeros_like(sdf)   # eval flavour: no sigma0 / dz stores
    lib.call("mli_tc_sdf_trunk_fused", X, *args, M, taps, None, h0f, None, sdfe)
    assert torch.equal(sdfe.cpu(), sdff.cpu())
    # second-order quantity (what the Hessian uses): sum of the 4 deltas of a symmetric stencil
    dxs = torch.randn(M, K) * 3e-4
    dxs[:, 131:] = 0
    sym = torch.stack([dxs, -dxs, dxs.flip(1) * 0 + dxs * 0.5, -dxs * 0.5])
    X2 = torch.cat([split_tcl_host(x0)] + [split_tcl_host(sym[i]) for i in range(taps)], dim=0).cuda()
    lib.call("mli_tc_sdf_trunk_fwd", X2[M // 128:], *args, taps * M, 1, M, s0, None, sdf[M:])
    got2 = sdf.cpu().double()[M:].view(taps, M).sum(0)
    ref2 = sum(_softplus100(z0 + rec(sym[i]) @ Wd.t()) @ w_sdf.double() - f0 for i in range(taps))
    assert float((got2 - ref2).abs().max()) < 2e-2 * float(ref2.abs().max()) + 1e-9, (float((got2 - ref2).abs().max()), float(ref2.abs().max()))


def test_sdf_trunk_bwd(lib):
    torch.manual_seed(8)
    M, taps = 256, 4
    z0 = torch.randn(M, 256) * 0.02
    z0[:, :16] += 0.3      # saturated units (softplus threshold)
    z0[:, 16:32] -= 0.3    # dead units
    s0 = torch.sigmoid(100 * z0)
    dzs = torch.randn(taps, M, 256) * 2e-3
    g = torch.randn(1 + taps, M)
    g[1:] *= 900.0
    dH0 = torch.randn(M, 256)
    h0 = _softplus100(z0)
    w = torch.randn(256) * 0.1
    Ed = torch.zeros((1 + taps) * M // 128, 32, 128, 8, dtype=torch.bfloat16, device="cuda")
    dw, db = torch.zeros(256, device="cuda"), torch.zeros(1, device="cuda")
    ws = torch.empty(lib.load().mli_tc_sdf_trunk_bwd_ws_bytes(M), dtype=torch.uint8, device="cuda")
    dz_t = torch.cat([to_tcl_host(dzs[i]) for i in range(taps)], dim=0).cuda()
    lib.call("mli_tc_sdf_trunk_bwd", g.reshape(-1).cuda(), M, taps, rows_to_tcl32_host(s0).cuda(), dz_t,
             to_tcl_host(dH0).cuda(), to_tcl_host(h0).cuda(), w.cuda(), Ed, dw, db, ws)
    # reference on the rounded inputs the kernel sees
    dzr, dHr, hr = bf(dzs).double(), bf(dH0).double(), bf(h0).double()
    s0d, gd, wd = s0.double(), g.double(), w.double()
    e0 = (gd[0][:, None] * wd + dHr) * s0d
    a = torch.expm1(100 * dzr)
    si = (1 + a) * s0d / (1 + a * s0d)
    ei = gd[1:, :, None] * wd * si
    E = e0 + ei.sum(0)
    got = from_tcl_host(Ed.cpu(), (1 + taps) * M).double().view(1 + taps, M, 256)
    assert float((got[0] - E).abs().max()) < 1e-2 * float(E.abs().max()), float((got[0] - E).abs().max())
    for i in range(taps):
        assert float((got[1 + i] - ei[i]).abs().max()) < 1e-2 * float(ei[i].abs().max())
    dhi = torch.log1p(a * s0d) / 100
    ref_dw = (gd.sum(0)[:, None] * hr).sum(0) + (gd[1:, :, None] * dhi).sum((0, 1))
    assert float((dw.cpu().double() - ref_dw).abs().max()) < 1e-4 * float(ref_dw.abs().max()) + 1e-3
    assert abs(float(db.cpu()) - float(gd.sum())) < 1e-3 * float(gd.abs().sum()) * 1e-3 + 1e-2
    # heads-only variant (no dw): same Ed
    Ed2 = torch.zeros_like(Ed)
    lib.call("mli_tc_sdf_trunk_bwd", g.reshape(-1).cuda(), M, taps, rows_to_tcl32_host(s0).cuda(), dz_t,
             to_tcl_host(dH0).cuda(), None, w.cuda(), Ed2, None, None, None)
    # (different instantiation -> different FMA contraction: equal to within one bf16 ulp)
    assert torch.allclose(Ed2.cpu().float(), Ed.cpu().float(), rtol=1e-2, atol=1e-2)


def _grid(lib, T=14):
    return lib.make_grid(16, 8, T, 32, math.exp((math.log(2048) - math.log(32)) / 15))


def _rays(R, seed=9):
    g = torch.Generator().manual_seed(seed)
    center = 3.0 * torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    ray = torch.nn.functional.normalize(0.3 * torch.randn(R, 3, generator=g) - center, dim=-1)
    dists = 2.0 + 2.0 * torch.rand(R, 16, generator=g).sort(dim=1).values
    return center, ray, dists


@pytest.mark.parametrize("taps", [0, 4, 6])
def test_encode_rays_tcl_matches_fp32_encode(lib, taps):
    R, n = 64, 16
    grid = _grid(lib)
    torch.manual_seed(10)
    table = ((torch.rand(int(grid.n_entries) * 8) * 2 - 1) * 0.05).cuda()
    center, ray, dists = _rays(R)
    eps = 1.0 / 2048 / (math.sqrt(3) if taps == 4 else 1.0)
    M, P = R * n, 1 + taps
    X32 = torch.zeros(P * M, 144, device="cuda")
    lib.call("mli_encode_rays", grid, table, center.cuda(), ray.cuda(), dists.cuda(), 16, R, n, taps, eps, -2.0, 2.0, X32, 144)
    Xt = torch.zeros(P * M // 128, 36, 128, 8, dtype=torch.bfloat16, device="cuda")
    lib.call("mli_encode_rays_tcl", grid, table, center.cuda(), ray.cuda(), dists.cuda(), 16, R, n, taps, eps, -2.0, 2.0,
             Xt, 36, 18)
    rec = (from_tcl_host(Xt.cpu()[:, :18], P * M) + from_tcl_host(Xt.cpu()[:, 18:], P * M)).view(P, M, 144)
    ref = X32.cpu().view(P, M, 144)
    assert float((rec[0] - ref[0]).abs().max()) < 2e-5 * float(ref[0].abs().max())
    assert bool((rec[:, :, 131:] == 0).all())
    for p in range(1, P):
        d = ref[p] - ref[0]   # fp32 difference of fp32 rows == what the kernel forms before splitting
        assert float((rec[p] - d).abs().max()) < 2e-5 * float(d.abs().max()) + 1e-12, p


@pytest.mark.parametrize("taps,eps_cells", [(4, 0.15), (4, 1.0), (4, 6.0), (6, 0.5)])
def test_encode_rays_tcl_corner_caching_variant_is_bit_exact(lib, taps, eps_cells, monkeypatch):
    """The default stencil encode (corner values cached in registers, taps outside the centre's cell deferred) must
    produce the very bits of the thread-per-plane kernel: tiny eps -> nearly all taps share the cell, eps of several fine
    cells -> nearly all are deferred; an inactive level range and a partial last CTA are included."""
    R, n = 72, 16     # 1152 samples = 9 tiles of 128
    grid = _grid(lib)
    grid.active_levels = 13
    torch.manual_seed(13)
    table = ((torch.rand(int(grid.n_entries) * 8) * 2 - 1) * 0.05).cuda()
    center, ray, dists = _rays(R, seed=14)
    eps = eps_cells / 2048
    P = 1 + taps
    outs = []
    for variant in ("1", "2"):
        monkeypatch.setenv("MLI_ENCODE_VARIANT", variant)
        Xt = torch.full((P * R * n // 128, 36, 128, 8), 7.0, dtype=torch.bfloat16, device="cuda")
        lib.call("mli_encode_rays_tcl", grid, table, center.cuda(), ray.cuda(), dists.cuda(), 16, R, n, taps, eps, -2.0,
                 2.0, Xt, 36, 18)
        outs.append(Xt.cpu().view(torch.int16))
    differing = int((outs[0] != outs[1]).sum())
    assert differing == 0, differing
    nonzero = int((outs[1][:, :13] != 0).sum())
    assert nonzero > 0


@pytest.mark.parametrize("taps,eps_cells", [(4, 0.15), (4, 1.0), (4, 6.0), (6, 0.5)])
def test_encode_rays_bwd_tcl_v2_equals_v1(lib, taps, eps_cells, monkeypatch):
    """Scatter kernel v2 (taps add into the centre cell's accumulators wherever they share a lattice point with it, the
    rest is scattered in warp-compacted rounds) against the first-generation kernel: same table gradient up to the
    summation order, for taps inside the centre's cell, one cell over and several cells away."""
    R, n = 72, 16
    grid = _grid(lib)
    grid.active_levels = 14
    center, ray, dists = _rays(R, seed=15)
    eps = eps_cells / 2048
    M, P = R * n, 1 + taps
    torch.manual_seed(16)
    dX = torch.randn(P, M, 128)
    dX[1:] *= 30.0
    dX[0] += dX[1:].sum(0)
    dXt = to_tcl_host(dX.reshape(P * M, 128)).cuda()
    n_par = int(grid.n_entries) * 8
    args = (grid, center.cuda(), ray.cuda(), dists.cuda(), 16, R, n, taps, eps, -2.0, 2.0)
    out = []
    for variant in ("1", "2"):
        monkeypatch.setenv("MLI_ENCODE_VARIANT", variant)
        tg = torch.zeros(n_par, device="cuda")
        for lv0, lv1 in ((0, 6), (6, 16)):
            lib.call("mli_encode_rays_bwd_tcl", *args, dXt, 16, tg, lv0, lv1)
        out.append(tg.cpu().double())
    scale = float(out[0].abs().max())
    err = float((out[0] - out[1]).abs().max())
    assert scale > 0 and err < 1e-5 * scale, (err, scale)
    untouched = int(((out[0] == 0) != (out[1] == 0)).sum())   # same support (exact zeros where nothing was scattered)
    assert untouched < 1e-4 * out[0].numel(), untouched


def test_encode_rays_bwd_delta_basis_equals_absolute_basis(lib):
    R, n, taps = 64, 16, 4
    grid = _grid(lib)
    center, ray, dists = _rays(R, seed=11)
    eps = 1.0 / 2048 / math.sqrt(3)
    M, P = R * n, 1 + taps
    torch.manual_seed(12)
    dX = torch.randn(P, M, 128)
    dX[1:] *= 100.0
    dX[0] -= dX[1:].sum(0) * 0.999   # large, nearly cancelling per-plane gradients like the real stencil's
    n_par = int(grid.n_entries) * 8
    tg_abs, tg_del = torch.zeros(n_par, device="cuda"), torch.zeros(n_par, device="cuda")
    args = (grid, center.cuda(), ray.cuda(), dists.cuda(), 16, R, n, taps, eps, -2.0, 2.0)
    lib.call("mli_encode_rays_bwd", *args, dX.reshape(P * M, 128).cuda(), 128, tg_abs, 0)
    dXd = dX.clone()
    dXd[0] = dX.sum(0)
    lib.call("mli_encode_rays_bwd", *args, dXd.reshape(P * M, 128).cuda(), 128, tg_del, 1)
    scale = float(tg_abs.abs().max())
    assert scale > 0
    assert float((tg_abs - tg_del).abs().max()) < 2e-4 * scale, (float((tg_abs - tg_del).abs().max()), scale)
    # bf16 TCL input variant (what the tensor-core path uses): equals the fp32 delta-basis kernel on bf16-rounded input
    tg_ref, tg_tcl = torch.zeros(n_par, device="cuda"), torch.zeros(n_par, device="cuda")
    lib.call("mli_encode_rays_bwd", *args, bf(dXd).reshape(P * M, 128).cuda(), 128, tg_ref, 1)
    dXt = to_tcl_host(dXd.reshape(P * M, 128)).cuda()
    for lv0, lv1 in ((0, 6), (6, 7), (7, 16)):  # level groups, as the overlapped multi-GPU backward launches them
        lib.call("mli_encode_rays_bwd_tcl", *args, dXt, 16, tg_tcl, lv0, lv1)
    assert float((tg_ref - tg_tcl).abs().max()) < 1e-5 * float(tg_ref.abs().max())


def test_relu_sign_mask_roundtrip(lib):
    """relu forward writes one sign bit per output; the relu data gradient with that mask equals the one that reads the
    bf16 activations (both batched, as the head layers use them)."""
    torch.manual_seed(13)
    M, N, K, batch = 640, 256, 256, 3
    X = torch.randn(M, batch * K) * 0.5
    W = torch.randn(batch, N, K) / math.sqrt(K)
    b = torch.randn(batch, N) * 0.1
    A = to_tcl_host(X).cuda()
    Wt = torch.stack([to_tcl_host(W[i], 256) for i in range(batch)]).cuda()
    out = torch.zeros(M // 128, batch * N // 8, 128, 8, dtype=torch.bfloat16, device="cuda")
    mask = torch.zeros(M // 128, batch * N // 32, 128, dtype=torch.int32, device="cuda")
    lib.call("mli_tc_linear", A, batch * K // 8, 0, K // 8, Wt, N * K, K, N, 256, b.cuda(), N, None, 0, 0, 0, 1, out, 0,
             batch * N // 8, 0, N // 8, 0, 0, 0, M, batch, 0, mask, batch * N // 32, 0, N // 32)
    y = from_tcl_host(out.cpu(), M)                                  # [M, batch*N] relu outputs
    bits = mask.cpu().permute(0, 2, 1).reshape(M, batch * N // 32)   # [M, chunks32]
    got = ((bits.unsqueeze(-1) >> torch.arange(32)) & 1).reshape(M, batch * N).bool()
    # bf16 rounding can flush a tiny positive fp32 output to +0: the mask (taken before rounding) may only have MORE ones
    assert bool((got | ~(y > 0)).all()) and float((got != (y > 0)).float().mean()) < 1e-3
    dZ = to_tcl_host(torch.randn(M, batch * N)).cuda()
    Wtt = torch.stack([to_tcl_host(W[i].t().contiguous(), 256) for i in range(batch)]).cuda()
    o1 = torch.zeros(M // 128, batch * K // 8, 128, 8, dtype=torch.bfloat16, device="cuda")
    o2 = torch.zeros_like(o1)
    # data gradient of a layer whose INPUT activations are `out` (N == K here): relu' from aux vs from the mask
    lib.call("mli_tc_linear", dZ, batch * N // 8, 0, N // 8, Wtt, N * K, N, K, 256, None, 0, out, batch * N // 8, 0, N // 8, 1,
             o1, 0, batch * K // 8, 0, K // 8, 0, 0, 0, M, batch, 1, None, 0, 0, 0)
    lib.call("mli_tc_linear", dZ, batch * N // 8, 0, N // 8, Wtt, N * K, N, K, 256, None, 0, None, 0, 0, 0, 1,
             o2, 0, batch * K // 8, 0, K // 8, 0, 0, 0, M, batch, 1, mask, batch * N // 32, 0, N // 32)
    d = (from_tcl_host(o1.cpu(), M) != from_tcl_host(o2.cpu(), M)).float().mean()
    assert float(d) < 1e-3, float(d)


@pytest.mark.parametrize("mode,R", [("rgb_r_s", 256), ("rgb", 131), ("r_s_re", 64), ("rgb_r_s", 2400), ("rgb_r_s", 2401)])
def test_fused_head_stack_matches_layer_by_layer(lib, mode, R):
    """csrc/heads_fused.cu (the whole head stack in one on-chip tcgen05 kernel, two 128-sample tiles per CTA) against the
    layer-by-layer tensor-core path on the same inputs: same bf16 rounding points -> the stored activations, relu sign
    bits and per-sample outputs agree to bf16 resolution.  R = 131 / 2401 give an ODD
    number of 128-sample tiles (the last tile pair is half empty) resp. fewer pairs than SMs; R = 2401 gives 1201 tile pairs
    = 8 or 9 per persistent CTA (barrier phases carried across pairs); rgb = one head, r_s_re = nine narrow outputs."""
    from mli_nerf_b200.engine import RenderEngine
    from tests.util import make_case, product_cfg
    from mli_nerf_b200.engine import head_layout
    case = make_case(R=R, mode=mode, progress=0.5)
    J = sum(h[2] for h in head_layout(mode if mode != "rgb" else None))
    res = {}
    for fused in (False, True):
        eng = RenderEngine(product_cfg(case["ocfg"], precision=1))
        eng.fuse_heads = eng.fuse_heads_train_fwd = fused  # also the (default-off) fused training forward
        p = {k: v.contiguous().cuda() for k, v in case["params"].items()}
        eng.pack_weights(p)
        c, r, l = (case[k][0].contiguous().cuda() for k in ("center", "ray_unit", "light"))
        near, far, outside = eng.bounds(c, r)
        if "dists" not in res:  # the product's own sampling (bit-reproducible): the same distances for both runs
            res["dists"] = eng.sample(p["neural_sdf.tcnn_encoding.params"], c, r, near, far, case["rands"].view(R, -1).cuda())
        dists = res["dists"]
        out, ctx = eng.forward(p, c, r, l, dists, near, far, outside, True, 0.5)
        torch.cuda.synchronize()
        res[fused] = (out, ctx)
        # forward-only call (no backward follows): nothing but S leaves the SM, same S
        out2, ctx2 = eng.forward(p, c, r, l, dists, near, far, outside, False, 0.5, keep_dz=False)
        assert torch.equal(out2["S"][:, :J], out["S"][:, :J])  # columns >= J of S are padding (never written)
        if fused:
            assert ctx2["A"][0] is None
    (o0, c0), (o1, c1) = res[False], res[True]
    # The fused kernel adds the bias inside the accumulator (it is pre-loaded into TMEM) instead of after the sum: a
    # different fp32 association, so a bf16 rounding flips now and then and the flips propagate through the layers --
    # the two paths agree to bf16 resolution, not bit for bit.
    assert float((o0["S"][:, :J] - o1["S"][:, :J]).abs().max()) < 2e-3
    assert float((o0["out"] - o1["out"]).abs().max()) < 2e-3
    for l in range(4):
        a0, a1 = c0["A"][l].float(), c1["A"][l].float()
        frac_a = float(a0.ne(a1).float().mean())
        err_a, max_a = float((a0 - a1).abs().max()), float(a0.abs().max())
        frac_m = float(c0["Am"][l].ne(c1["Am"][l]).float().mean())
        assert frac_a < 0.25, (l, frac_a)            # 1-ulp bf16 flips (growing with depth) are tolerated ...
        assert err_a < 2e-2 * max_a, (l, err_a)      # ... a wrong value is not
        assert frac_m < 1e-2, (l, frac_m)
    # the fused data-gradient chain (mli_tc_heads_bwd) against the layer-by-layer chain: every parameter gradient
    torch.manual_seed(5)
    d_out = torch.randn_like(o0["out"]) * 1e-2
    d_grad = torch.randn_like(o0["gradients"]) * 1e-4
    grads = {}
    for fused in (False, True):
        eng = RenderEngine(product_cfg(case["ocfg"], precision=1))
        eng.fuse_heads = fused
        p = {k: v.contiguous().cuda() for k, v in case["params"].items()}
        eng.pack_weights(p)
        grads[fused] = eng.backward(p, res[fused][1], d_out, d_grad, None, None)
        torch.cuda.synchronize()
    errs = {k: float((g0 - grads[True][k]).norm() / (g0.norm() + 1e-30)) for k, g0 in grads[False].items()}
    bad = {k: (e, bool(torch.isfinite(grads[False][k]).all()), bool(torch.isfinite(grads[True][k]).all()))
           for k, e in errs.items() if not e < 2e-2}
    assert not bad, bad  # (relative error, layer-by-layer finite, fused finite)


@pytest.mark.parametrize("nbytes", [0, 4, 16, 4096 + 12, (1 << 22) + 20])
def test_zero_fill_background(lib, nbytes):
    """Small-footprint zero-fill: every byte of the range, nothing after it, odd tails included."""
    buf = torch.full((nbytes + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    lib.call("mli_zero_fill_background", buf, nbytes, 8)
    host = buf.cpu()
    assert int(host[:nbytes].sum()) == 0
    assert bool((host[nbytes:] == 0x5A).all())
