"""CPU checks of the kernels' scalar math (mli_nerf_b200/csrc/mli_math.h compiled by g++, tests/hostsim) against
the oracle.  Integer decisions (hash-grid corner rows, inverse-CDF bins, outside flags) must be bit-exact."""
import ctypes as C
import math

import numpy as np
import torch

from oracle import port
from oracle.torch_hashgrid import corner_indices, level_table
from mli_nerf_b200 import _lib


def _p(t):
    return C.c_void_p(t.data_ptr())


def _pls():
    return math.exp((math.log(2048) - math.log(32)) / 15)


def test_level_table_matches_oracle():
    for T in (14, 19, 22):
        lv, n = level_table(16, T, 32, _pls())
        g = _lib.make_grid(16, 8, T, 32, _pls())
        assert g.n_entries == n
        for i in range(16):
            assert (lv[i]["scale"], lv[i]["res"], lv[i]["size"], lv[i]["offset"], int(lv[i]["hashed"])) == \
                (g.level[i].scale, g.level[i].res, g.level[i].size, g.level[i].offset, g.level[i].hashed)


def test_corner_rows_bit_exact(hostsim):
    torch.manual_seed(0)
    x = torch.rand(4000, 3)
    # edge cases: exact 0/1, outside [0,1] (uint32 wrap), cell boundaries
    x[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [-0.25, 0.5, 1.5], [1.0, 0.0, 0.5], [0.5, 0.5, 0.5],
                          [-1e-7, 1 + 1e-7, 0.3], [0.999999, 1e-8, 0.25], [2.5, -3.0, 0.1]])
    for T in (14, 22):
        lv, _ = level_table(16, T, 32, _pls())
        g = _lib.make_grid(16, 8, T, 32, _pls())
        for level in range(16):
            idx = torch.zeros(x.shape[0], 8, dtype=torch.int32)
            wt = torch.zeros(x.shape[0], 8)
            hostsim.hs_corners(C.byref(g), C.c_uint32(level), _p(x), C.c_int64(x.shape[0]), _p(idx), _p(wt))
            ref_idx, ref_w = corner_indices(x, lv[level])
            got = idx.to(torch.int64) & 0xFFFFFFFF
            assert torch.equal(got, ref_idx + lv[level]["offset"]), f"T={T} level={level}"
            assert torch.allclose(wt, ref_w, rtol=0, atol=1e-6)


def test_bounds_bit_exact(hostsim):
    c, r, _ = port.synthetic_rays(5000, seed=4)
    c, r = c[0].contiguous(), r[0].contiguous()
    r[:50] = torch.nn.functional.normalize(torch.randn(50, 3), dim=-1)  # many misses
    for cfg in (port.PathConfig(), port.PathConfig(bounding="box", aabb=(-0.66, -0.516, -0.18, 0.66, 0.42, 0.3))):
        near, far, out = torch.zeros(5000), torch.zeros(5000), torch.zeros(5000, dtype=torch.uint8)
        aabb = (C.c_float * 6)(*cfg.aabb) if cfg.aabb else None
        hostsim.hs_bounds(_p(c), _p(r), C.c_int64(5000), aabb, _p(near), _p(far), _p(out))
        n_ref, f_ref, o_ref = port.dist_bounds(cfg, c[None], r[None])
        assert torch.equal(out.bool(), o_ref[0, :, 0])
        assert 0 < int(out.sum()) < 5000
        # near/far: identical up to torch's vectorised sqrt (1 ulp on a handful of rays); the flag is exact
        assert torch.allclose(near, n_ref[0, :, 0], rtol=3e-7, atol=0) and torch.allclose(far, f_ref[0, :, 0], rtol=3e-7, atol=0)
        assert (near != n_ref[0, :, 0]).float().mean() < 0.01


def test_sample_points_bit_exact(hostsim):
    c, r, _ = port.synthetic_rays(2000, seed=5)
    c, r = c[0].contiguous(), r[0].contiguous()
    d = torch.rand(2000) * 4
    e = (1.0 / 2048) / math.sqrt(3)
    ks = ([1, -1, -1], [-1, -1, 1], [-1, 1, -1], [1, 1, 1])
    for plane in range(5):
        p = torch.zeros(2000, 3)
        hostsim.hs_points(_p(c), _p(r), _p(d), C.c_int64(2000), 4, plane, C.c_float(e), _p(p))
        ref = c + r * d[:, None]
        if plane:
            ref = ref + torch.tensor(ks[plane - 1], dtype=torch.float32) * e
        assert torch.equal(p, ref)


def test_unif_matches_torch(hostsim):
    for n_fine in (16, 8, 12, 32, 5):
        u = torch.zeros(n_fine)
        hostsim.hs_unif(n_fine, _p(u))
        grid = torch.linspace(0, 1, n_fine + 1)
        assert torch.equal(u, 0.5 * (grid[:-1] + grid[1:])), n_fine


def _rand_ray_state(R, n, seed):
    g = torch.Generator().manual_seed(seed)
    near = torch.rand(R, 1, generator=g) * 2
    d = (near + torch.sort(torch.rand(R, n, generator=g), dim=1).values * 2).contiguous()
    # sdf that crosses zero somewhere for most rays, flat/positive for some (empty rays -> flat cdf)
    t0 = torch.rand(R, 1, generator=g) * 2 + near
    s = (t0 - d) * (0.2 + torch.rand(R, 1, generator=g)) + 0.01 * torch.randn(R, n, generator=g)
    s[: R // 10] = 0.5 + 0.1 * torch.rand(R // 10, n, generator=g)
    s[R // 10: R // 5] = 50.0
    return d, s.contiguous()


def test_hierarchical_bins_bit_exact_given_weights(hostsim):
    """sample_dists_from_pdf: idx/low/high must equal torch's searchsorted on the same weights."""
    for n_w in (63, 79, 95, 111):
        g = torch.Generator().manual_seed(n_w)
        w = torch.rand(3000, n_w, generator=g) * (torch.rand(3000, n_w, generator=g) > 0.5)
        w[:100] = 0.0  # all-zero weights: flat cdf -> idx = N
        w[100:200, 10:] = 0.0
        w[200:300] = w[200:300] * 1e-30
        idx, low, high = (torch.zeros(3000, 16, dtype=torch.int32) for _ in range(3))
        cdf = torch.zeros(3000, n_w + 1)
        hostsim.hs_pdf_bins(_p(w), C.c_int64(n_w), C.c_int64(3000), n_w, 16, _p(idx), _p(low), _p(high), _p(cdf))
        bins = torch.arange(n_w + 1, dtype=torch.float32).expand(3000, -1)[None, ..., None]
        _, info = port.sample_from_pdf(bins, w[None], 16)
        assert torch.equal(cdf, info["cdf"][0])
        assert torch.equal(idx.long(), info["idx"][0])
        assert torch.equal(low.long(), info["low"][0]) and torch.equal(high.long(), info["high"][0])


def test_sample_fine_matches_oracle(hostsim):
    for n, inv_s in ((64, 64.0), (80, 128.0), (96, 256.0), (112, 512.0)):
        d, s = _rand_ray_state(2000, n, seed=n)
        fine = torch.zeros(2000, 16)
        idx, low, high = (torch.zeros(2000, 16, dtype=torch.int32) for _ in range(3))
        cdf, w = torch.zeros(2000, n), torch.zeros(2000, n - 1)
        hostsim.hs_sample_fine(_p(d), _p(s), C.c_int64(n), C.c_int64(2000), n, 16, C.c_float(inv_s), _p(fine), _p(idx),
                               _p(low), _p(high), _p(cdf), _p(w))
        w_ref = port.hierarchical_weights(d[None, ..., None], s[None, ..., None], inv_s)
        fine_ref, info = port.sample_from_pdf(d[None, ..., None], w_ref, 16)
        # weights agree to float rounding (expf vs torch's vectorised exp) ...
        assert torch.allclose(w, w_ref[0], rtol=1e-4, atol=1e-7)
        # ... and the bins are identical wherever the two cdfs are (they are computed from ulp-different weights)
        same = torch.equal(idx.long(), info["idx"][0])
        frac = (idx.long() == info["idx"][0]).float().mean().item()
        assert same or frac > 0.999, frac
        ok = (idx.long() == info["idx"][0]).all(dim=1)
        assert torch.allclose(fine[ok], fine_ref[0, ok, :, 0], rtol=1e-4, atol=1e-5)


def test_sh_and_activations(hostsim):
    torch.manual_seed(1)
    d = torch.randn(1000, 3) * 2
    out = torch.zeros(1000, 16)
    hostsim.hs_sh16(_p(d), C.c_int64(1000), _p(out))
    assert torch.allclose(out, port.sh_basis(d, 3), rtol=1e-5, atol=1e-5)
    x = torch.cat([torch.randn(2000) * 0.1, torch.tensor([0.0, 0.2, 0.2000001, 0.25, -1.0, 5.0, -5.0])])
    for act, f in ((1, torch.relu), (2, lambda v: torch.nn.functional.softplus(v, beta=100)), (3, torch.sigmoid)):
        xx = x.clone().requires_grad_(True)
        y_ref = f(xx)
        (g_ref,) = torch.autograd.grad(y_ref.sum(), xx)
        y, dy = torch.zeros_like(x), torch.zeros_like(x)
        hostsim.hs_act(_p(x), C.c_int64(x.numel()), act, _p(y), _p(dy))
        assert torch.allclose(y, y_ref.detach(), rtol=1e-5, atol=1e-7), act
        assert torch.allclose(dy, g_ref, rtol=2e-4, atol=1e-6), act


def test_neus_alpha_forward_backward(hostsim):
    torch.manual_seed(2)
    n = 4000
    sdf = (torch.randn(n) * 0.05).requires_grad_(True)
    g = torch.randn(n, 3).requires_grad_(True)
    r = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
    intv = torch.rand(n) * 0.05
    for progress in (0.03, 0.5):
        s_var = torch.tensor(3.0, requires_grad=True)
        anneal = min(progress / 0.1, 1.0)
        inv_s = s_var.exp()
        true_cos = (r * g).sum(-1)
        iter_cos = -((-true_cos * 0.5 + 0.5).relu() * (1.0 - anneal) + (-true_cos).relu() * anneal)
        p_cdf = ((sdf - iter_cos * intv * 0.5) * inv_s).sigmoid()
        n_cdf = ((sdf + iter_cos * intv * 0.5) * inv_s).sigmoid()
        alpha_ref = ((p_cdf - n_cdf) / (p_cdf + 1e-5)).clip(0.0, 1.0)
        g_sdf, g_g, g_s = torch.autograd.grad(alpha_ref.sum(), (sdf, g, s_var))
        alpha, d_sdf, d_g, d_is = torch.zeros(n), torch.zeros(n), torch.zeros(n, 3), torch.zeros(n)
        hostsim.hs_neus_alpha(_p(sdf.detach()), _p(g.detach().contiguous()), _p(r), _p(intv), C.c_int64(n),
                              C.c_float(float(inv_s)), C.c_float(anneal), _p(alpha), _p(d_sdf), _p(d_g), _p(d_is))
        assert torch.allclose(alpha, alpha_ref.detach(), rtol=1e-4, atol=1e-6)
        assert torch.allclose(d_sdf, g_sdf, rtol=1e-3, atol=1e-4)
        assert torch.allclose(d_g, g_g, rtol=1e-3, atol=1e-5)
        assert abs(float(d_is.sum() * inv_s) - float(g_s)) < 1e-3 * max(1.0, abs(float(g_s)))


def test_level_table_matches_float32_host_arithmetic_of_tcnn():
    """tcnn sizes the table on the host with exp2f(l * log2f(s)) * base - 1 in float32 steps; the float64 formula rounded
    once (oracle and product) must give the same resolutions -- incl. the knife-edge levels 5 / 10 / 15 of the
    Neuralangelo config (2^0.4 growth: exact values 127 / 511 / 2047) -- so that reference checkpoints load."""
    import math
    import numpy as np
    from oracle.torch_hashgrid import level_table
    for lo, hi, n in ((5, 11, 16), (5, 12, 16), (4, 11, 16), (5, 11, 8)):
        pls = math.exp((math.log(2 ** hi) - math.log(2 ** lo)) / (n - 1))
        s = np.float32(pls)
        l2 = np.float32(math.log2(float(s)))
        res32 = []
        for lv in range(n):
            e = np.float32(math.pow(2.0, float(np.float32(np.float32(lv) * l2))))
            scale = np.float32(np.float32(e * np.float32(2 ** lo)) - np.float32(1))
            res32.append(int(math.ceil(float(scale))) + 1)
        table, _ = level_table(n, 22, 2 ** lo, pls)
        assert res32 == [t["res"] for t in table], (lo, hi, n)
    table, total = level_table(16, 22, 32, math.exp((math.log(2048) - math.log(32)) / 15))
    assert [table[i]["res"] for i in (5, 10, 15)] == [129, 513, 2049] and total == 45724048
