"""Golden vectors produced by the UNMODIFIED reference (oracle/gen_golden.py, run in the build container):
  * CPU: oracle/port.py and the plain-C checker must reproduce them (pins the oracle on any box),
  * GPU (-m gpu): the CUDA path through the C ABI must reproduce them.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import port
from tests.util import loss_cfg, product_cfg, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "render_hotdog_b_train": dict(cfg=dict(log2_hashmap_size=14)),
    "render_hotdog_b_eval": dict(cfg=dict(log2_hashmap_size=14)),
    "render_rene_b_train": dict(cfg=dict(log2_hashmap_size=14, bounding="box", aabb=(-0.66, -0.516, -0.18, 0.66, 0.42, 0.3),
                                         white_background=False)),
}


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def digest(p):
    h = hashlib.sha256()
    for k in sorted(p):
        h.update(k.encode())
        h.update(p[k].detach().numpy().tobytes())
    return h.hexdigest()


def case_params(name):
    g = load(name)
    ocfg = port.PathConfig(**CASES[name]["cfg"])
    p = port.init_params(ocfg, seed=int(g["seed"]), generic=True, table_scale=5e-3)
    assert digest(p) == bytes(g["params_sha256"]).decode(), "seeded weights differ from the ones the fixture was made with"
    return g, ocfg, p


@pytest.fixture(scope="module")
def oracle_c():
    d = os.path.join(os.path.dirname(GOLD), "..", "oracle", "c")
    subprocess.check_call(["make", "-s", "-C", d])
    return C.CDLL(os.path.join(d, "liboracle_int.so"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------------------------
def test_sampling_kat_port_and_c(oracle_c):
    g = load("sampling_kat")
    for n_w in (63, 79, 95, 111):
        w, bins = torch.from_numpy(g[f"w{n_w}"]), torch.from_numpy(g[f"bins{n_w}"])
        d, info = port.sample_from_pdf(bins[None, ..., None], w[None], 16)
        assert np.array_equal(info["idx"][0].numpy(), g[f"idx{n_w}"]) and np.array_equal(info["cdf"][0].numpy(), g[f"cdf{n_w}"])
        assert np.array_equal(d[0, :, :, 0].numpy(), g[f"dists{n_w}"])
        cdf = np.zeros((200, n_w + 1), np.float32)
        idx = np.zeros((200, 16), np.int32)
        dist = np.zeros((200, 16), np.float32)
        oracle_c.oc_pdf_bins(_p(np.ascontiguousarray(g[f"w{n_w}"])), _p(np.ascontiguousarray(g[f"bins{n_w}"])), C.c_int64(200),
                             n_w, 16, _p(cdf), _p(idx), _p(dist))
        assert np.array_equal(cdf, g[f"cdf{n_w}"]) and np.array_equal(idx, g[f"idx{n_w}"])
        assert np.array_equal(dist, g[f"dists{n_w}"])


def test_hash_index_kat_c_and_product_table(oracle_c, mli_lib):
    import math
    g = load("hash_index_kat")
    x = np.ascontiguousarray(g["x"])
    pls = math.exp((math.log(2048) - math.log(32)) / 15)

    class Lv(C.Structure):
        _fields_ = [("scale", C.c_float), ("res", C.c_uint32), ("size", C.c_uint32), ("offset", C.c_uint32), ("hashed", C.c_uint32)]
    for T in (14, 22):
        grid = mli_lib.make_grid(16, 8, T, 32, pls)
        for level in range(16):
            s, res, size, off, hashed = g[f"levels{T}"][level]
            assert (grid.level[level].scale, grid.level[level].res, grid.level[level].size, grid.level[level].offset,
                    grid.level[level].hashed) == (np.float32(s), int(res), int(size), int(off), int(hashed))
            lv = Lv(float(s), int(res), int(size), int(off), int(hashed))
            rows = np.zeros((x.shape[0], 8), np.uint32)
            oracle_c.oc_grid_rows(C.byref(lv), _p(x), C.c_int64(x.shape[0]), _p(rows))
            assert np.array_equal(rows.astype(np.int64), g[f"rows{T}"][level]), (T, level)


@pytest.mark.parametrize("name", list(CASES))
def test_port_reproduces_reference_render(name):
    g, ocfg, p = case_params(name)
    training = bool(g["training"])
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    out = port.render_rays(pp, ocfg, torch.from_numpy(g["center"]), torch.from_numpy(g["ray_unit"]), torch.from_numpy(g["light"]),
                           rands=torch.from_numpy(g["rands"]), training=training, progress=float(g["progress"]))
    assert np.array_equal(out["outside"].numpy(), g["out_outside"])
    same = np.abs(out["dists"].numpy() - g["out_dists"]).max(axis=(2, 3))[0] < 1e-6  # rays with identical samples
    assert same.mean() > 0.3  # the rest differ by the fp32-noise amplification of hierarchical sampling
    for k in ("rgb", "o_r", "o_s", "o_re", "weights", "gradients", "opacity", "gradient"):
        if "out_" + k in g:
            a, b = out[k].detach().numpy()[0][same], g["out_" + k][0][same]
            # numerical SDF gradients carry fp32 cancellation noise (differences of SDFs / 1e-3): looser absolute floor
            atol = (5e-3 if k in ("gradients", "gradient") else 2e-4) * max(1.0, np.abs(b).max())
            assert np.allclose(a, b, rtol=1e-3, atol=atol), k
    if training:
        R = g["center"].shape[1]
        total, losses, _ = port.total_loss(ocfg, out, port.synthetic_targets(R, seed=int(g["seed"]) + 2))
        assert abs(float(total) - float(g["loss_total"])) < 2e-3 * abs(float(g["loss_total"]))
        total.backward()
        for k, v in pp.items():
            n_ref = float(g["gnorm_" + k])
            assert abs(float(v.grad.norm()) - n_ref) < 5e-2 * n_ref + 1e-9, k


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_reproduces_reference_render(name, prec):
    """CUDA path (through the C ABI) vs the reference's own outputs, fed the reference's sample distances.
    fp32 mode: rtol 1e-3.  bf16 tcgen05 mode (the measured one), stated looser bounds: per-ray colours |err| <= 2e-2
    (mean <= 3e-3), compositing weights |err| <= 2e-2, numerical gradients relative L2 <= 2e-3, losses 2 %, parameter
    gradient norms 6 % (SDF-side: see `loose`)."""
    from mli_nerf_b200.engine import RenderEngine
    g, ocfg, p = case_params(name)
    training = bool(g["training"])
    bf = prec == "bf16"
    eng = RenderEngine(product_cfg(ocfg, precision=1 if bf else 0))
    pc = {k: v.cuda() for k, v in p.items()}
    eng.pack_weights(pc)
    c, r, l = (torch.from_numpy(g[k][0]).cuda() for k in ("center", "ray_unit", "light"))
    near, far, outside = eng.bounds(c, r)
    assert np.array_equal(outside.cpu().numpy().astype(bool), g["out_outside"][0, :, 0])
    res, ctx = eng.forward(pc, c, r, l, torch.from_numpy(g["out_dists"][0, :, :, 0]).cuda(), near, far, outside, training,
                           float(g["progress"]))
    out = res["out"].cpu().numpy()
    R = c.shape[0]
    for k, (a, b) in dict(rgb=(0, 3), o_r=(3, 6), o_s=(6, 7), o_re=(7, 10)).items():
        if bf:
            err = np.abs(out[:, a:b] - g["out_" + k][0])
            assert err.max() < 2e-2 and err.mean() < 3e-3, (k, err.max(), err.mean())
        else:
            assert np.allclose(out[:, a:b], g["out_" + k][0], rtol=1e-3, atol=2e-5), k
    assert np.allclose(res["weights"].cpu().numpy(), g["out_weights"][0, :, :, 0], rtol=1e-3, atol=2e-2 if bf else 2e-5)
    inside = ~g["out_outside"][0, :, 0]
    gr = res["gradients"].cpu().numpy().reshape(R, 128, 3)
    assert np.abs(gr[inside] - g["out_gradients"][0][inside]).max() < 2e-3 * np.abs(g["out_gradients"][0][inside]).max()
    if not training:
        ex = res["extras"].cpu().numpy()
        assert np.allclose(ex[:, 0:1], g["out_opacity"][0], rtol=1e-3, atol=2e-2 if bf else 2e-5)
        assert np.allclose(ex[:, 1:4], g["out_gradient"][0], rtol=2e-3, atol=2e-2 if bf else 2e-3)
        return
    tg = {k: v[0].cuda() for k, v in port.synthetic_targets(R, seed=int(g["seed"]) + 2).items()}
    losses, d_out, d_grad, d_hess = eng.losses(loss_cfg(ocfg), res["out"], res["gradients"], res["hessians"], outside, tg)
    lc = losses.cpu().numpy()
    assert abs(lc[0] - float(g["loss_total"])) < (2e-2 if bf else 2e-3) * abs(float(g["loss_total"]))
    for i, k in ((1, "render"), (2, "eikonal"), (4, "intrinsic"), (5, "regularize_re")):
        assert abs(lc[i] - float(g["loss_" + k])) < (2e-2 if bf else 1e-3) * abs(float(g["loss_" + k])) + 1e-6, k
    grads = eng.backward(pc, ctx, d_out, d_grad, d_hess, None)
    for k in p:
        gn, ref = float(grads[k].norm()), float(g["gnorm_" + k])
        loose = "neural_sdf" in k or k == "s_var"  # curvature seed = sign(laplacian): fp32-noise sensitive, see parity test
        scalar = grads[k].numel() == 1  # a scalar gradient is a heavily cancelling sum of the noisy per-sample terms
        tol_n = ((0.5 if scalar else 1.5e-1 if bf else 1e-1) if loose else 6e-2 if bf else 1e-2)
        assert abs(gn - ref) < tol_n * ref + 1e-9, (k, gn, ref)
        if grads[k].numel() <= 4096 and not (loose and scalar):
            tol_e = (1.5e-1 if bf else 1e-1) if loose else (6e-2 if bf else 5e-3)
            assert rel_err(grads[k].cpu().reshape(-1), torch.from_numpy(g["grad_" + k]).reshape(-1)) < tol_e, k


# ---------------------------------------------------------------------------------------------------------------
# light visibility: Model.get_light_visibility of the reference, three camera_ray_types
# ---------------------------------------------------------------------------------------------------------------
LV_TYPES = ("blend_z_sphere_tracing", "sphere_tracing", "blend_z")


def _lv_case():
    g = load("light_visibility_hotdog_b")
    ocfg = port.PathConfig(log2_hashmap_size=14)
    p = port.init_params(ocfg, seed=int(g["seed"]), generic=False)
    assert digest(p) == bytes(g["params_sha256"]).decode(), "seeded weights differ from the ones the fixture was made with"
    return g, ocfg, p


@pytest.mark.parametrize("ray_type", LV_TYPES)
def test_port_reproduces_reference_light_visibility(ray_type):
    g, ocfg, p = _lv_case()
    center, ray_unit, light = (torch.from_numpy(g[k]) for k in ("center", "ray_unit", "light"))
    with torch.no_grad():
        near, far, _ = port.dist_bounds(ocfg, center, ray_unit)
        vis, nxl, idist, imask = port.light_visibility(p, ocfg, center, ray_unit, light, near, far, torch.from_numpy(g["blend"]),
                                                       torch.from_numpy(g["gradient"]), ray_type, float(g["radius"]))
    assert np.array_equal(imask.numpy(), g[ray_type + "_inter_mask"])
    assert np.allclose(idist.numpy(), g[ray_type + "_inter_dist"], rtol=1e-4, atol=1e-5)
    assert np.allclose(nxl.numpy(), g[ray_type + "_normal_x_light"], rtol=1e-3, atol=1e-4)
    assert (vis.numpy() == g[ray_type + "_visibility"]).mean() >= 0.97


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("ray_type", LV_TYPES)
def test_cuda_reproduces_reference_light_visibility(ray_type, prec):
    """RenderEngine.light_visibility (csrc/visibility.cu + the encode / SDF-trunk kernels in a 20 + 20 iteration loop)
    against the reference's own outputs, fed the reference's composited distance and gradient."""
    from mli_nerf_b200.engine import RenderEngine
    g, ocfg, p = _lv_case()
    eng = RenderEngine(product_cfg(ocfg, precision=1 if prec == "bf16" else 0))
    pc = {k: v.cuda() for k, v in p.items()}
    eng.pack_weights(pc)
    c, r, l = (torch.from_numpy(g[k][0]).cuda() for k in ("center", "ray_unit", "light"))
    near, far, _ = eng.bounds(c, r)
    vis, nxl, idist, imask = eng.light_visibility(pc["neural_sdf.tcnn_encoding.params"], c, r, l, near, far,
                                                  torch.from_numpy(g["blend"][0, :, 0]).cuda(),
                                                  torch.from_numpy(g["gradient"][0]).cuda(), ray_type, float(g["radius"]))
    R = c.shape[0]
    ref_d, ref_m = g[ray_type + "_inter_dist"][0, :, 0], g[ray_type + "_inter_mask"][0, :, 0]
    tol = 1e-4 if prec == "fp32" else 2e-3
    ok = np.abs(idist.cpu().numpy().reshape(R) - ref_d) < tol * (1 + np.abs(ref_d))
    assert ok.mean() > 0.97, ok.mean()
    assert (imask.cpu().numpy().reshape(R).astype(bool) == ref_m).mean() > 0.97
    assert (vis.cpu().numpy().reshape(R).astype(bool) == g[ray_type + "_visibility"][0, :, 0]).mean() > 0.95
    d_n = np.abs(nxl.cpu().numpy().reshape(R) - g[ray_type + "_normal_x_light"][0, :, 0])
    assert (d_n < (1e-3 if prec == "fp32" else 5e-3)).mean() > 0.97


# ---------------------------------------------------------------------------------------------------------------
# full-image inference: the reference's own Model.inference maps on a 12 x 14 view
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_cuda_reproduces_reference_inference_maps(prec):
    from mli_nerf_b200 import config
    from mli_nerf_b200.model import Model
    g = load("inference_hotdog_b")
    ocfg = port.PathConfig(log2_hashmap_size=14)
    p = port.init_params(ocfg, seed=int(g["seed"]), generic=False)
    assert digest(p) == bytes(g["params_sha256"]).decode(), "seeded weights differ from the ones the fixture was made with"
    H, W = (int(v) for v in g["image_size"])
    cfg = config.experiment("syn_hotdog_b", dict_size=14)
    cfg.data.val.image_size = [H, W]
    cfg.model.render.rand_rays_val = 50  # 168 rays: three full chunks + a ragged one, like the fixture
    cfg.model.mli_precision = prec
    model = Model(cfg.model, cfg.data)
    model.load_state_dict(p)
    model = model.cuda()
    data = dict(pose=torch.from_numpy(g["pose"]).cuda(), intr=torch.from_numpy(g["intr"]).cuda(),
                pose_light=torch.from_numpy(g["pose_light"]).cuda(), idx=torch.zeros(1, dtype=torch.long))
    out = model.inference(data)
    assert np.array_equal(out["outside"].cpu().numpy(), g["outside"])
    tight = dict(fp32=2e-3, bf16=2e-2)[prec]
    for k in ("rgb_map", "opacity_map", "depth_map", "o_r_map", "o_s_map", "o_re_map", "normal_map"):
        a, b = out[k].cpu().numpy(), g[k]
        assert a.shape == b.shape, k
        err = np.abs(a - b).max(axis=1).reshape(-1)  # per pixel
        # independently sampled distances: most pixels agree tightly, a bin flip moves a pixel by ~1e-2
        loose = 0.15 if k == "normal_map" else 5e-2
        assert err.max() < loose and (err < tight * (1 + np.abs(b).max())).mean() > 0.9, (k, err.max(), (err < tight).mean())
