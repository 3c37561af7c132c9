"""GPU parity at the BENCHMARKED configuration and for the network modes no shipped YAML uses.

* log2_hashmap_size = 22 (the syn_hotdog_b table bench.py measures): levels 0-5 take the dense-stride path, levels 6-15
  are hashed; every other render test runs at T = 2^14 where all 16 levels are hashed.  The CUDA path (fp32 CUDA-core
  mode and the bf16 tcgen05 mode: encode_rays_tcl -> fused trunk -> head GEMMs -> encode_rays_bwd_tcl) is compared with
  the CPU oracle on the same rays, weights and sample distances, the 1.46 GB table gradient separately on the dense and
  the hashed levels.
* network_mode r_s / r_s_re / rgb_r (/root/reference/projects/NeuralLumen/utils/modules.py:110-174, merge at
  projects/NeuralLumen/model.py:266-310; note r_s's o_s head has no sigmoid, modules.py:119).

Tolerances: fp32 mode rtol 1e-3 on outputs (absolute floors at each check), parameter gradients max-error 5e-3 of the
gradient's max; bf16 mode: colours |err| <= 2e-2 (mean <= 3e-3), parameter gradients relative L2 <= 6e-2.
"""
import pytest
import torch

from oracle import port
from tests.util import loss_cfg, make_case, product_cfg, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from mli_nerf_b200 import _lib
    _lib.load()
    assert torch.cuda.is_available() and _lib.device_ok(), "these tests need a B200 (sm_100a)"
    return _lib


def cu(t):
    return t.contiguous().cuda()


def _l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def t22_case():
    """R = 64 rays at T = 2^22 with the oracle's forward + backward (w_curvature = 0: the curvature seed sign(laplacian)
    flips on fp32 noise, see test_gpu_parity.py) -- computed once for both precision modes."""
    case = make_case(R=64, log2_T=22, progress=0.5, miss_rays=4)
    ocfg = port.PathConfig(**{**case["ocfg"].__dict__, "w_curvature": 0.0})
    pp = {k: v.clone().requires_grad_(True) for k, v in case["params"].items()}
    out_ref = port.render_rays(pp, ocfg, case["center"], case["ray_unit"], case["light"], rands=case["rands"],
                               training=True, progress=case["progress"], keep=True)
    total, losses, _ = port.total_loss(ocfg, out_ref, case["targets"])
    total.backward()
    grads = {k: v.grad for k, v in pp.items()}
    out_ref = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out_ref.items()}
    return dict(case=case, ocfg=ocfg, out_ref=out_ref, total=float(total), grads=grads)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_t22_forward_backward_vs_oracle(lib, t22_case, prec):
    from mli_nerf_b200.engine import RenderEngine
    case, ocfg, out_ref, g_ref = t22_case["case"], t22_case["ocfg"], t22_case["out_ref"], t22_case["grads"]
    R, N = 64, 128
    eng = RenderEngine(product_cfg(ocfg, precision=1 if prec == "bf16" else 0))
    lv = eng.grid.level
    assert [int(lv[l].hashed) for l in range(16)] == [0] * 6 + [1] * 10, "T=22: levels 0-5 dense, 6-15 hashed"
    p = {k: cu(v) for k, v in case["params"].items()}
    eng.pack_weights(p)
    c, r, l = cu(case["center"][0]), cu(case["ray_unit"][0]), cu(case["light"][0])
    near, far, outside = eng.bounds(c, r)
    assert torch.equal(outside.cpu().bool(), out_ref["outside"][0, :, 0])
    # the sampling-side query (encode + SDF trunk only) on the coarse samples
    d_ref = out_ref["dists"][0, :, :, 0]
    pts = case["center"][..., None, :] + case["ray_unit"][..., None, :] * out_ref["dists"]
    sdf_q = eng.sdf_query(p["neural_sdf.tcnn_encoding.params"], c, r, cu(d_ref), N, N)
    with torch.no_grad():
        sdf_ref = port.sdf_only(case["params"], ocfg, pts)[0, :, :, 0]
    assert torch.allclose(sdf_q.cpu().view(R, N), sdf_ref, rtol=1e-3, atol=1e-5)
    # the render, fed the oracle's distances
    res, ctx = eng.forward(p, c, r, l, cu(d_ref), near, far, outside, True, case["progress"])
    inside = ~out_ref["outside"][0, :, 0]
    assert torch.allclose(res["sdf"].cpu()[:R * N].view(R, N), out_ref["sdfs"][0, :, :, 0], rtol=1e-3, atol=1e-5)
    g_got, g_want = res["gradients"].cpu().view(R, N, 3)[inside], out_ref["gradients"][0][inside]
    assert _l2(g_got, g_want) < (2e-3 if prec == "fp32" else 5e-4), _l2(g_got, g_want)
    out = res["out"].cpu()
    for k, (a, b) in dict(rgb=(0, 3), o_r=(3, 6), o_s=(6, 7), o_re=(7, 10)).items():
        if prec == "fp32":
            assert torch.allclose(out[:, a:b], out_ref[k][0], rtol=1e-3, atol=2e-5), k
        else:
            err = (out[:, a:b] - out_ref[k][0]).abs()
            assert float(err.max()) < 2e-2 and float(err.mean()) < 3e-3, (k, float(err.max()), float(err.mean()))
    tg = {k: cu(v[0]) for k, v in case["targets"].items()}
    losses, d_out, d_grad, d_hess = eng.losses(loss_cfg(ocfg), res["out"], res["gradients"], res["hessians"], outside, tg)
    tol_l = 1e-3 if prec == "fp32" else 2e-2
    assert abs(float(losses[0]) - t22_case["total"]) < tol_l * abs(t22_case["total"]) + 1e-5
    grads = eng.backward(p, ctx, d_out, d_grad, d_hess, None)
    # hash-table gradient: dense (stride-indexed) levels and hashed levels separately
    F = ocfg.feat_per_level
    split = int(lv[6].offset) * F
    tg_got = grads["neural_sdf.tcnn_encoding.params"].cpu().view(-1)
    tg_ref = g_ref["neural_sdf.tcnn_encoding.params"].view(-1)
    tol_t = 5e-3 if prec == "fp32" else 6e-2
    e_dense, e_hash = _l2(tg_got[:split], tg_ref[:split]), _l2(tg_got[split:], tg_ref[split:])
    print(f"T=22 [{prec}] table gradient rel-L2 error: dense levels 0-5 {e_dense:.2e}, hashed levels 6-15 {e_hash:.2e}")
    assert float(tg_ref[:split].abs().max()) > 0 and float(tg_ref[split:].abs().max()) > 0
    assert e_dense < tol_t and e_hash < tol_t, (e_dense, e_hash)
    # the zero pattern must agree too: entries no sample touches stay zero (an entry whose oracle contributions cancel
    # to exactly 0.0 may hold rounding noise).  NOTE: no tensor comparison inside an `assert` -- pytest's failure report
    # would iterate over the 365 M-element operands.
    extra = (tg_got != 0) & (tg_ref == 0)
    n_extra = int(extra.sum())
    max_extra = float(tg_got[extra].abs().max()) if n_extra else 0.0
    ref_max = float(tg_ref.abs().max())
    n_ref = int((tg_ref != 0).sum())
    print(f"T=22 [{prec}] entries non-zero only in the product: {n_extra} of {n_ref} touched, max |g| {max_extra:.2e} "
          f"(max |g_ref| {ref_max:.2e})")
    assert n_extra <= 1e-3 * n_ref and max_extra <= 1e-4 * ref_max, (n_extra, n_ref, max_extra, ref_max)
    worst = {}
    for k, v in g_ref.items():
        if k == "neural_sdf.tcnn_encoding.params":
            continue
        got = grads[k].cpu().view_as(v)
        worst[k] = rel_err(got, v) if prec == "fp32" else _l2(got, v)
    bad = {k: e for k, e in worst.items() if e > (5e-3 if prec == "fp32" else 6e-2)}
    assert not bad, bad


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_t22_model_fused_step_vs_oracle(lib, prec):
    """The exact call bench.py times -- Model.fused_train_step at dict_size 22 (own rays, own sampling, in-kernel losses,
    full backward) -- against the oracle pipeline on the same frame."""
    from mli_nerf_b200 import config
    from mli_nerf_b200.losses import loss_cfg_from_trainer
    from mli_nerf_b200.model import Model
    R = 96
    ocfg = port.PathConfig(log2_hashmap_size=22)
    params = port.init_params(ocfg, seed=0, generic=True, table_scale=5e-3)
    pose = torch.tensor([[[1, 0, 0, 0.0], [0, -1, 0, 0.0], [0, 0, -1, 3.0]]], dtype=torch.float32)
    intr = torch.tensor([[[711.0, 0, 256], [0, 711.0, 256], [0, 0, 1]]])
    pose_light = torch.tensor([[[1, 0, 0, 1.0], [0, 1, 0, -2.0], [0, 0, 1, 3.0]]], dtype=torch.float32)
    ray_idx = torch.randperm(512 * 512, generator=torch.Generator().manual_seed(0))[:R][None]
    tg = port.synthetic_targets(R)
    data = dict(pose=pose, intr=intr, pose_light=pose_light, ray_idx=ray_idx, **tg)
    c, ray, l = port.rays_from_pose(pose, intr, pose_light, (512, 512), ray_idx)
    pp = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref = port.render_rays(pp, ocfg, c, torch.nn.functional.normalize(ray, dim=-1), l, rands=None, training=True,
                           progress=0.5)
    total, _, _ = port.total_loss(ocfg, ref, tg)
    total.backward()
    cfg = config.experiment("syn_hotdog_b", dict_size=22, rand_rays=R)
    cfg.model.render.stratified = False
    cfg.model.mli_precision = prec
    model = Model(cfg.model, cfg.data)
    model.load_state_dict(params)
    model = model.cuda().train()
    model.progress = 0.5
    losses = model.fused_train_step({k: cu(v) for k, v in data.items()}, loss_cfg_from_trainer(cfg.trainer))
    torch.cuda.synchronize()
    tol = 2e-2 if prec == "fp32" else 6e-2  # independently sampled distances
    assert abs(float(losses[0]) - float(total)) < tol * abs(float(total)) + 1e-4, (float(losses[0]), float(total))
    for name in ("neural_rgb.mlp.linears.4.weight_v", "neural_rgb.mlp_r.linears.0.weight_v"):
        got, want = dict(model.named_parameters())[name].grad.cpu(), pp[name].grad
        assert float((got - want).abs().max()) < tol * float(want.abs().max()) + 1e-7, name
    tgrad = model.neural_sdf.tcnn_encoding.params.grad
    assert tgrad.numel() == ocfg.n_table_params() and bool(torch.isfinite(tgrad).all())
    # touched-entry pattern of the dense levels: identical sample positions on most rays -> nearly the same support
    split = int(model.engine.grid.level[6].offset) * 8
    nz_got, nz_ref = tgrad[:split].cpu() != 0, pp["neural_sdf.tcnn_encoding.params"].grad[:split] != 0
    assert float((nz_got & nz_ref).sum()) > 0.9 * float(nz_ref.sum())


MODE_COLS = {"r_s": dict(rgb=(0, 3), o_r=(3, 6), o_s=(6, 9)), "rgb_r": dict(rgb=(0, 3), o_r=(3, 6), o_s=(6, 9)),
             "r_s_re": dict(rgb=(0, 3), o_r=(3, 6), o_s=(6, 9), o_re=(9, 12))}


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("mode", ["r_s", "r_s_re", "rgb_r"])
def test_other_network_modes_vs_oracle(lib, mode, prec):
    """Heads, merge and backward of the three network modes that no shipped config selects."""
    from mli_nerf_b200.engine import RenderEngine
    case = make_case(R=192, mode=mode, progress=0.5)
    ocfg = port.PathConfig(**{**case["ocfg"].__dict__, "w_curvature": 0.0})
    pp = {k: v.clone().requires_grad_(True) for k, v in case["params"].items()}
    out_ref = port.render_rays(pp, ocfg, case["center"], case["ray_unit"], case["light"], rands=case["rands"],
                               training=True, progress=case["progress"], keep=True)
    total_ref, losses_ref, _ = port.total_loss(ocfg, out_ref, case["targets"])
    total_ref.backward()
    assert "intrinsic" in losses_ref and (("regularize_re" in losses_ref) == (mode == "r_s_re"))
    eng = RenderEngine(product_cfg(ocfg, precision=1 if prec == "bf16" else 0))
    p = {k: cu(v) for k, v in case["params"].items()}
    eng.pack_weights(p)
    c, r, l = cu(case["center"][0]), cu(case["ray_unit"][0]), cu(case["light"][0])
    near, far, outside = eng.bounds(c, r)
    res, ctx = eng.forward(p, c, r, l, cu(out_ref["dists"][0, :, :, 0]), near, far, outside, True, case["progress"])
    out = res["out"].cpu()
    assert out.shape[1] == max(b for _, b in MODE_COLS[mode].values())
    for k, (a, b) in MODE_COLS[mode].items():
        want = out_ref[k][0].detach()
        if prec == "fp32":
            # rgb_r's o_s = rgb / o_r divides by an accumulated colour that can be small: relative bound only
            assert torch.allclose(out[:, a:b], want, rtol=1e-3, atol=2e-5 * max(1.0, float(want.abs().max()))), k
        else:
            err = (out[:, a:b] - want).abs() / (1.0 + want.abs())
            assert float(err.max()) < 2e-2 and float(err.mean()) < 3e-3, (k, float(err.max()), float(err.mean()))
    if mode == "r_s":  # the un-squashed shading head really is unbounded
        s = out_ref["s_o_s"].detach()
        assert float(s.min()) < 0.0 or float(s.max()) > 1.0
    tg = {k: cu(v[0]) for k, v in case["targets"].items()}
    losses, d_out, d_grad, d_hess = eng.losses(loss_cfg(ocfg), res["out"], res["gradients"], res["hessians"], outside, tg)
    lc = losses.cpu()
    tol = 1e-3 if prec == "fp32" else 2e-2
    assert abs(float(lc[0]) - float(total_ref)) < tol * abs(float(total_ref)) + 1e-5
    for i, k in ((1, "render"), (2, "eikonal"), (4, "intrinsic"), (5, "regularize_re")):
        if k in losses_ref:
            assert abs(float(lc[i]) - float(losses_ref[k])) < tol * abs(float(losses_ref[k])) + 1e-6, k
    grads = eng.backward(p, ctx, d_out, d_grad, d_hess, None)
    worst = {}
    for k, v in pp.items():
        assert k in grads, k
        got = grads[k].cpu().view_as(v.grad)
        worst[k] = rel_err(got, v.grad) if prec == "fp32" else _l2(got, v.grad)
    bad = {k: e for k, e in worst.items() if e > (5e-3 if prec == "fp32" else 6e-2)}
    assert not bad, bad


@pytest.mark.parametrize("mode", ["r_s", "r_s_re", "rgb_r"])
def test_other_network_modes_model_outputs(lib, mode):
    """Drop-in Model for those modes: output dictionary keys / shapes of NeuralLumen/model.py:266-310 in train and eval."""
    from mli_nerf_b200 import config
    from mli_nerf_b200.model import Model
    cfg = config.experiment("syn_hotdog_b", dict_size=14, rand_rays=64)
    cfg.model.object.rgb.network_mode = mode
    cfg.model.object.rgb.shading_dim = 3
    cfg.model.mli_precision = "bf16"
    ocfg = port.PathConfig(log2_hashmap_size=14, network_mode=mode)
    params = port.init_params(ocfg, seed=0, generic=True, table_scale=5e-3)
    model = Model(cfg.model, cfg.data)
    model.load_state_dict(params, strict=True)
    model = model.cuda()
    center, ray_unit, light = port.synthetic_rays(64, seed=3)
    for training in (True, False):
        model.train(training)
        with torch.no_grad():
            out = model.render_rays_lumen(cu(center), cu(ray_unit), cu(light), stratified=False)
            ref = port.render_rays(params, ocfg, center, ray_unit, light, rands=None, training=training, progress=1.0)
        assert set(k for k, v in out.items() if v is not None and not k.startswith("_")) == \
            set(k for k, v in ref.items() if v is not None)
        for k in MODE_COLS[mode]:
            assert out[k].shape == ref[k].shape, k


def test_light_visibility_box_bound_and_training_mode(lib):
    """visibility_bounding_type 'box' uses the DATA box like the reference (NeuralLumen/model.py:188-191; the oracle
    branch is pinned on the live reference by test_oracle_vs_reference.py::test_light_visibility_box_bound_matches_reference),
    and training mode produces the visibility outputs + the composited gradient too (model.py:325-334, 369-372)."""
    from mli_nerf_b200 import config
    from mli_nerf_b200.model import Model
    cfg = config.experiment("rene_savannah_b", dict_size=14, rand_rays=64)
    lv = cfg.model.light_visibility
    lv.enabled, lv.camera_ray_type, lv.visibility_bounding_type = True, "sphere_tracing", "box"
    lv.visibility_bounding_box_aabb = [-0.3, -0.21, -0.18, 0.3, 0.21, 0.18]  # present in the YAML, NOT what is used
    aabb = tuple(float(v) for v in cfg.data.bounding_box_aabb)
    ocfg = port.PathConfig(log2_hashmap_size=14, bounding="box", aabb=aabb, white_background=False)
    p = port.init_params(ocfg, seed=5, generic=False)
    p["neural_sdf.mlp.linear_sdf.bias"] = torch.tensor([-0.3])
    model = Model(cfg.model, cfg.data)
    model.load_state_dict(p, strict=True)
    model = model.cuda().eval()
    R = 64
    center, ray_unit, light = port.synthetic_rays(R, seed=6)
    center = center * 0.5
    g = torch.Generator().manual_seed(1)
    ray_unit = torch.nn.functional.normalize(-center + 0.05 * torch.randn(center.shape, generator=g), dim=-1)
    light = light * 0.3
    with torch.no_grad():
        out = model.render_rays_lumen(cu(center), cu(ray_unit), cu(light), stratified=False)
        ref = port.render_rays(p, ocfg, center, ray_unit, light, rands=None, training=False, progress=1.0, keep=True)
        near, far, _ = port.dist_bounds(ocfg, center, ray_unit)
        blend = port.composite(ref["dists"], ref["weights"])
        vis, nxl, idist, imask = port.light_visibility(p, ocfg, center, ray_unit, light, near, far, blend, ref["gradient"],
                                                       "sphere_tracing", aabb=aabb)
    assert 0 < int(imask.sum()) and 0 < int(vis.sum()) < R
    assert float((out["inter_mask"].cpu() == imask).float().mean()) > 0.97
    ok = (out["inter_dist"].cpu() - idist).abs() < 1e-4 * (1 + idist.abs())
    assert float(ok.float().mean()) > 0.97
    assert float((out["visibility"].cpu() == vis).float().mean()) > 0.95
    assert float(((out["normal_x_light"].cpu() - nxl).abs() < 2e-3).float().mean()) > 0.95
    model.train()
    with torch.no_grad():
        out_t = model.render_rays_lumen(cu(center), cu(ray_unit), cu(light), stratified=False)
    for k in ("visibility", "normal_x_light", "pseudo_shading", "inter_dist", "inter_mask"):
        assert out_t[k].shape == (1, R, 1), k
    assert out_t["gradient"].shape == (1, R, 3) and out_t["opacity"] is None and out_t["hessians"] is not None
    assert float((out_t["inter_mask"].cpu() == imask).float().mean()) > 0.97


def test_sharded_table_adamw_single_rank_group(lib):
    """The rank-owned optimizer path of the reduce-scatter exchange on ONE GPU (a world-size-1 NCCL group): slab shards
    at table offsets through mli_adamw_step_batch + the in-place parameter all-gather == torch.optim.AdamW.  (The
    N-rank run of the same code is tests/multi/train_step_check.py, which needs >= 2 GPUs.)"""
    import torch.distributed as dist
    from mli_nerf_b200.optim import ShardedTableAdamW
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29771", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        torch.manual_seed(0)
        n = 1 << 22
        table = torch.nn.Parameter(torch.randn(n, device="cuda") * 0.01)
        small = torch.nn.Parameter(torch.randn(256, 131, device="cuda"))
        ref_p = [table.detach().clone().requires_grad_(True), small.detach().clone().requires_grad_(True)]
        ref = torch.optim.AdamW(ref_p, lr=1e-3, weight_decay=1e-2)

        class Red:
            world, rank = 1, 0
        red = Red()
        opt = ShardedTableAdamW(red, [table, small], lr=1e-3, weight_decay=1e-2)
        cuts = [0, n // 8, n // 2, n]
        for it in range(3):
            g = torch.randn(n, device="cuda") * (torch.rand(n, device="cuda") > 0.8)  # mostly-zero gradient, like the table's
            gs = torch.randn_like(small)
            table.grad, small.grad = torch.full_like(table, float("nan")), gs.clone()  # the dense .grad must not be read
            red._last_shards = [(a, b, g[a:b].clone()) for a, b in zip(cuts[:-1], cuts[1:])]
            ref_p[0].grad, ref_p[1].grad = g.clone(), gs.clone()
            opt.step()
            ref.step()
        assert torch.allclose(table, ref_p[0], rtol=1e-5, atol=1e-7), float((table - ref_p[0]).abs().max())
        assert torch.allclose(small, ref_p[1], rtol=1e-5, atol=1e-6)
        m, v = opt.gather_state()
        # beta1 * m and (1 - beta1) * g can cancel: absolute floor = a few ulps of the larger term
        assert torch.allclose(m, ref.state[ref_p[0]]["exp_avg"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(v, ref.state[ref_p[0]]["exp_avg_sq"], rtol=1e-5, atol=1e-8)
    finally:
        if own_group:
            dist.destroy_process_group()


def test_reduce_slots_mean_of_staged_shards(lib):
    """mli_reduce_slots, the summation step of the copy-engine exchange, emulated on one GPU: own shard + (W-1) staged
    peer shards -> mean, in place, for ragged sizes."""
    torch.manual_seed(1)
    for W, n in ((2, 1000), (4, 8 * 1_000_003 // 4), (8, 4)):
        slot = (n + 3) // 4 * 4
        own = torch.randn(n, device="cuda")
        stage = torch.randn((W - 1) * slot, device="cuda")
        want = (own + sum(stage[k * slot:k * slot + n] for k in range(W - 1))) / W
        lib.call("mli_reduce_slots", own, stage, W - 1, slot, n, 1.0 / W)
        assert torch.allclose(own, want, rtol=1e-6, atol=1e-7)
