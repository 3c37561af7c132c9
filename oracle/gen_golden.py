"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* from the UNMODIFIED reference (run in the build container):

    python -m oracle.gen_golden

Everything here comes from /root/reference code executed on CPU (imported through oracle/ref_import.py) except the
hash-grid encoding, which the reference takes from the un-vendored tiny-cuda-nn and which is therefore supplied by
oracle/torch_hashgrid.py ("parity unpinned" at that one boundary).  The fixtures are small (a few hundred kB) so that
they can be committed; weights are NOT stored: they are regenerated from a seed by oracle.port.init_params and guarded
by a checksum.
"""
import hashlib
import os
import warnings

import numpy as np
import torch

from oracle import port, ref_import

warnings.filterwarnings("ignore")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def params_digest(p):
    h = hashlib.sha256()
    for k in sorted(p):
        h.update(k.encode())
        h.update(p[k].detach().numpy().tobytes())
    return h.hexdigest()


def strided(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].numpy().copy()


def gen_sampling():
    """nerf_util.sample_dists_from_pdf on fixed weights: idx/low/high/cdf/dists are the integer-exact KAT."""
    ref = ref_import.load()
    out = {}
    for n_w in (63, 79, 95, 111):
        g = torch.Generator().manual_seed(n_w)
        w = torch.rand(1, 200, n_w, generator=g) * (torch.rand(1, 200, n_w, generator=g) > 0.5)
        w[0, :10] = 0.0
        w[0, 10:20, 7:] = 0.0
        w[0, 20:30] *= 1e-30
        bins = (torch.rand(1, 200, n_w + 1, 1, generator=g).sort(dim=2).values * 3 + 0.5)
        # replicate the function body's intermediates by calling it and re-deriving idx the way it does
        d = ref.nerf_util.sample_dists_from_pdf(bins, w, intvs_fine=16)
        pdf = torch.nn.functional.normalize(w, p=1, dim=-1)
        cdf = torch.cat([torch.zeros_like(pdf[..., :1]), pdf.cumsum(dim=-1)], dim=-1)
        grid = torch.linspace(0, 1, 17)
        unif = (0.5 * (grid[:-1] + grid[1:])).repeat(1, 200, 1)
        idx = torch.searchsorted(cdf, unif, right=True)
        out[f"w{n_w}"] = w[0].numpy()
        out[f"bins{n_w}"] = bins[0, :, :, 0].numpy()
        out[f"cdf{n_w}"] = cdf[0].numpy()
        out[f"idx{n_w}"] = idx[0].numpy().astype(np.int32)
        out[f"dists{n_w}"] = d[0, :, :, 0].numpy()
    np.savez_compressed(os.path.join(OUT, "sampling_kat.npz"), **out)


def gen_hash_index():
    """Corner rows of the tcnn HashGrid restatement (oracle/torch_hashgrid.py) -- KAT for the C oracle and the CUDA
    kernels; NOT a reference output (tcnn is not in /root/reference)."""
    from oracle.torch_hashgrid import corner_indices, level_table
    import math
    pls = math.exp((math.log(2048) - math.log(32)) / 15)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(512, 3, generator=g)
    x[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [-0.25, 0.5, 1.5], [1.0, 0.0, 0.5], [0.5, 0.5, 0.5],
                          [-1e-7, 1 + 1e-7, 0.3], [0.999999, 1e-8, 0.25], [2.5, -3.0, 0.1]])
    out = {"x": x.numpy()}
    for T in (14, 22):
        lv, n = level_table(16, T, 32, pls)
        out[f"levels{T}"] = np.array([[l["scale"], l["res"], l["size"], l["offset"], int(l["hashed"])] for l in lv],
                                     dtype=np.float64)
        rows = [(corner_indices(x, lv[level])[0] + lv[level]["offset"]).numpy() for level in range(16)]
        out[f"rows{T}"] = np.stack(rows).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "hash_index_kat.npz"), **out)


def gen_render(name, cfg_name, overrides, cfgkw, R, training, progress, seed):
    """Full render + losses + parameter gradients of the reference Model on seeded rays/weights."""
    cfg_ref = ref_import.load_config(cfg_name, overrides)
    ocfg = port.PathConfig(**cfgkw)
    p = port.init_params(ocfg, seed=seed, generic=True, table_scale=5e-3)
    model = ref_import.build_model(cfg_ref, progress=progress, training=training)
    model.load_state_dict(p, strict=True)
    center, ray_unit, light = port.synthetic_rays(R, seed=seed + 1)
    if ocfg.bounding == "box":
        center = center * 0.5
        ray_unit = torch.nn.functional.normalize(-center + 0.15 * torch.randn(center.shape,
                                                 generator=torch.Generator().manual_seed(seed + 9)), dim=-1)
    ray_unit[0, :4] = torch.nn.functional.normalize(torch.randn(4, 3, generator=torch.Generator().manual_seed(seed + 7)),
                                                    dim=-1)
    g = torch.Generator().manual_seed(seed + 3)
    rands = torch.rand(1, R, 64, 1, generator=g)
    # the reference draws its stratified samples with torch.rand inside sample_dists: patch the global RNG stream
    real_rand = torch.rand
    torch.rand = lambda *a, **k: rands.clone() if tuple(a) == (1, R, 64, 1) else real_rand(*a, **k)
    try:
        out = model.render_rays_lumen(center, ray_unit, light, stratified=True)
    finally:
        torch.rand = real_rand
    store = dict(center=center.numpy(), ray_unit=ray_unit.numpy(), light=light.numpy(), rands=rands.numpy(),
                 params_sha256=np.frombuffer(params_digest(p).encode(), dtype=np.uint8), seed=np.array(seed),
                 progress=np.array(progress), training=np.array(int(training)))
    for k, v in out.items():
        if v is not None:
            store["out_" + k] = v.detach().numpy()
    if training:
        ref = ref_import.load()
        from projects.neuralangelo.utils.misc import curvature_loss, eikonal_loss
        tg = port.synthetic_targets(R, seed=seed + 2)
        losses = dict(render=torch.nn.functional.l1_loss(out["rgb"], tg["image_sampled"]) * 3,
                      eikonal=eikonal_loss(out["gradients"], outside=out["outside"]),
                      curvature=curvature_loss(out["hessians"], outside=out["outside"]))
        wts = dict(render=1.0, eikonal=0.1, curvature=5e-4)
        if "o_re" in out:
            losses["intrinsic"] = ref.lumen_utils.intrinsic_loss(
                out["o_r"], out["o_s"], tg["pseudo_ref_sampled"], tg["pseudo_sha_sampled"],
                tg["pseudo_visibility_certainty_sampled"], weight_map_range_shading=(0.0, 1.0),
                weight_map_range_visibility=(0.0, 1.0), factor_ref=1.0, factor_sha=1.0)
            losses["regularize_re"] = ref.lumen_utils.regularize_re_loss(out["o_re"], factor_negative=10.0,
                                                                        factor_positive=1.0, exponent_positive=1.0)
            wts.update(intrinsic=1.0, regularize_re=1.0)
        total = sum(wts[k] * v for k, v in losses.items())  # imaginaire/trainers/base.py:534-544
        model.zero_grad()
        total.backward()
        store["loss_total"] = total.detach().numpy()
        for k, v in losses.items():
            store["loss_" + k] = v.detach().numpy()
        for k, v in model.named_parameters():
            store["gnorm_" + k] = v.grad.norm().numpy()
            store["grad_" + k] = v.grad.numpy() if v.grad.numel() <= 4096 else strided(v.grad)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)


def gen_light_visibility(name, R=64, seed=3):
    """Model.get_light_visibility of the reference (stage-a export of pseudo shading labels, NeuralLumen/model.py:133-180)
    reached through render_rays_lumen in eval mode; geometric init (sphere tracing is a fixed-point iteration that is only
    stable for |grad sdf| ~ 1)."""
    store = {}
    ocfg = port.PathConfig(log2_hashmap_size=14)
    p = port.init_params(ocfg, seed=seed, generic=False)
    center, ray_unit, light = port.synthetic_rays(R, seed=seed + 1)
    for ray_type in ("blend_z_sphere_tracing", "sphere_tracing", "blend_z"):
        over = {"model.object.sdf.encoding.hashgrid.dict_size": 14, "model.light_visibility.enabled": True,
                "model.light_visibility.camera_ray_type": ray_type}
        cfg_ref = ref_import.load_config("syn_hotdog_b", over)
        model = ref_import.build_model(cfg_ref, progress=1.0, training=False)
        model.load_state_dict(p, strict=True)
        with torch.no_grad():
            out = model.render_rays_lumen(center, ray_unit, light, stratified=False)
            blend = (out["dists"] * out["weights"]).sum(dim=2)  # render.composite(dists, weights), model.py:138
        for k in ("visibility", "normal_x_light", "inter_dist", "inter_mask"):
            store[f"{ray_type}_{k}"] = out[k].numpy()
        store["radius"] = np.array(float(cfg_ref.model.light_visibility.visibility_sphere_radius))
    store.update(center=center.numpy(), ray_unit=ray_unit.numpy(), light=light.numpy(), blend=blend.numpy(),
                 gradient=out["gradient"].numpy(), rgb=out["rgb"].numpy(),
                 params_sha256=np.frombuffer(params_digest(p).encode(), dtype=np.uint8), seed=np.array(seed))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)


def gen_inference(name, H=12, W=14, seed=3):
    """Model.inference of the reference (NeuralLumen/model.py:60-83): full-image eval render in chunks, depth =
    composited distance / |ray|, camera-space normal map, [B,HW,C] -> [B,C,H,W] maps."""
    over = {"model.object.sdf.encoding.hashgrid.dict_size": 14, "data.val.image_size": [H, W],
            "model.render.rand_rays_val": 50}  # 168 rays in chunks of 50: ragged last chunk
    cfg_ref = ref_import.load_config("syn_hotdog_b", over)
    ocfg = port.PathConfig(log2_hashmap_size=14)
    p = port.init_params(ocfg, seed=seed, generic=False)
    model = ref_import.build_model(cfg_ref, progress=1.0, training=False)
    model.load_state_dict(p, strict=True)
    pose = torch.tensor([[[0.96, 0.0, 0.28, 0.05], [0.0, -1.0, 0.0, 0.02], [0.28, 0.0, -0.96, 3.0]]], dtype=torch.float32)
    intr = torch.tensor([[[17.0, 0.0, W / 2], [0.0, 17.0, H / 2], [0.0, 0.0, 1.0]]])
    pose_light = torch.tensor([[[1.0, 0, 0, 1.0], [0, 1.0, 0, -2.0], [0, 0, 1.0, 3.0]]], dtype=torch.float32)
    data = dict(pose=pose, intr=intr, pose_light=pose_light, idx=torch.zeros(1, dtype=torch.long))
    out = model.inference(data)
    store = dict(pose=pose.numpy(), intr=intr.numpy(), pose_light=pose_light.numpy(), image_size=np.array([H, W]),
                 params_sha256=np.frombuffer(params_digest(p).encode(), dtype=np.uint8), seed=np.array(seed))
    for k in ("rgb_map", "opacity_map", "depth_map", "normal_map", "o_r_map", "o_s_map", "o_re_map", "outside"):
        store[k] = out[k].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_light_visibility("light_visibility_hotdog_b")
    gen_inference("inference_hotdog_b")
    gen_sampling()
    gen_hash_index()
    t14 = {"model.object.sdf.encoding.hashgrid.dict_size": 14}
    gen_render("render_hotdog_b_train", "syn_hotdog_b", t14, dict(log2_hashmap_size=14), 48, True, 0.03, 0)
    gen_render("render_hotdog_b_eval", "syn_hotdog_b", t14, dict(log2_hashmap_size=14), 48, False, 1.0, 0)
    gen_render("render_rene_b_train", "rene_savannah_b", t14,
               dict(log2_hashmap_size=14, bounding="box", aabb=(-0.66, -0.516, -0.18, 0.66, 0.42, 0.3),
                    white_background=False), 48, True, 0.5, 10)
    print(sorted(os.listdir(OUT)), sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)))


if __name__ == "__main__":
    main()
