"""Shared helpers for the parity tests: build matching (oracle cfg, product cfg, params) triples."""
import torch

from oracle import port


def make_case(R=256, log2_T=14, mode="rgb_r_s", bounding="unit_sphere", taps=4, white=True, seed=0, table_scale=5e-3,
              progress=0.5, miss_rays=8):
    """Oracle PathConfig + a parameter set that exercises every input + rays (some of which miss the bounds)."""
    aabb = (-0.66, -0.516, -0.18, 0.66, 0.42, 0.3) if bounding == "box" else None
    ocfg = port.PathConfig(log2_hashmap_size=log2_T, network_mode=mode, bounding=bounding, aabb=aabb, taps=taps,
                           white_background=white,
                           normal_eps=1.0 / 2048)
    params = port.init_params(ocfg, seed=seed, generic=True, table_scale=table_scale)
    center, ray_unit, light = port.synthetic_rays(R, seed=seed + 1)
    if bounding == "box":
        center = center * 0.5  # closer, so most rays hit the small box
        ray_unit = torch.nn.functional.normalize(-center + 0.15 * torch.randn_like(center), dim=-1)
    if miss_rays:
        g = torch.Generator().manual_seed(seed + 7)
        ray_unit[0, :miss_rays] = torch.nn.functional.normalize(torch.randn(miss_rays, 3, generator=g), dim=-1)
    g = torch.Generator().manual_seed(seed + 3)
    rands = torch.rand(1, R, ocfg.coarse, 1, generator=g)
    targets = port.synthetic_targets(R, seed=seed + 2)
    return dict(ocfg=ocfg, params=params, center=center, ray_unit=ray_unit, light=light, rands=rands, targets=targets,
                progress=progress)


def product_cfg(ocfg, precision=0):
    from mli_nerf_b200.engine import PathCfg
    return PathCfg(n_levels=ocfg.n_levels, feat_per_level=ocfg.feat_per_level, log2_hashmap_size=ocfg.log2_hashmap_size,
                   min_logres=ocfg.min_logres, max_logres=ocfg.max_logres, vol_range=ocfg.vol_range, hidden=ocfg.hidden,
                   taps=ocfg.taps, coarse=ocfg.coarse, fine=ocfg.fine, hierarchy=ocfg.hierarchy,
                   network_mode=ocfg.network_mode, white_background=ocfg.white_background, anneal_end=ocfg.anneal_end,
                   outside_val=ocfg.outside_val, bounding=ocfg.bounding, aabb=ocfg.aabb, c2f_enabled=ocfg.c2f_enabled,
                   precision=precision)


def loss_cfg(ocfg, has_intrinsic=True):
    from mli_nerf_b200 import _lib
    c = _lib.LossCfg()
    c.w_render, c.w_eikonal, c.w_curvature = ocfg.w_render, ocfg.w_eikonal, ocfg.w_curvature
    c.w_intrinsic, c.w_regularize_re = ocfg.w_intrinsic, ocfg.w_regularize_re
    c.range_sha[0], c.range_sha[1] = ocfg.range_shading
    c.range_vis[0], c.range_vis[1] = ocfg.range_visibility
    c.factor_ref, c.factor_sha = ocfg.factor_ref, ocfg.factor_sha
    c.factor_negative, c.factor_positive, c.exponent_positive = (ocfg.factor_negative, ocfg.factor_positive,
                                                                 ocfg.exponent_positive)
    c.has_intrinsic = int(has_intrinsic)
    return c


def rel_err(a, b):
    """max |a-b| / max|b| -- scale-aware error for gradient tensors."""
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
