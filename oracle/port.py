"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of MLI-NeRF's per-ray render hot path.

This is the oracle that travels to the GPU box (the reference itself is Python that only exists under
/root/reference in the build container).  It is a from-scratch, functional restatement (tensors + a flat
state-dict, no nn.Modules); every function cites the reference lines it follows.  It is pinned against the
UNMODIFIED reference by tests/test_oracle_vs_reference.py (runs wherever /root/reference exists) and by the
golden vectors under tests/golden/ which were produced by the reference itself (oracle/gen_golden.py).

The hash-grid encoding is the exception: tcnn is not under /root/reference, see torch_hashgrid.py
("parity unpinned" at that boundary).

Citations are relative to /root/reference/.
"""
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from oracle.torch_hashgrid import corner_indices, level_table
from oracle.torch_hashgrid import hash_encode as _hash_encode


@dataclass
class PathConfig:
    """Resolved hyper-parameters of the hot path."""
    n_levels: int = 16
    feat_per_level: int = 8
    log2_hashmap_size: int = 22
    min_logres: int = 5
    max_logres: int = 11
    vol_range: Tuple[float, float] = (-2.0, 2.0)
    hidden: int = 256
    taps: int = 4
    coarse: int = 64
    fine: int = 16
    hierarchy: int = 4
    sh_levels: int = 3
    network_mode: str = "rgb_r_s"
    white_background: bool = True
    anneal_end: float = 0.1
    outside_val: float = 1000.0
    bounding: str = "unit_sphere"  # or "box"
    aabb: Optional[Tuple[float, ...]] = None
    c2f_enabled: bool = False
    active_levels: int = 16
    normal_eps: float = 1.0 / 2048
    # loss hyper-parameters (projects/NeuralLumen/configs/*_b.yaml: trainer.*)
    w_render: float = 1.0
    w_eikonal: float = 0.1
    w_curvature: float = 5e-4
    w_intrinsic: float = 1.0
    w_regularize_re: float = 1.0
    range_shading: Tuple[float, float] = (0.0, 1.0)
    range_visibility: Tuple[float, float] = (0.0, 1.0)
    factor_ref: float = 1.0
    factor_sha: float = 1.0
    factor_negative: float = 10.0
    factor_positive: float = 1.0
    exponent_positive: float = 1.0
    _levels: list = field(default=None, repr=False)

    @property
    def n_samples(self):
        return self.coarse + self.fine * self.hierarchy

    @property
    def growth_rate(self):  # projects/neuralangelo/utils/modules.py:38-41
        return float(math.exp((math.log(2 ** self.max_logres) - math.log(2 ** self.min_logres)) / (self.n_levels - 1)))

    def levels(self):
        if self._levels is None:
            self._levels, _ = level_table(self.n_levels, self.log2_hashmap_size, 2 ** self.min_logres,
                                          self.growth_rate)
        return self._levels

    def n_table_params(self):
        lv = self.levels()[-1]
        return (lv["offset"] + lv["size"]) * self.feat_per_level

    def resolutions(self):  # modules.py:51-54
        return [int(math.floor(2 ** self.min_logres * self.growth_rate ** lv)) + 1 for lv in range(self.n_levels)]


# ----------------------------------------------------------------------------------------------------------
# encoding + SDF network
# ----------------------------------------------------------------------------------------------------------
def wn_weight(p, prefix):
    """old-style torch.nn.utils.weight_norm(dim=0): W = g * v / ||v||_row  (mlp.py:42-44, nerf_util.py:177-178)."""
    v, g = p[prefix + ".weight_v"], p[prefix + ".weight_g"]
    return v * (g / v.norm(dim=1, keepdim=True))


def hash_encode(p, cfg, x01):
    """tcnn HashGrid stand-in, see torch_hashgrid.py; table = p['neural_sdf.tcnn_encoding.params']."""
    return _hash_encode(p["neural_sdf.tcnn_encoding.params"], x01, cfg.levels(), cfg.feat_per_level)


def sdf_encode(p, cfg, pts):
    """NeuralSDF.encode (projects/neuralangelo/utils/modules.py:76-95): normalise to [0,1], hash grid,
    coarse-to-fine mask (:110-113), concat xyz in front."""
    lo, hi = cfg.vol_range
    x01 = (pts - lo) / (hi - lo)
    enc = hash_encode(p, cfg, x01.reshape(-1, 3)).view(*pts.shape[:-1], -1)
    if cfg.c2f_enabled:
        mask = torch.zeros_like(enc)
        mask[..., :cfg.active_levels * cfg.feat_per_level] = 1
        enc = enc * mask
    return torch.cat([pts, enc], dim=-1)


def softplus100(x):
    return F.softplus(x, beta=100)


def sdf_network(p, cfg, pts, with_feat=True):
    """NeuralSDF.forward + MLPforNeuralSDF.forward (modules.py:68-74, mlp.py:55-69) for num_layers=1:
    h0 = softplus100(W0 x + b0); sdf = w_sdf . h0 + b_sdf (fed from the *input* of the last layer);
    feat = softplus100(W1 h0 + b1)."""
    x = sdf_encode(p, cfg, pts)
    h0 = softplus100(F.linear(x, wn_weight(p, "neural_sdf.mlp.linears.0"), p["neural_sdf.mlp.linears.0.bias"]))
    sdf = F.linear(h0, p["neural_sdf.mlp.linear_sdf.weight"], p["neural_sdf.mlp.linear_sdf.bias"])
    feat = None
    if with_feat:
        feat = softplus100(F.linear(h0, wn_weight(p, "neural_sdf.mlp.linears.1"), p["neural_sdf.mlp.linears.1.bias"]))
    return sdf, feat


def sdf_only(p, cfg, pts):
    return sdf_network(p, cfg, pts, with_feat=False)[0]


def sdf_gradients(p, cfg, pts, sdf, training):
    """NeuralSDF.compute_gradients, numerical mode (modules.py:131-177)."""
    if cfg.taps == 4:
        e = cfg.normal_eps / math.sqrt(3)
        ks = [torch.tensor(k, dtype=pts.dtype) for k in ([1, -1, -1], [-1, -1, 1], [-1, 1, -1], [1, 1, 1])]
        s = [sdf_only(p, cfg, pts + k * e) for k in ks]
        grad = (ks[0] * s[0] + ks[1] * s[1] + ks[2] * s[2] + ks[3] * s[3]) / (4.0 * e)
        hess = None
        if training:
            hxx = ((s[0] + s[1] + s[2] + s[3]) / 2.0 - 2 * sdf) / e ** 2
            hess = torch.cat([hxx, hxx, hxx], dim=-1) / 3.0
        return grad, hess
    if cfg.taps == 6:
        eps = cfg.normal_eps
        g, h = [], []
        for d in range(3):
            off = torch.zeros(3, dtype=pts.dtype)
            off[d] = eps
            sp, sn = sdf_only(p, cfg, pts + off), sdf_only(p, cfg, pts - off)
            g.append((sp - sn) / (2 * eps))
            if training:
                h.append((sp + sn - 2 * sdf) / (eps ** 2))
        return torch.cat(g, dim=-1), (torch.cat(h, dim=-1) if training else None)
    raise ValueError("Only support 4 or 6 taps.")


# ----------------------------------------------------------------------------------------------------------
# colour / intrinsic heads
# ----------------------------------------------------------------------------------------------------------
_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def sh_basis(d, levels=3):
    """Real SH basis, levels<=3 (projects/neuralangelo/utils/spherical_harmonics.py:47-70)."""
    x, y, z = d.unbind(-1)
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    v = [torch.full_like(x, _C0), -_C1 * y, _C1 * z, -_C1 * x,
         _C2[0] * xy, _C2[1] * yz, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy),
         _C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
         _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy), _C3[5] * z * (xx - yy),
         _C3[6] * x * (xx - 3 * yy)]
    return torch.stack(v[:(levels + 1) ** 2], dim=-1)


def head_mlp(p, prefix, x, n_layers=5):
    """MLPwithSkipConnection.forward, no skips, ReLU hidden (projects/nerf/utils/nerf_util.py:186-196)."""
    h = x
    for li in range(n_layers):
        h = F.linear(h, wn_weight(p, f"{prefix}.linears.{li}"), p[f"{prefix}.linears.{li}.bias"])
        if li != n_layers - 1:
            h = torch.relu(h)
    return h


def lumen_heads(p, cfg, pts, normals, rays_unit, feats, pts_light):
    """LumenRGB.forward (projects/NeuralLumen/utils/modules.py:106-174).  Returns dict of per-sample outputs."""
    view = sh_basis(rays_unit, cfg.sh_levels)
    light = sh_basis(pts_light, cfg.sh_levels)  # raw light *position*, not a direction (:108-109)
    full = torch.cat([pts, view, normals, feats, light], dim=-1)
    geo = torch.cat([pts, normals, feats], dim=-1)
    geo_l = torch.cat([pts, normals, feats, light], dim=-1)
    m = cfg.network_mode
    if m == "rgb_r_s":
        return dict(rgbs=head_mlp(p, "neural_rgb.mlp", full).sigmoid(),
                    o_r=head_mlp(p, "neural_rgb.mlp_r", geo).sigmoid(),
                    o_s=head_mlp(p, "neural_rgb.mlp_s", geo_l).sigmoid())
    if m == "rgb_r":
        return dict(rgbs=head_mlp(p, "neural_rgb.mlp", full).sigmoid(),
                    o_r=head_mlp(p, "neural_rgb.mlp_r", geo).sigmoid())
    if m == "r_s":
        return dict(o_r=head_mlp(p, "neural_rgb.mlp_r", geo).sigmoid(),
                    o_s=head_mlp(p, "neural_rgb.mlp_s", full))  # no sigmoid on o_s here (:119)
    if m == "r_s_re":
        return dict(o_r=head_mlp(p, "neural_rgb.mlp_r", geo).sigmoid(),
                    o_s=head_mlp(p, "neural_rgb.mlp_s", geo_l).sigmoid(),
                    o_re=head_mlp(p, "neural_rgb.mlp_re", full).sigmoid())
    if m in (None, "rgb"):
        return dict(rgbs=head_mlp(p, "neural_rgb.mlp", full).sigmoid())
    raise NotImplementedError(m)


# ----------------------------------------------------------------------------------------------------------
# rays, bounds, sampling
# ----------------------------------------------------------------------------------------------------------
def rays_from_pose(pose, intr, pose_light, image_size, ray_idx):
    """camera.get_center_and_ray + slice_by_ray_idx + get_center (projects/nerf/utils/camera.py:283-311,263-266,
    46-52; nerf_util.py:127-131; projects/NeuralLumen/utils/utils.py:61-79).  pose = world->camera [R|t]."""
    H, W = image_size
    B = pose.shape[0]
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5,
                            indexing="ij")
    pix = torch.stack([xx, yy, torch.ones_like(xx)], dim=-1).view(1, -1, 3).repeat(B, 1, 1)
    cam = pix @ intr.inverse().transpose(-1, -2)

    def to_world(X, P):
        R, t = P[..., :3], P[..., 3:]
        Rinv = R.transpose(-1, -2)
        tinv = (-Rinv @ t)[..., 0]
        Pinv = torch.cat([Rinv, tinv[..., None]], dim=-1)
        Xh = torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)
        return Xh @ Pinv.transpose(-1, -2)

    grid = to_world(cam, pose)
    center = to_world(torch.zeros_like(cam), pose)
    light = to_world(torch.zeros_like(cam), pose_light)
    bi = torch.arange(B)[:, None].expand_as(ray_idx)
    ray = (grid - center)[bi, ray_idx]
    return center[bi, ray_idx], ray, light[bi, ray_idx]


def dist_bounds(cfg, center, ray_unit):
    """Model.get_dist_bounds (projects/neuralangelo/model.py:420-430) with intersect_with_sphere
    (nerf_util.py:199-205) or intersect_aabb (projects/NeuralLumen/utils/utils.py:86-123)."""
    if cfg.bounding == "box":
        aabb = torch.tensor(cfg.aabb, dtype=torch.float32)
        t0 = (aabb[:3] - center) / ray_unit
        t1 = (aabb[3:] - center) / ray_unit
        near = torch.minimum(t0, t1).amax(dim=-1, keepdim=True).clamp(min=0, max=1e10)
        far = torch.maximum(t0, t1).amin(dim=-1, keepdim=True).clamp(min=0, max=1e10)
        outside = far <= near
    else:
        ctc = (center * center).sum(dim=-1, keepdim=True)
        ctv = (center * ray_unit).sum(dim=-1, keepdim=True)
        disc = ctv ** 2 - (ctc - 1.0)
        near = (-ctv - disc.sqrt()).relu()
        far = -ctv + disc.sqrt()
        outside = near.isnan()
    near = torch.where(outside, torch.ones_like(near), near)
    far = torch.where(outside, torch.full_like(far, 1.2), far)
    return near, far, outside


def compositing_weights(alphas):
    """render.alpha_compositing_weights (projects/nerf/utils/render.py:87-99): w_i = a_i * prod_{j<i}(1-a_j)."""
    front = torch.cat([torch.zeros_like(alphas[..., :1]), alphas[..., :-1]], dim=2)
    return (alphas * (1 - front).cumprod(dim=2))[..., None]


def composite(q, w):
    """render.composite (render.py:102-112)."""
    return (q * w).sum(dim=2)


def sample_from_pdf(bins, weights, n_fine):
    """nerf_util.sample_dists_from_pdf (projects/nerf/utils/nerf_util.py:41-68).  Also returns idx/low/high."""
    pdf = F.normalize(weights, p=1, dim=-1)
    cdf = pdf.cumsum(dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    grid = torch.linspace(0, 1, n_fine + 1)
    unif = (0.5 * (grid[:-1] + grid[1:])).repeat(*cdf.shape[:-1], 1)
    idx = torch.searchsorted(cdf, unif, right=True)
    low = (idx - 1).clamp(min=0)
    high = idx.clamp(max=cdf.shape[-1] - 1)
    b = bins[..., 0]
    d_lo, d_hi = b.gather(2, low), b.gather(2, high)
    c_lo, c_hi = cdf.gather(2, low), cdf.gather(2, high)
    t = (unif - c_lo) / (c_hi - c_lo + 1e-8)
    return (d_lo + t * (d_hi - d_lo))[..., None], dict(idx=idx, low=low, high=high, cdf=cdf)


def hierarchical_weights(dists, sdfs, inv_s):
    """First half of Model.sample_dists_hierarchical (projects/neuralangelo/model.py:467-482)."""
    s = sdfs[..., 0]
    d = dists[..., 0]
    ps, ns = s[..., :-1], s[..., 1:]
    pd, nd = d[..., :-1], d[..., 1:]
    mid = (ps + ns) * 0.5
    cos = (ns - ps) / (nd - pd + 1e-5)
    prev_cos = torch.cat([torch.zeros_like(cos[..., :1]), cos[..., :-1]], dim=-1)
    cos = torch.minimum(prev_cos, cos)
    intv = nd - pd
    p_cdf = ((mid - cos * intv * 0.5) * inv_s).sigmoid()
    n_cdf = ((mid + cos * intv * 0.5) * inv_s).sigmoid()
    alphas = ((p_cdf - n_cdf) / (p_cdf + 1e-5)).clip(0.0, 1.0)
    return compositing_weights(alphas)[..., 0]


@torch.no_grad()
def sample_dists_all(p, cfg, center, ray_unit, near, far, rands=None, trace=None):
    """Model.sample_dists_all + nerf_util.sample_dists (projects/neuralangelo/model.py:449-465;
    nerf_util.py:20-38).  ``rands`` [B,R,coarse,1] replaces torch.rand (None -> 0.5 = not stratified)."""
    B, R = ray_unit.shape[:2]
    if rands is None:
        rands = torch.full((B, R, cfg.coarse, 1), 0.5)
    r = rands + torch.arange(cfg.coarse, dtype=torch.float32)[None, None, :, None]
    dists = r / cfg.coarse * (far[..., None] - near[..., None]) + near[..., None]
    pts = center[..., None, :] + ray_unit[..., None, :] * dists
    sdfs = sdf_only(p, cfg, pts)
    for h in range(cfg.hierarchy):
        w = hierarchical_weights(dists, sdfs, inv_s=64 * 2 ** h)
        fine, info = sample_from_pdf(dists, w, cfg.fine)
        if trace is not None:
            trace.append(dict(dists_in=dists.clone(), sdfs_in=sdfs.clone(), weights=w, fine=fine, **info))
        dists, order = torch.cat([dists, fine], dim=2).sort(dim=2)
        if h != cfg.hierarchy - 1:
            pts_f = center[..., None, :] + ray_unit[..., None, :] * fine
            sdfs = torch.cat([sdfs, sdf_only(p, cfg, pts_f)], dim=2).gather(2, order)
    return dists


# ----------------------------------------------------------------------------------------------------------
# NeuS alpha, compositing, the render itself
# ----------------------------------------------------------------------------------------------------------
def neus_alphas(p, cfg, ray_unit, sdfs, gradients, dists, far, progress):
    """Model.compute_neus_alphas + _get_iter_cos (projects/neuralangelo/model.py:492-515)."""
    s = sdfs[..., 0]
    inv_s = p["s_var"].exp()
    true_cos = (ray_unit[..., None, :] * gradients).sum(dim=-1)
    a = min(progress / cfg.anneal_end, 1.0)
    iter_cos = -((-true_cos * 0.5 + 0.5).relu() * (1.0 - a) + (-true_cos).relu() * a)
    d = torch.cat([dists, far[..., None]], dim=2)[..., 0]
    intv = d[..., 1:] - d[..., :-1]
    p_cdf = ((s - iter_cos * intv * 0.5) * inv_s).sigmoid()
    n_cdf = ((s + iter_cos * intv * 0.5) * inv_s).sigmoid()
    return ((p_cdf - n_cdf) / (p_cdf + 1e-5)).clip(0.0, 1.0)


def render_rays(p, cfg, center, ray_unit, pts_light, rands=None, training=True, progress=1.0, keep=False):
    """Model.render_rays_lumen + render_rays_object_lumen (projects/NeuralLumen/model.py:232-336, 338-403),
    background disabled, light visibility disabled."""
    with torch.no_grad():
        near, far, outside = dist_bounds(cfg, center, ray_unit)
        dists = sample_dists_all(p, cfg, center, ray_unit, near, far, rands)
    pts = center[..., None, :] + ray_unit[..., None, :] * dists
    sdfs, feats = sdf_network(p, cfg, pts)
    sdfs = torch.where(outside[..., None].expand_as(sdfs), torch.full_like(sdfs, cfg.outside_val), sdfs)  # :343
    gradients, hessians = sdf_gradients(p, cfg, pts, sdfs, training)
    normals = F.normalize(gradients, dim=-1)
    rays_unit = ray_unit[..., None, :].expand_as(pts)
    light = pts_light[..., None, :].expand_as(pts)
    heads = lumen_heads(p, cfg, pts, normals, rays_unit, feats, light)
    alphas = neus_alphas(p, cfg, ray_unit, sdfs, gradients, dists, far, progress)
    weights = compositing_weights(alphas)
    opacity_all = composite(1.0, weights)
    white = (1 - opacity_all) if cfg.white_background else 0.0
    out = dict(outside=outside, dists=dists, weights=weights, gradients=gradients, hessians=hessians,
               opacity=None, gradient=None)
    m = cfg.network_mode
    if m == "rgb_r_s":  # model.py:294-305
        out["rgb"] = composite(heads["rgbs"], weights) + white
        out["o_r"] = composite(heads["o_r"], weights) + white
        out["o_s"] = composite(heads["o_s"], weights) + white
        out["o_re"] = out["rgb"] - out["o_r"] * out["o_s"]
    elif m == "rgb_r":  # :284-293
        out["rgb"] = composite(heads["rgbs"], weights) + white
        out["o_r"] = composite(heads["o_r"], weights) + white
        out["o_s"] = out["rgb"] / out["o_r"]
    elif m in ("r_s", "r_s_re"):  # :269-283
        acc = {k: composite(v, weights) + white for k, v in heads.items()}
        out.update(acc)
        out["rgb"] = acc["o_r"] * acc["o_s"] + (acc["o_re"] if m == "r_s_re" else 0.0)
    else:  # :306-310
        out["rgb"] = composite(heads["rgbs"], weights) + white
    if not training:  # :365-368
        out["opacity"] = opacity_all
        out["gradient"] = composite(gradients, weights)
    if keep:
        out.update(sdfs=sdfs, feats=feats, alphas=alphas, near=near, far=far, **{"s_" + k: v for k, v in heads.items()})
    return out


# ----------------------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------------------
def sphere_tracing_intersection(p, cfg, center, ray_unit, near, far, num_iters=20, dist_start=None):
    """Model.sphere_tracing_intersection (projects/neuralangelo/model.py:298-325)."""
    dist = dist_start.clone() if dist_start is not None else near.clone()
    mask = torch.ones_like(dist, dtype=torch.bool)
    for _ in range(num_iters):
        pts = center + ray_unit * dist
        sdfs = sdf_only(p, cfg, pts)
        dist[mask] += sdfs[mask]
        mask[dist > far] = False
        mask[dist < near] = False
    dist = torch.clamp(dist, near, far)
    return dist, center + ray_unit * dist, mask


def light_visibility(p, cfg, center, ray_unit, pts_light, near, far, blend_dist, gradient,
                     camera_ray_type="blend_z_sphere_tracing", radius=0.95, aabb=None):
    """Model.get_light_visibility with type 'sphere_tracing' (projects/NeuralLumen/model.py:133-200) and a sphere
    visibility bound, or -- `aabb` given -- the box bound, which the reference takes from the DATA box
    `self.bounding_box_aabb` (model.py:188-191).  blend_dist = composite(dists, weights), gradient = composited gradient.
    -> visibility (bool), normal_x_light, inter_dist, inter_mask   (all [B,R,1])"""
    if camera_ray_type == "blend_z_sphere_tracing":
        inter_dist, inter_pts, inter_mask = sphere_tracing_intersection(p, cfg, center, ray_unit, near, far,
                                                                        dist_start=blend_dist)
    elif camera_ray_type == "blend_z":
        inter_dist = blend_dist
        inter_pts = center + ray_unit * inter_dist
        inter_mask = inter_dist > 0.0
    elif camera_ray_type == "sphere_tracing":
        inter_dist, inter_pts, inter_mask = sphere_tracing_intersection(p, cfg, center, ray_unit, near, far)
    else:
        raise NotImplementedError
    light_ray = inter_pts - pts_light
    light_ray_unit = F.normalize(light_ray, dim=-1)
    if aabb is not None:  # get_dist_bounds_visibility, box branch (model.py:189-191; intersect_aabb, utils.py:86-123)
        box = torch.tensor(aabb, dtype=torch.float32)
        t0 = (box[:3] - pts_light) / light_ray_unit
        t1 = (box[3:] - pts_light) / light_ray_unit
        near_l = torch.minimum(t0, t1).amax(dim=-1, keepdim=True).clamp(min=0, max=1e10)
        far_l = torch.maximum(t0, t1).amin(dim=-1, keepdim=True).clamp(min=0, max=1e10)
        outside_l = far_l <= near_l
    else:  # sphere branch (model.py:192-197)
        ctc = (pts_light * pts_light).sum(dim=-1, keepdim=True)
        ctv = (pts_light * light_ray_unit).sum(dim=-1, keepdim=True)
        disc = ctv ** 2 - (ctc - radius ** 2)
        near_l = (-ctv - disc.sqrt()).relu()
        far_l = -ctv + disc.sqrt()
        outside_l = near_l.isnan()
    near_l = torch.where(outside_l, torch.ones_like(near_l), near_l)
    far_l = torch.where(outside_l, torch.full_like(far_l, 1.2), far_l)
    far_tracing = light_ray.norm(dim=-1, keepdim=True) - 1e-3
    inside_bounding = (near_l < far_tracing) & (far_tracing < far_l) & ~outside_l
    _, _, mask_light = sphere_tracing_intersection(p, cfg, pts_light, light_ray_unit, near_l, far_tracing)
    visibility = ~mask_light | ~inside_bounding
    normal_ray_unit = F.normalize(-gradient, dim=-1)
    normal_x_light = (normal_ray_unit * light_ray_unit).sum(dim=-1, keepdim=True).relu()
    return visibility, normal_x_light, inter_dist, inter_mask


def eikonal_loss(gradients, outside):
    """projects/neuralangelo/utils/misc.py:74-80."""
    err = ((gradients.norm(dim=-1) - 1.0) ** 2).nan_to_num(nan=0.0, posinf=0.0, neginf=0.0)
    return (err * (~outside).float()).mean()


def curvature_loss(hessians, outside):
    """projects/neuralangelo/utils/misc.py:83-89."""
    lap = hessians.sum(dim=-1).abs().nan_to_num(nan=0.0, posinf=0.0, neginf=0.0)
    return (lap * (~outside).float()).mean()


def _minmax(x, lo, hi):
    return lo + (x - x.min()) / torch.clamp(x.max() - x.min(), min=1e-6) * (hi - lo)


def intrinsic_loss(cfg, o_r, o_s, ref, sha, vis):
    """projects/NeuralLumen/utils/utils.py:142-162."""
    w_sha = _minmax(sha.detach(), *cfg.range_shading)
    w_vis = _minmax(vis.detach(), *cfg.range_visibility)
    w_ref = torch.minimum(w_vis, w_sha)
    return (torch.mean(torch.abs(o_r - ref) * w_ref) * cfg.factor_ref
            + torch.mean(torch.abs(o_s - sha) * w_sha) * cfg.factor_sha)


def regularize_re_loss(cfg, o_re):
    """projects/NeuralLumen/utils/utils.py:165-174."""
    zero = torch.zeros((), dtype=o_re.dtype)
    neg = torch.where(o_re < 0.0, o_re, zero).abs().mean()
    pos = torch.where(o_re >= 0.0, o_re, zero).pow(cfg.exponent_positive).mean()
    return neg * cfg.factor_negative + pos * cfg.factor_positive


def total_loss(cfg, out, targets):
    """Trainer._compute_loss(mode='train') + _get_total_loss (projects/NeuralLumen/trainer.py:133-149;
    imaginaire/trainers/base.py:534-544)."""
    losses = dict(render=F.l1_loss(out["rgb"], targets["image_sampled"]) * 3,
                  eikonal=eikonal_loss(out["gradients"], out["outside"]),
                  curvature=curvature_loss(out["hessians"], out["outside"]))
    weights = dict(render=cfg.w_render, eikonal=cfg.w_eikonal, curvature=cfg.w_curvature)
    if "o_r" in out and "pseudo_ref_sampled" in targets:
        losses["intrinsic"] = intrinsic_loss(cfg, out["o_r"], out["o_s"], targets["pseudo_ref_sampled"],
                                             targets["pseudo_sha_sampled"],
                                             targets["pseudo_visibility_certainty_sampled"])
        weights["intrinsic"] = cfg.w_intrinsic
    if "o_re" in out:
        losses["regularize_re"] = regularize_re_loss(cfg, out["o_re"])
        weights["regularize_re"] = cfg.w_regularize_re
    total = sum(weights[k] * v for k, v in losses.items())
    psnr = -10 * F.mse_loss(out["rgb"], targets["image_sampled"]).log10()
    return total, losses, psnr


# ----------------------------------------------------------------------------------------------------------
# parameter construction (same shapes / init distributions / key names as the reference's state_dict)
# ----------------------------------------------------------------------------------------------------------
def head_in_dims(cfg):
    sh = (cfg.sh_levels + 1) ** 2
    full, geo, geo_l = 3 + sh + 3 + cfg.hidden + sh, 3 + 3 + cfg.hidden, 3 + 3 + cfg.hidden + sh
    return {"rgb_r_s": {"mlp": (full, 3), "mlp_r": (geo, 3), "mlp_s": (geo_l, 1)},
            "rgb_r": {"mlp": (full, 3), "mlp_r": (geo, 3)},
            "r_s": {"mlp_r": (geo, 3), "mlp_s": (full, 3)},
            "r_s_re": {"mlp_r": (geo, 3), "mlp_s": (geo_l, 3), "mlp_re": (full, 3)},
            "rgb": {"mlp": (full, 3)}, None: {"mlp": (full, 3)}}[cfg.network_mode]


def init_params(cfg, seed=0, table_scale=1e-4, generic=False):
    """State-dict with the reference's key names/shapes.  ``generic=False`` follows the
    reference's init (geometric SDF init, mlp.py:71-84; default Linear init for heads).  ``generic=True`` draws
    every tensor from a distribution that exercises all inputs (encoding columns of W0 non-zero, larger table),
    which is what parity tests want."""
    g = torch.Generator().manual_seed(seed)
    H, enc = cfg.hidden, cfg.n_levels * cfg.feat_per_level
    p = {"s_var": torch.tensor(3.0)}
    p["neural_sdf.tcnn_encoding.params"] = (torch.rand(cfg.n_table_params(), generator=g) * 2 - 1) * table_scale

    def wn_layer(prefix, k_in, k_out, w):
        p[prefix + ".weight_v"] = w
        p[prefix + ".weight_g"] = w.norm(dim=1, keepdim=True).clone()
        p[prefix + ".bias"] = torch.zeros(k_out)

    w0 = torch.randn(H, 3 + enc, generator=g) * math.sqrt(2 / H)
    if not generic:
        w0[:, 3:] = 0.0
    else:
        w0[:, 3:] *= 4.0
    wn_layer("neural_sdf.mlp.linears.0", 3 + enc, H, w0)
    wn_layer("neural_sdf.mlp.linears.1", H, H, torch.randn(H, H, generator=g) * math.sqrt(2 / H))
    p["neural_sdf.mlp.linear_sdf.weight"] = math.sqrt(math.pi / H) + 1e-4 * torch.randn(1, H, generator=g)
    p["neural_sdf.mlp.linear_sdf.bias"] = torch.tensor([-0.5])
    for name, (k_in, k_out) in head_in_dims(cfg).items():
        dims = [k_in] + [H] * 4 + [k_out]
        for li in range(5):
            bound = 1 / math.sqrt(dims[li])
            w = (torch.rand(dims[li + 1], dims[li], generator=g) * 2 - 1) * bound
            wn_layer(f"neural_rgb.{name}.linears.{li}", dims[li], dims[li + 1], w)
            if li != 4:
                p[f"neural_rgb.{name}.linears.{li}.bias"] = (torch.rand(dims[li + 1], generator=g) * 2 - 1) * bound
    if generic:
        for k in list(p):
            if k.endswith("weight_g"):
                p[k] = p[k] * (0.75 + 0.5 * torch.rand(p[k].shape, generator=g))
            if k.endswith(".bias") and "linears.4" not in k and p[k].numel() > 1:
                p[k] = p[k] + 0.01 * torch.randn(p[k].shape, generator=g)
    return p


def synthetic_rays(n_rays, seed=1, batch=1):
    """origins on a radius-3 sphere looking at a 0.3-sigma blob; light at radius 4."""
    g = torch.Generator().manual_seed(seed)
    origin = 3 * F.normalize(torch.randn(batch, n_rays, 3, generator=g), dim=-1)
    target = 0.3 * torch.randn(batch, n_rays, 3, generator=g)
    ray_unit = F.normalize(target - origin, dim=-1)
    light = (4 * F.normalize(torch.randn(3, generator=g), dim=0)).expand(batch, n_rays, 3).contiguous()
    return origin, ray_unit, light


def synthetic_targets(n_rays, seed=2, batch=1):
    g = torch.Generator().manual_seed(seed)
    return dict(image_sampled=torch.rand(batch, n_rays, 3, generator=g),
                pseudo_ref_sampled=torch.rand(batch, n_rays, 3, generator=g),
                pseudo_sha_sampled=torch.rand(batch, n_rays, 1, generator=g),
                pseudo_visibility_certainty_sampled=torch.rand(batch, n_rays, 1, generator=g))
