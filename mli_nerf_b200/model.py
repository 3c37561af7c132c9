"""Drop-in replacement for ``projects.NeuralLumen.model`` (select it with ``--model.type=mli_nerf_b200.model``).

Same constructor, ``forward(data)`` / ``inference(data)`` output dictionaries, trainer-facing attributes and
``state_dict`` keys as the reference model (/root/reference/projects/NeuralLumen/model.py:17-131,422-438;
projects/neuralangelo/model.py:29-60) -- but every FLOP of the render path runs in
libmli_b200.so (hand-written sm_100a CUDA behind a C ABI).  PyTorch provides parameters, device memory, the stream
and autograd glue only.  There is no CPU fallback: constructing works anywhere (checkpoints can be inspected on a
CPU box), calling forward/inference without a B200-class GPU raises.
"""
import math
import os
from functools import partial

import weakref

import numpy as np
import torch

from . import _lib
from .engine import PathCfg, RenderEngine, head_layout, _col_map
from .hashgrid import Encoding


# ----------------------------------------------------------------------------------------------------------------
# parameter containers: only hold tensors under the reference's names (no forward of their own)
# ----------------------------------------------------------------------------------------------------------------
class _WNLinear(torch.nn.Module):
    """Old-style torch.nn.utils.weight_norm(Linear): parameters ``bias``, ``weight_g`` [out,1], ``weight_v`` [out,in]."""

    def __init__(self, k_in, k_out, weight=None, bias=None):
        super().__init__()
        lin = torch.nn.Linear(k_in, k_out)  # default init (kaiming_uniform a=sqrt(5)), as the reference's heads
        w = lin.weight.data if weight is None else weight
        b = lin.bias.data if bias is None else bias
        self.bias = torch.nn.Parameter(b.clone())
        self.weight_g = torch.nn.Parameter(w.norm(dim=1, keepdim=True).clone())
        self.weight_v = torch.nn.Parameter(w.clone())


class _Params(torch.nn.Module):
    pass


def _head_mlp(k_in, k_out, hidden, n_hidden):
    m = _Params()
    dims = [k_in] + [hidden] * n_hidden + [k_out]
    m.linears = torch.nn.ModuleList([_WNLinear(a, b) for a, b in zip(dims[:-1], dims[1:])])
    m.linears[-1].bias.data.fill_(0.0)  # nerf_util.py:182-183
    return m


class NeuralSDF(torch.nn.Module):
    """Parameter holder + schedule state of projects/neuralangelo/utils/modules.py:24-113 (NeuralSDF)."""

    def __init__(self, cfg_sdf):
        super().__init__()
        self.cfg_sdf = cfg_sdf
        enc, hg = cfg_sdf.encoding, cfg_sdf.encoding.hashgrid
        if enc.type != "hashgrid":
            raise NotImplementedError("Unknown encoding type")  # fourier SDF encoding: not on the shipped path
        r_min, r_max = 2 ** hg.min_logres, 2 ** hg.max_logres
        self.growth_rate = np.exp((np.log(r_max) - np.log(r_min)) / (enc.levels - 1))
        self.tcnn_encoding = Encoding(3, dict(otype="HashGrid", n_levels=enc.levels, n_features_per_level=hg.dim,
                                              log2_hashmap_size=hg.dict_size, base_resolution=r_min,
                                              per_level_scale=self.growth_rate))
        self.resolutions = [np.floor(r_min * self.growth_rate ** lv).astype(int) + 1 for lv in range(enc.levels)]
        cfg_mlp = cfg_sdf.mlp
        if cfg_mlp.num_layers != 1 or list(cfg_mlp.skip) or cfg_mlp.activ != "softplus" or not cfg_mlp.weight_norm:
            raise NotImplementedError("SDF MLP variants other than the shipped 1-layer softplus/weight_norm one")
        k_in, H = 3 + hg.dim * enc.levels, cfg_mlp.hidden_dim
        # geometric init (mlp.py:71-84)
        w0 = torch.randn(H, k_in) * math.sqrt(2 / H)
        w0[:, 3:] = 0.0
        w1 = torch.randn(H, H) * math.sqrt(2 / H)
        self.mlp = _Params()
        self.mlp.linears = torch.nn.ModuleList([_WNLinear(k_in, H, w0, torch.zeros(H)), _WNLinear(H, H, w1, torch.zeros(H))])
        self.mlp.linear_sdf = torch.nn.Linear(H, 1)
        torch.nn.init.normal_(self.mlp.linear_sdf.weight, mean=math.sqrt(math.pi / H), std=0.0001)
        torch.nn.init.constant_(self.mlp.linear_sdf.bias, -cfg_mlp.out_bias)
        if cfg_mlp.inside_out:
            self.mlp.linear_sdf.weight.data *= -1
            self.mlp.linear_sdf.bias.data *= -1
        self.active_levels = enc.levels
        self.anneal_levels = enc.levels
        self.warm_up_end = 0
        self.normal_eps = 1.0 / self.resolutions[-1]

    def __getstate__(self):  # the back-reference to the owning Model is re-created by Model.__setstate__
        state = dict(self.__dict__)
        state.pop("_owner", None)
        return state

    def sdf(self, points_3D):
        """`neural_sdf.sdf(x)` (modules.py:73-74) -- what scripts/extract_mesh.py:101 calls; the query runs in the owning
        Model's engine (encode + SDF-trunk kernels), see Model.sdf."""
        owner = self.__dict__.get("_owner")
        model = owner() if owner is not None else None
        if model is None:
            raise _lib.MliError("NeuralSDF.sdf: this module is not attached to a mli_nerf_b200 Model")
        return model.sdf(points_3D)

    def set_active_levels(self, current_iter=None):  # modules.py:97-100
        c2f = self.cfg_sdf.encoding.coarse2fine
        anneal_levels = max((current_iter - self.warm_up_end) // c2f.step, 1)
        self.anneal_levels = min(self.cfg_sdf.encoding.levels, anneal_levels)
        self.active_levels = max(c2f.init_active_level, self.anneal_levels)

    def set_normal_epsilon(self):  # modules.py:102-107
        if self.cfg_sdf.encoding.coarse2fine.enabled:
            epsilon_res = self.resolutions[self.anneal_levels - 1]
        else:
            epsilon_res = self.resolutions[-1]
        self.normal_eps = 1. / epsilon_res


class LumenRGB(torch.nn.Module):
    """Parameter holder of projects/NeuralLumen/utils/modules.py:9-104 (LumenRGB)."""

    def __init__(self, cfg_rgb, feat_dim, appear_embed):
        super().__init__()
        if appear_embed.enabled:
            raise NotImplementedError("appearance embedding (disabled in every NeuralLumen config)")
        if cfg_rgb.encoding_view.type != "spherical" or cfg_rgb.encoding_view.levels != 3:
            raise NotImplementedError("Unknown encoding type")
        self.network_mode = getattr(cfg_rgb, "network_mode", None) or "rgb"
        if self.network_mode == "rgb" and cfg_rgb.mode != "idr":
            raise NotImplementedError("rgb mode variants no_view_dir / no_normal")
        shading_dim = getattr(cfg_rgb, "shading_dim", 3)
        for name, kind, odim, _ in head_layout(self.network_mode if self.network_mode != "rgb" else None):
            if self.network_mode == "rgb_r_s" and name == "mlp_s":
                odim = shading_dim
                if odim != 1:
                    raise NotImplementedError("rgb_r_s with shading_dim != 1")
            setattr(self, name, _head_mlp(len(_col_map(kind)), odim, cfg_rgb.mlp.hidden_dim, cfg_rgb.mlp.num_layers))


# ----------------------------------------------------------------------------------------------------------------
def _aliases(obj):
    """Same storages, new tensor objects (no autograd history attaches to them later), containers rebuilt."""
    if isinstance(obj, torch.Tensor):
        return obj.detach()
    if isinstance(obj, dict):
        return {k: _aliases(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_aliases(v) for v in obj)
    return obj


class _RenderFunction(torch.autograd.Function):
    """One autograd node for the whole render: forward and hand-written backward both run in libmli_b200."""

    @staticmethod
    def forward(ctx, model, center, ray_unit, pts_light, rands, training, names, *params):
        eng = model.engine
        p = dict(zip(names, params))
        with torch.no_grad():
            eng.pack_weights(p)
            need_bwd = any(ctx.needs_input_grad[7:])  # no parameter needs a gradient (eval / no_grad): skip the saves
            if training and ctx.needs_input_grad[7 + names.index("neural_sdf.tcnn_encoding.params")]:
                eng.start_table_grad_zero()  # the backward's 1.46 GB gradient buffer: zeroed on a side stream meanwhile
            near, far, outside = eng.bounds(center, ray_unit)
            dists = eng.sample(p["neural_sdf.tcnn_encoding.params"], center, ray_unit, near, far, rands)
            res, saved = eng.forward(p, center, ray_unit, pts_light, dists, near, far, outside, training, model.progress,
                                     keep_dz=need_bwd)
        # Several saved tensors (gradients, weights, dists, outside ...) are also OUTPUTS of this node.  An output gets this
        # node as its grad_fn, so keeping the same tensor object on ctx would close a reference cycle output -> node -> ctx
        # -> output, and a step's activations (~3.7 GB at the bench shape) would live until Python's cycle collector
        # runs: the caching allocator then grows by cudaMalloc every step.  Aliases break the cycle.
        ctx.model, ctx.names, ctx.saved, ctx.params = model, names, _aliases(saved), params
        ctx.W = eng.W
        out = res["out"]
        hess = res["hessians"] if res["hessians"] is not None else out.new_zeros(0)
        extras = res["extras"] if res["extras"] is not None else out.new_zeros(0)
        ctx.mark_non_differentiable(dists, outside, extras)
        return out, res["weights"], res["gradients"], hess, dists, outside, extras

    @staticmethod
    def backward(ctx, d_out, d_weights, d_gradients, d_hessians, *_):
        model, eng, names = ctx.model, ctx.model.engine, ctx.names
        p = dict(zip(names, ctx.params))
        need_flags = ctx.needs_input_grad[7:]
        need = set()
        for n, f in zip(names, need_flags):
            if not f:
                continue
            if n == "s_var":
                need.add("s_var")
            elif n == "neural_sdf.tcnn_encoding.params":
                need.add("table")
            elif n.startswith("neural_sdf."):
                need.add("sdf")
            else:
                need.add("heads")
        eng.W = ctx.W
        if d_hessians is not None and d_hessians.numel() == 0:
            d_hessians = None
        with torch.no_grad():
            grads = eng.backward(p, ctx.saved, d_out.contiguous() if d_out is not None else eng._z(ctx.saved["R"], eng.n_out),
                                 d_gradients, d_hessians.contiguous() if d_hessians is not None else None,
                                 d_weights.contiguous() if d_weights is not None else None, need=tuple(need))
        ret = [grads.get(n) if f else None for n, f in zip(names, need_flags)]
        return (None,) * 7 + tuple(ret)


class Model(torch.nn.Module):

    def __init__(self, cfg_model, cfg_data):
        super().__init__()
        self.cfg_render = cfg_model.render
        self.white_background = cfg_model.background.white
        self.with_background = cfg_model.background.enabled
        self.with_appear_embed = cfg_model.appear_embed.enabled
        self.anneal_end = cfg_model.object.s_var.anneal_end
        self.outside_val = 1000. * (-1 if cfg_model.object.sdf.mlp.inside_out else 1)
        self.image_size_train = cfg_data.train.image_size
        self.image_size_val = cfg_data.val.image_size
        if self.with_background:
            raise NotImplementedError  # NeuralLumen/model.py:247-249 raises for every network_mode as well
        if self.with_appear_embed:
            raise NotImplementedError("appearance embedding (disabled in every NeuralLumen config)")
        sdf_cfg = cfg_model.object.sdf
        if sdf_cfg.gradient.mode != "numerical":
            raise NotImplementedError("analytical gradient mode (not used by any shipped config)")
        if sdf_cfg.gradient.taps not in (4, 6):
            raise ValueError("Only support 4 or 6 taps.")
        self.neural_sdf = NeuralSDF(sdf_cfg)
        self.neural_sdf.__dict__["_owner"] = weakref.ref(self)  # plain attribute: no module cycle, not in the state_dict
        self.rgb_network_mode = getattr(cfg_model.object.rgb, "network_mode", None)
        self.neural_rgb = LumenRGB(cfg_model.object.rgb, feat_dim=sdf_cfg.mlp.hidden_dim,
                                   appear_embed=cfg_model.appear_embed)
        self.background_nerf = None
        self.appear_embed = self.appear_embed_outside = None
        self.s_var = torch.nn.Parameter(torch.tensor(cfg_model.object.s_var.init_val, dtype=torch.float32))
        if getattr(cfg_data, "bounding_type", None) == "box":
            self.bounding_type = "box"
            self.bounding_box_aabb = torch.tensor(cfg_data.bounding_box_aabb)
        else:
            self.bounding_type = "unit_sphere"
        self.rand_rays_val = getattr(cfg_model.render, "rand_rays_val", cfg_model.render.rand_rays)
        lv = getattr(cfg_model, "light_visibility", None)
        self.flag_light_visibility = bool(lv is not None and lv.enabled)
        self.flag_gamma_correlation = False
        if self.flag_light_visibility:  # NeuralLumen/model.py:25-35
            self.para_light_visibility = lv
            if lv.type != "sphere_tracing":
                raise NotImplementedError("light_visibility.type 'render_light_visibility' reads an attribute the reference "
                                          "never defines (model.py:217); only 'sphere_tracing' is usable")
            if lv.visibility_bounding_type == "box":
                self.visibility_bounding_box_aabb = torch.tensor(lv.visibility_bounding_box_aabb)
            elif lv.visibility_bounding_type != "sphere":
                raise NotImplementedError
            if hasattr(lv, "gamma_correlation"):
                self.flag_gamma_correlation = True
                self.gamma_for_shading = lv.gamma_correlation
        hg = sdf_cfg.encoding.hashgrid
        ns = cfg_model.render.num_samples
        self.path_cfg = PathCfg(
            n_levels=sdf_cfg.encoding.levels, feat_per_level=hg.dim, log2_hashmap_size=hg.dict_size,
            min_logres=hg.min_logres, max_logres=hg.max_logres, vol_range=(float(hg.range[0]), float(hg.range[1])),
            hidden=sdf_cfg.mlp.hidden_dim, taps=sdf_cfg.gradient.taps, coarse=ns.coarse, fine=ns.fine,
            hierarchy=cfg_model.render.num_sample_hierarchy, sh_levels=cfg_model.object.rgb.encoding_view.levels,
            network_mode=self.rgb_network_mode, white_background=bool(self.white_background),
            anneal_end=self.anneal_end, outside_val=self.outside_val,
            bounding="box" if self.bounding_type == "box" else "unit_sphere",
            aabb=tuple(float(v) for v in cfg_data.bounding_box_aabb) if self.bounding_type == "box" else None,
            c2f_enabled=bool(sdf_cfg.encoding.coarse2fine.enabled),
            precision={"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[
                getattr(cfg_model, "mli_precision", os.environ.get("MLI_PRECISION", "fp32"))])
        self.progress = 1.0
        self._engine = None
        self.max_rays_per_launch = 8192

    # -- trainer-facing helpers (imaginaire/models/base.py:16-30; NeuralLumen/model.py:422-438) ---------------------
    def device(self):
        return next(self.parameters()).device

    def __getstate__(self):  # engines hold CUDA streams / ctypes structs: rebuilt lazily after unpickling / deepcopy
        state = dict(self.__dict__)
        state["_engine"] = None
        state.pop("_graph_state", None)   # captured CUDA graphs / their static buffers do not travel
        state.pop("_graph_last_key", None)
        state.pop("_last_render", None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self.neural_sdf.__dict__["_owner"] = weakref.ref(self)

    def get_param_groups(self, cfg_optim):
        if hasattr(cfg_optim, "partial_training"):
            keyword_list = cfg_optim.partial_training
            return [param for name, param in self.named_parameters() if any(k in name for k in keyword_list)]
        return self.parameters()

    @property
    def engine(self):
        if self._engine is None:
            dev = self.device()
            if dev.type != "cuda" or not _lib.device_ok():
                raise _lib.MliError("mli_nerf_b200.Model needs a B200-class CUDA device: the render path has no CPU "
                                    "fallback (parameters live on %s)" % dev)
            self._engine = RenderEngine(self.path_cfg, dev)
        eng = self._engine
        eng.normal_eps = float(self.neural_sdf.normal_eps)
        eng.set_active_levels(self.neural_sdf.active_levels)
        return eng

    def _named(self):
        names, params = [], []
        for n, p in self.named_parameters():
            names.append(n)
            params.append(p)
        return tuple(names), params

    # -- rays ---------------------------------------------------------------------------------------------------------
    def _rays(self, pose, intr, pose_light, image_size, ray_idx):
        B = pose.shape[0]
        R = ray_idx.shape[1] if ray_idx is not None else image_size[0] * image_size[1]
        f = partial(torch.empty, dtype=torch.float32, device=pose.device)
        center, ray_unit, light, norm = f(B * R, 3), f(B * R, 3), f(B * R, 3), f(B * R)
        if ray_idx is not None:  # the kernel reads int64 indices
            ray_idx = ray_idx.to(device=pose.device, dtype=torch.int64).contiguous()
        _lib.call("mli_rays_from_pose", pose.contiguous().float(), intr.contiguous().float(),
                  pose_light.contiguous().float(), ray_idx, B, R,
                  int(image_size[1]), center, ray_unit, norm, light)
        return center, ray_unit, light, norm

    def _empty_render(self, B, R, device):
        """Output dict of an EMPTY ray batch (B*R == 0): the keys and trailing dimensions of the regular path, no launch
        (the reference's torch code returns empty tensors as well)."""
        N = self.path_cfg.n_samples
        z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=device)  # noqa: E731
        res = dict(rgb=z(B, R, 3), opacity=None, outside=torch.zeros(B, R, 1, dtype=torch.bool, device=device),
                   dists=z(B, R, N, 1), weights=z(B, R, N, 1), gradient=None, gradients=z(B, R, N, 3),
                   hessians=z(B, R, N, 3) if self.training else None)
        m = self.rgb_network_mode
        if m == "rgb_r_s":
            res.update(o_r=z(B, R, 3), o_s=z(B, R, 1), o_re=z(B, R, 3))
        elif m in ("rgb_r", "r_s"):
            res.update(o_r=z(B, R, 3), o_s=z(B, R, 3))
        elif m == "r_s_re":
            res.update(o_r=z(B, R, 3), o_s=z(B, R, 3), o_re=z(B, R, 3))
        if not self.training:
            res.update(opacity=z(B, R, 1), gradient=z(B, R, 3), _dist=z(B, R, 1))
            if self.flag_light_visibility:
                res.update(visibility=torch.zeros(B, R, 1, dtype=torch.bool, device=device), normal_x_light=z(B, R, 1),
                           pseudo_shading=z(B, R, 1), inter_dist=z(B, R, 1),
                           inter_mask=torch.zeros(B, R, 1, dtype=torch.bool, device=device))
        return res

    def render_rays_lumen(self, center, ray_unit, pts_light, sample_idx=None, stratified=False, rands=None):
        """[B,R,3] rays -> the reference's output dict (NeuralLumen/model.py:232-336)."""
        B, R = center.shape[:2]
        if B * R == 0:
            return self._empty_render(B, R, center.device)
        c, r, l = (t.reshape(B * R, 3).contiguous().float() for t in (center, ray_unit, pts_light))
        if stratified and rands is None:
            rands = torch.rand(B, R, self.path_cfg.coarse, 1, device=c.device)  # nerf_util.py:33, same shape/order
        if rands is not None:
            rands = rands.reshape(B * R, self.path_cfg.coarse).contiguous()
        names, params = self._named()
        out, weights, gradients, hess, dists, outside, extras = _RenderFunction.apply(
            self, c, r, l, rands, self.training, names, *params)
        N = self.path_cfg.n_samples
        res = dict(rgb=out[:, 0:3].view(B, R, 3), opacity=None, outside=outside.view(B, R, 1).bool(),
                   dists=dists.view(B, R, N, 1), weights=weights.view(B, R, N, 1), gradient=None,
                   gradients=gradients.view(B, R, N, 3), hessians=hess.view(B, R, N, 3) if self.training else None)
        m = self.rgb_network_mode
        if m == "rgb_r_s":
            res.update(o_r=out[:, 3:6].view(B, R, 3), o_s=out[:, 6:7].view(B, R, 1), o_re=out[:, 7:10].view(B, R, 3))
        elif m in ("rgb_r", "r_s"):
            res.update(o_r=out[:, 3:6].view(B, R, 3), o_s=out[:, 6:9].view(B, R, 3))
        elif m == "r_s_re":
            res.update(o_r=out[:, 3:6].view(B, R, 3), o_s=out[:, 6:9].view(B, R, 3), o_re=out[:, 9:12].view(B, R, 3))
        if not self.training:
            res["opacity"] = extras[:, 0:1].view(B, R, 1)
            res["gradient"] = extras[:, 1:4].view(B, R, 3)
            res["_dist"] = extras[:, 4:5].view(B, R, 1)
        if self.flag_light_visibility:  # NeuralLumen/model.py:325-334 (the stage-a export of pseudo shading labels)
            lv, eng = self.para_light_visibility, self.engine
            with torch.no_grad():
                if self.training:
                    # training mode composites the gradient only when visibility is on (NeuralLumen/model.py:369-372);
                    # the eval extras are not produced there, so the two per-ray blends are formed from the weights
                    grad_c = (weights.view(B * R, N, 1) * gradients.view(B * R, N, 3)).sum(dim=1)
                    blend = (weights.view(B * R, N) * dists.view(B * R, N)).sum(dim=1)
                    res["gradient"] = grad_c.view(B, R, 3)
                else:
                    grad_c, blend = extras[:, 1:4].contiguous(), extras[:, 4].contiguous()
                near, far, _ = eng.bounds(c, r)
                aabb = None
                if lv.visibility_bounding_type == "box":
                    # get_dist_bounds_visibility intersects the light rays with the DATA box (`self.bounding_box_aabb`,
                    # NeuralLumen/model.py:190), not with light_visibility.visibility_bounding_box_aabb
                    if not hasattr(self, "bounding_box_aabb"):
                        raise AttributeError("light_visibility.visibility_bounding_type='box' needs data.bounding_type='box' "
                                             "(the reference reads self.bounding_box_aabb, NeuralLumen/model.py:190)")
                    aabb = [float(v) for v in self.bounding_box_aabb]
                vis, nxl, inter_dist, inter_mask = eng.light_visibility(
                    dict(zip(names, params))["neural_sdf.tcnn_encoding.params"], c, r, l, near, far,
                    blend, grad_c, lv.camera_ray_type, getattr(lv, "visibility_sphere_radius", 1.0), aabb)
            res["visibility"] = vis.view(B, R, 1).bool()
            res["normal_x_light"] = nxl.view(B, R, 1)
            res["pseudo_shading"] = res["normal_x_light"] * res["visibility"].float()
            res["inter_dist"] = inter_dist.view(B, R, 1)
            res["inter_mask"] = inter_mask.view(B, R, 1).bool()
            if self.flag_gamma_correlation:
                res["pseudo_shading"] = torch.pow(res["pseudo_shading"], 1.0 / self.gamma_for_shading)
        return res

    @torch.no_grad()
    def sdf(self, points, chunk=1 << 20):
        """``neural_sdf.sdf(points)`` of the reference (projects/neuralangelo/utils/modules.py:73-74), the query the mesh
        extraction sweeps over its 512^3 lattice (projects/neuralangelo/utils/mesh.py:25-49): points [...,3] -> [...,1].
        Runs the encode + SDF-trunk kernels of the sampling path (no head layers)."""
        eng = self.engine
        names, params = self._named()
        p = dict(zip(names, params))
        eng.pack_weights(p)
        pts = points.reshape(-1, 3).contiguous().float()
        out = torch.empty(pts.shape[0], dtype=torch.float32, device=pts.device)
        for s in range(0, pts.shape[0], chunk):
            c = pts[s:s + chunk]
            zeros3, zeros1 = torch.zeros_like(c), torch.zeros(c.shape[0], 1, device=c.device)
            # a "ray" of one sample at distance 0 along a zero direction is the point itself
            out[s:s + chunk] = eng.sdf_query(p["neural_sdf.tcnn_encoding.params"], c, zeros3, zeros1, 1, 1)
        return out.view(*points.shape[:-1], 1)

    def forward(self, data):
        """NeuralLumen/model.py:113-131."""
        pose = data["pose"]
        B, R = data["ray_idx"].shape
        c, r, l, _ = self._rays(pose, data["intr"], data["pose_light"], self.image_size_train, data["ray_idx"])
        return self.render_rays_lumen(c.view(B, R, 3), r.view(B, R, 3), l.view(B, R, 3), sample_idx=data.get("idx"),
                                      stratified=self.cfg_render.stratified)

    @torch.no_grad()
    def fused_train_step(self, data, loss_cfg, accumulate=False, use_graph=False, after_backward=None):
        """forward + in-kernel losses + hand-written backward in one pass (no autograd graph).

        Equivalent to ``total = trainer.model_forward(data); total.backward()`` of the reference
        (projects/nerf/trainers/base.py:99-107, NeuralLumen/trainer.py:133-149,189-196): fills ``.grad`` of every
        parameter with ``requires_grad`` and returns the device tensor of losses (see _lib.LOSS_NAMES).
        ``use_graph``: replay the ~75 kernel launches of the step from ONE CUDA graph (see _graphed_train_step: the
        per-iteration schedule scalars are device resident, only a coarse-to-fine level change re-captures)."""
        if use_graph and not accumulate:
            return self._graphed_train_step(data, loss_cfg, after_backward)
        eng = self.engine
        B, R = data["ray_idx"].shape
        if B * R == 0:  # the reference's mean-reduced losses are NaN for an empty batch: refuse instead of training on NaN
            raise ValueError("fused_train_step: empty ray batch")
        c, r, l, _ = self._rays(data["pose"], data["intr"], data["pose_light"], self.image_size_train, data["ray_idx"])
        names, params = self._named()
        p = dict(zip(names, params))
        rands = None
        if self.cfg_render.stratified:
            rands = torch.rand(B, R, self.path_cfg.coarse, 1, device=c.device).view(B * R, self.path_cfg.coarse)
        eng.pack_weights(p)
        if p["neural_sdf.tcnn_encoding.params"].requires_grad:
            eng.start_table_grad_zero()  # side stream: hidden behind the sampling rounds
        near, far, outside = eng.bounds(c, r)
        dists = eng.sample(p["neural_sdf.tcnn_encoding.params"], c, r, near, far, rands)
        res, ctx = eng.forward(p, c, r, l, dists, near, far, outside, True, self.progress)
        tg = {k: v.reshape(B * R, v.shape[-1]).to(device=c.device, dtype=torch.float32).contiguous()
              for k, v in data.items() if k.endswith("_sampled")}
        losses, d_out, d_grad, d_hess = eng.losses(loss_cfg, res["out"], res["gradients"], res["hessians"], outside, tg)
        need = set()
        for n, q in zip(names, params):
            if q.requires_grad:
                need.add("s_var" if n == "s_var" else "table" if n == "neural_sdf.tcnn_encoding.params"
                         else "sdf" if n.startswith("neural_sdf.") else "heads")
        grads = eng.backward(p, ctx, d_out, d_grad, d_hess, None, need=tuple(need))
        for n, q in zip(names, params):
            if q.requires_grad and n in grads:
                g = grads[n].view_as(q)
                if accumulate and q.grad is not None:
                    q.grad.add_(g)
                else:
                    q.grad = g
        self._last_render = res
        if after_backward is not None:  # e.g. GradReducer.allreduce_grads: part of the step (and of its CUDA graph)
            after_backward()
        return losses

    def _graphed_train_step(self, data, loss_cfg, after_backward=None):
        """CUDA-graph replay of the step.  The two schedule values that move EVERY iteration for long stretches of
        training -- the s_var anneal ratio (first anneal_end*max_iter iterations) and the loss weights (curvature
        warm-up) -- are read by the kernels from a small device buffer that is refreshed before each replay
        (engine.set_dynamic_scalars), so they are not part of the graph.  What remains baked into the captured launches is
        piecewise constant (tap epsilon and active levels change 16 times over a coarse-to-fine run; the non-weight
        fields of the loss config, requires_grad flags, shapes): the graph is keyed on it, ONE graph is kept (the previous
        one is dropped when the key changes), and a new key is only captured once it has been seen on two consecutive
        steps -- the step in between runs eagerly."""
        tensors = {k: v for k, v in data.items() if isinstance(v, torch.Tensor) and v.is_cuda}
        eng = self.engine
        static_cfg = _lib.LossCfg.from_buffer_copy(loss_cfg)
        static_cfg.w_render = static_cfg.w_eikonal = static_cfg.w_curvature = 0.0
        static_cfg.w_intrinsic = static_cfg.w_regularize_re = 0.0
        static_cfg.weights_dev = None
        key = (tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(tensors.items())),
               float(self.neural_sdf.normal_eps), int(eng.grid.active_levels),
               tuple(p.requires_grad for p in self.parameters()), bytes(static_cfg), self.path_cfg.precision)
        st = self.__dict__.get("_graph_state")
        if st is not None and st[0] != key:
            st = None
            self.__dict__["_graph_state"] = None  # frees the old graph and its private memory pool
        if st is None:
            seen_before = self.__dict__.get("_graph_last_key") == key
            self.__dict__["_graph_last_key"] = key
            if not seen_before:  # first step with this key: do not capture yet
                return self.fused_train_step(data, loss_cfg, after_backward=after_backward)
            static = {k: v.clone() for k, v in tensors.items()}
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            eng.dynamic_scalars = True
            try:
                eng.set_dynamic_scalars(self.progress, loss_cfg)
                side.wait_stream(cur)
                with torch.cuda.stream(side):  # eager warm-up off the capture stream (lazy loads, smem attributes)
                    for _ in range(2):
                        self.fused_train_step(static, loss_cfg, after_backward=after_backward)
                cur.wait_stream(side)
                # kernels captured from this stream inherit its access-policy window (L2 residency of the dense table levels)
                eng.apply_l2_window(self.neural_sdf.tcnn_encoding.params)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    losses = self.fused_train_step(static, loss_cfg, after_backward=after_backward)
            finally:
                eng.dynamic_scalars = False
            grads = {n: p.grad for n, p in self.named_parameters() if p.grad is not None}
            st = (key, graph, static, losses, grads)
            self.__dict__["_graph_state"] = st
        _, graph, static, losses, grads = st
        for k, v in tensors.items():
            static[k].copy_(v, non_blocking=True)
        eng.set_dynamic_scalars(self.progress, loss_cfg)
        graph.replay()
        for n, p in self.named_parameters():
            if n in grads:
                p.grad = grads[n]
        return losses

    @torch.no_grad()
    def inference(self, data, per_sample=True, shard=None):
        """NeuralLumen/model.py:60-111: full-image render in chunks of rand_rays_val rays, eval outputs + *_map.
        ``per_sample=False`` drops the [B,HW,128,.] debug tensors (dists / weights / gradients) that the reference
        concatenates but nothing downstream reads.
        ``shard=(rank, world)``: every rank renders one contiguous range of the frame's rays (dist.shard_rays) and the
        per-ray outputs are all-gathered (torch.distributed, NCCL over NVLink; the reference gathers whole frames the
        same way, projects/nerf/utils/misc.py:25-34), so every rank returns the full maps.  ``shard=True`` takes rank
        and world size from the default process group."""
        self.eval()
        pose = data["pose"]
        B = pose.shape[0]
        H, W = self.image_size_val
        rank, world = 0, 1
        if shard is True:
            import torch.distributed as dist
            rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        elif shard:
            rank, world = int(shard[0]), int(shard[1])
        if world > 1:
            from .dist import shard_rays
            r0, r1 = shard_rays(H * W, rank, world)
            ray_idx = torch.arange(r0, r1, device=pose.device, dtype=torch.int64)[None].expand(B, -1).contiguous()
            n_loc = r1 - r0
        else:
            ray_idx, n_loc = None, H * W
        c, r, l, norm = self._rays(pose, data["intr"], data["pose_light"], self.image_size_val, ray_idx)
        c, r, l, norm = c.view(B, n_loc, 3), r.view(B, n_loc, 3), l.view(B, n_loc, 3), norm.view(B, n_loc, 1)
        chunks = []
        for s in range(0, n_loc, self.rand_rays_val):
            e = min(n_loc, s + self.rand_rays_val)
            o = self.render_rays_lumen(c[:, s:e], r[:, s:e], l[:, s:e], stratified=False)
            o["depth"] = o.pop("_dist") / norm[:, s:e]
            if not per_sample:
                for k in ("dists", "weights", "gradients"):
                    o.pop(k, None)
            chunks.append(o)
        if chunks:
            output = {k: torch.cat([ch[k] for ch in chunks], dim=1) for k, v in chunks[0].items() if v is not None}
        else:  # a rank whose range is empty (more ranks than rays): zero-length tensors with the regular keys
            output = {k: v for k, v in self._empty_render(B, 0, pose.device).items() if v is not None}
            output["depth"] = output.pop("_dist")
            if not per_sample:
                for k in ("dists", "weights", "gradients"):
                    output.pop(k, None)
        if world > 1:
            output = self._gather_rays(output, H * W, world)
        rot = pose[..., :3, :3]
        normal_cam = -output["gradient"] @ rot.transpose(-1, -2)
        to_img = lambda x: x.unflatten(dim=1, sizes=(H, W)).moveaxis(-1, 1)  # misc.py:110-117
        output.update(rgb_map=to_img(output["rgb"]), opacity_map=to_img(output["opacity"]),
                      depth_map=to_img(output["depth"]), normal_map=to_img(normal_cam))
        for key in ("o_r", "o_s", "o_re"):
            if key in output:
                output[key + "_map"] = to_img(output[key])
        if self.flag_light_visibility:  # NeuralLumen/model.py:78-83
            for key in ("visibility", "normal_x_light", "pseudo_shading", "inter_dist", "inter_mask"):
                output[key + "_map"] = to_img(output[key]).float()
        return output

    @staticmethod
    def _gather_rays(output, n_rays, world):
        """all-gather of per-ray outputs [B, n_local, ...] over the ranks' contiguous ray ranges -> [B, n_rays, ...].
        Ranges are padded to the common shard size so that one all_gather_into_tensor per key does it."""
        import torch.distributed as dist
        from .dist import shard_rays
        per = shard_rays(n_rays, 0, world)[1]
        full = {}
        for k in sorted(output):  # same key order on every rank
            v = output[k]
            was_bool = v.dtype == torch.bool
            v = v.to(torch.uint8) if was_bool else v
            B, n_loc = v.shape[:2]
            tail = tuple(v.shape[2:])
            send = v.new_zeros((B, per) + tail)
            send[:, :n_loc] = v
            recv = v.new_empty((world, B, per) + tail)
            dist.all_gather_into_tensor(recv.view(-1), send.view(-1).contiguous())
            g = recv.transpose(0, 1).reshape((B, world * per) + tail)[:, :n_rays]
            full[k] = g.bool() if was_bool else g.contiguous()
        return full
