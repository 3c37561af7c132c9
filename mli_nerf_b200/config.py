"""Resolved configs of the shipped NeuralLumen experiments, as attribute trees.

The drop-in ``Model(cfg_model, cfg_data)`` reads the same keys as the reference model
(/root/reference/projects/NeuralLumen/model.py:19-58, projects/neuralangelo/model.py:31-60), so the reference's own
``imaginaire.config.Config`` objects work unchanged.  This module only exists so the path can be exercised where the
reference's YAML loader is not present (tests, bench, smoke on the GPU box): it reproduces the values of
projects/neuralangelo/configs/base.yaml + projects/NeuralLumen/configs/*.yaml after ``_parent_`` inheritance
(SURVEY.md Appendix D).
"""
from types import SimpleNamespace as NS


def _ns(d):
    if isinstance(d, dict):
        return NS(**{k: _ns(v) for k, v in d.items()})
    return d


_BOUNDS = {
    "syn_hotdog": dict(bounding_type="unit_sphere"),
    "NRHints_Pikachu": dict(bounding_type="unit_sphere"),
    "rene_savannah": dict(bounding_type="box", bounding_box_aabb=[-0.66, -0.516, -0.18, 0.66, 0.42, 0.3]),
}
_IMAGE = {"syn_hotdog": [512, 512], "NRHints_Pikachu": [512, 512], "rene_savannah": [270, 360]}
_WHITE = {"syn_hotdog": True, "NRHints_Pikachu": False, "rene_savannah": False}


def experiment(name="syn_hotdog_b", dict_size=22, rand_rays=2048, taps=4):
    """cfg with .model / .data / .trainer / .optim / .max_iter for ``<scene>_<a|b>``."""
    scene, stage = name.rsplit("_", 1)
    if scene not in _BOUNDS or stage not in ("a", "b"):
        raise KeyError(name)
    rgb = dict(mlp=dict(num_layers=4, hidden_dim=256, skip=[], activ="relu_", activ_params={}, weight_norm=True),
               mode="idr", encoding_view=dict(type="spherical", levels=3))
    if stage == "b":
        rgb.update(network_mode="rgb_r_s", shading_dim=1)
    model = dict(
        type="mli_nerf_b200.model",
        object=dict(
            sdf=dict(
                mlp=dict(num_layers=1, hidden_dim=256, skip=[], activ="softplus", activ_params=dict(beta=100),
                         geometric_init=True, weight_norm=True, out_bias=0.5, inside_out=False),
                encoding=dict(type="hashgrid", levels=16,
                              hashgrid=dict(min_logres=5, max_logres=11, dict_size=dict_size, dim=8, range=[-2, 2]),
                              coarse2fine=dict(enabled=(stage == "a"), init_active_level=8, step=5000)),
                gradient=dict(mode="numerical", taps=taps)),
            rgb=rgb,
            s_var=dict(init_val=3.0, anneal_end=0.1)),
        background=dict(enabled=False, white=_WHITE[scene]),
        render=dict(rand_rays=rand_rays, rand_rays_val=20000,
                    num_samples=dict(coarse=64, fine=16, background=32), num_sample_hierarchy=4, stratified=True),
        appear_embed=dict(enabled=False, dim=8),
        light_visibility=dict(enabled=False, camera_ray_type="blend_z_sphere_tracing", type="sphere_tracing",
                              visibility_bounding_type="sphere", visibility_sphere_radius=0.95),
    )
    data = dict(train=dict(image_size=_IMAGE[scene], batch_size=1), val=dict(image_size=_IMAGE[scene], batch_size=1),
                **_BOUNDS[scene])
    weights = dict(render=1.0, eikonal=0.1, curvature=5e-4)
    trainer = dict(loss_weight=weights)
    if stage == "b":
        weights.update(intrinsic=1.0, regularize_re=1.0)
        trainer.update(para_intrinsic_loss=dict(weight_map_range_shading=[0.0, 1.0],
                                                weight_map_range_visibility=[0.0, 1.0], factor_ref=1.0,
                                                factor_sha=1.0),
                       para_regularize_re_loss=dict(factor_negative=10.0, factor_positive=1.0, exponent_positive=1.0),
                       partial_grad=["neural_rgb"])
    optim = dict(type="AdamW", params=dict(lr=1e-3, weight_decay=1e-2), sched=dict(warm_up_end=5000))
    if stage == "b":
        optim["partial_training"] = ["neural_rgb"]
    return _ns(dict(model=model, data=data, trainer=trainer, optim=optim, max_iter=500000))
