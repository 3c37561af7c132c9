"""Fused training losses (SURVEY.md section 8 row a13 / north star item (e)).

``loss_cfg_from_trainer`` reads the same keys as the reference trainer
(/root/reference/projects/NeuralLumen/trainer.py:56-71,133-149: cfg.trainer.loss_weight, para_intrinsic_loss,
para_regularize_re_loss); ``fused_losses`` is what a maintainer calls instead of ``Trainer._compute_loss`` +
``_get_total_loss`` when the model was run through ``Model.fused_train_step``.
"""
from . import _lib


def loss_cfg_from_trainer(cfg_trainer, weights=None):
    """`weights`: the trainer's LIVE `self.weights` dict (projects/nerf/trainers/base.py:47 keeps the non-zero entries of
    cfg.trainer.loss_weight; projects/neuralangelo/trainer.py:56-63 re-schedules `curvature` every iteration when
    coarse-to-fine is on).  Pass it from inside the trainer; without it the static YAML weights are used."""
    w = cfg_trainer.loss_weight
    c = _lib.LossCfg()

    def weight(name):
        if weights is not None:
            return float(weights.get(name, 0.0))  # a loss that is not in self.weights does not enter the total
        return float(getattr(w, name, 0.0))

    c.w_render = weight("render")
    c.w_eikonal = weight("eikonal")
    c.w_curvature = weight("curvature")
    c.w_intrinsic = weight("intrinsic")
    c.w_regularize_re = weight("regularize_re")
    c.has_intrinsic = int(hasattr(w, "intrinsic"))
    pi = getattr(cfg_trainer, "para_intrinsic_loss", None)
    rs = getattr(pi, "weight_map_range_shading", (0.25, 1.0)) if pi is not None else (0.25, 1.0)
    rv = getattr(pi, "weight_map_range_visibility", (0.25, 1.0)) if pi is not None else (0.25, 1.0)
    c.range_sha[0], c.range_sha[1] = float(rs[0]), float(rs[1])
    c.range_vis[0], c.range_vis[1] = float(rv[0]), float(rv[1])
    c.factor_ref = float(getattr(pi, "factor_ref", 1.0)) if pi is not None else 1.0
    c.factor_sha = float(getattr(pi, "factor_sha", 1.0)) if pi is not None else 1.0
    pr = getattr(cfg_trainer, "para_regularize_re_loss", None)
    c.factor_negative = float(getattr(pr, "factor_negative", 10.0)) if pr is not None else 10.0
    c.factor_positive = float(getattr(pr, "factor_positive", 1.0)) if pr is not None else 1.0
    c.exponent_positive = float(getattr(pr, "exponent_positive", 1.0)) if pr is not None else 1.0
    return c


def losses_to_dict(losses_tensor):
    vals = losses_tensor.tolist()
    return {name: vals[i] for i, name in enumerate(_lib.LOSS_NAMES)}
