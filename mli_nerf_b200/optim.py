"""``FusedAdamW``: torch.optim.AdamW semantics (the optimizer the reference builds from ``cfg.optim``:
/root/reference/projects/neuralangelo/configs/base.yaml:117-121, imaginaire/trainers/utils/get_trainer.py:106-150) with
the update of every parameter of a group done by ONE ``mli_adamw_step_batch`` launch -- a single pass over (param,
grad, exp_avg, exp_avg_sq), 28 bytes per parameter, i.e. 10.2 GB of HBM traffic per step for the 365 M floats of the
hash table.  ``grad_scale`` folds the 1/world_size of a sum-all-reduce into the same pass.

Drop-in: ``optim = FusedAdamW(model.get_param_groups(cfg.optim), lr=1e-3, weight_decay=1e-2)``; state_dict keys
(`step`, `exp_avg`, `exp_avg_sq`) are those of torch.optim.AdamW, so the reference's checkpoints resume unchanged.
"""
import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        # the inert torch.optim.AdamW group keys ride along so that a state_dict saved here loads into torch's AdamW
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                                      decoupled_weight_decay=True, grad_scale=grad_scale))

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.AdamW checkpoint (the reference's): its param_groups carry no `grad_scale` (kept from
        this optimizer) and may carry amsgrad / maximize / capturable flags -- the two that would change the update
        rule are refused instead of being silently ignored."""
        for g in state_dict["param_groups"]:
            if g.get("amsgrad", False) or g.get("maximize", False) or not g.get("decoupled_weight_decay", True):
                raise ValueError("FusedAdamW: checkpoints saved with amsgrad=True, maximize=True or coupled weight decay "
                                 "are not supported")
        scales = [g.get("grad_scale", 1.0) for g in self.param_groups]
        super().load_state_dict(state_dict)
        for g, s in zip(self.param_groups, scales):
            g.setdefault("grad_scale", s)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            by_step = {}  # parameters that share a step count share the bias corrections -> one launch
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.MliError("FusedAdamW updates contiguous fp32 CUDA parameters only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                by_step.setdefault(int(st["step"]), []).append((p, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"]))
            for step, items in by_step.items():
                for s in range(0, len(items), _lib.ADAMW_MAX_TENSORS):
                    part = items[s:s + _lib.ADAMW_MAX_TENSORS]
                    descs = (_lib.AdamwDesc * len(part))()
                    for d, (p, g, m, v) in zip(descs, part):
                        d.param, d.grad, d.exp_avg, d.exp_avg_sq, d.n = p.data_ptr(), g.data_ptr(), m.data_ptr(), \
                            v.data_ptr(), p.numel()
                    _lib.call("mli_adamw_step_batch", _lib.C.addressof(descs), len(part), group["lr"], b1, b2,
                              group["eps"], group["weight_decay"], step, group.get("grad_scale", 1.0))
        return loss
