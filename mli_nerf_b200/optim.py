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


class ShardedTableAdamW:
    """AdamW for ray-sharded training whose hash-table gradient was exchanged by REDUCE-SCATTER
    (`GradReducer(table_mode="reduce_scatter")`, mli_nerf_b200/dist.py): every rank updates only the table shards it owns
    -- parameter entries, first and second moments of 1/W of the table: 1/W of the 28 B/parameter optimizer pass and of
    the moment memory -- then the updated PARAMETER shards are all-gathered in place into every replica (NCCL over
    NVLink; input = this rank's slice of the output slab).  All other parameters (MLPs, s_var: ~3.6 MB, all-reduced
    gradients) are updated redundantly on every rank by a regular FusedAdamW.

    Same arithmetic as the reference's all-reduce + torch.optim.AdamW (base.yaml:117-121): every table entry is updated
    exactly once, by its owner, from the same mean gradient, with the same bias corrections; replicas stay identical
    because they all receive the owner's bytes.  `state_dict()` is rank-local (each rank holds the moments of its own
    shards); `gather_state()` rebuilds the dense torch.optim.AdamW moments for a checkpoint."""

    def __init__(self, reducer, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        import torch.distributed as dist
        self.reducer, self.dist = reducer, dist
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        params = list(params)
        self.table = max(params, key=lambda p: p.numel())
        if self.table.numel() < (1 << 22):
            raise ValueError("ShardedTableAdamW: no large table parameter among `params`")
        self.small = FusedAdamW([p for p in params if p is not self.table], lr=lr, betas=betas, eps=eps,
                                weight_decay=weight_decay)
        self.step_count = 0
        self.state = {}  # (elem_begin, elem_end) -> (exp_avg, exp_avg_sq)

    @torch.no_grad()
    def step(self):
        red = self.reducer
        shards = getattr(red, "_last_shards", None)
        if not shards:
            raise _lib.MliError("ShardedTableAdamW.step: no reduce-scattered table gradient (call the fused step with "
                                "after_backward=reducer.exchange_grads first)")
        table = self.table.data.view(-1)
        if not table.is_cuda or table.dtype != torch.float32 or not table.is_contiguous():
            raise _lib.MliError("ShardedTableAdamW updates a contiguous fp32 CUDA table only (no CPU fallback)")
        # the dense .grad of the table holds this rank's UNREDUCED gradient in this mode: keep the dense optimizer off it
        g_table, self.table.grad = self.table.grad, None
        self.small.step()
        self.table.grad = g_table
        self.step_count += 1
        descs = (_lib.AdamwDesc * len(shards))()
        for d, (a, b, g) in zip(descs, shards):
            st = self.state.get((a, b))
            if st is None:
                st = (torch.zeros(b - a, dtype=torch.float32, device=table.device),
                      torch.zeros(b - a, dtype=torch.float32, device=table.device))
                self.state[(a, b)] = st
            d.param, d.grad, d.exp_avg, d.exp_avg_sq, d.n = table.data_ptr() + 4 * a, g.data_ptr(), st[0].data_ptr(), \
                st[1].data_ptr(), b - a
        _lib.call("mli_adamw_step_batch", _lib.C.addressof(descs), len(shards), self.lr, self.betas[0], self.betas[1],
                  self.eps, self.weight_decay, self.step_count, 1.0)
        # all-gather of the updated parameter shards, slab by slab, in place (slab = W consecutive equal shards)
        W, r = red.world, red.rank
        for a, b, _ in shards:
            n = b - a
            slab = table[a - r * n:a + (W - r) * n]
            self.dist.all_gather_into_tensor(slab, table[a:b])

    def zero_grad(self, set_to_none=True):
        self.small.zero_grad(set_to_none=set_to_none)
        self.table.grad = None

    def state_dict(self):
        return {"step": self.step_count, "small": self.small.state_dict(),
                "table_shards": {k: (m.clone(), v.clone()) for k, (m, v) in self.state.items()}}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.small.load_state_dict(sd["small"])
        self.state = {k: (m.clone(), v.clone()) for k, (m, v) in sd["table_shards"].items()}

    @torch.no_grad()
    def gather_state(self):
        """Dense (exp_avg, exp_avg_sq) of the table on every rank, as torch.optim.AdamW keeps them (checkpointing)."""
        W, r = self.reducer.world, self.reducer.rank
        n_tab = self.table.numel()
        dense = [torch.zeros(n_tab, dtype=torch.float32, device=self.table.device) for _ in range(2)]
        for (a, b), st in sorted(self.state.items()):
            n = b - a
            for k in range(2):
                self.dist.all_gather_into_tensor(dense[k][a - r * n:a + (W - r) * n], st[k])
        return dense[0].view_as(self.table), dense[1].view_as(self.table)
