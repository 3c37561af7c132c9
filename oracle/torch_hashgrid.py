"""TEST INFRASTRUCTURE ONLY -- fp32 torch stand-in for ``tinycudann.Encoding(3, {"otype": "HashGrid", ...})``.

PARITY UNPINNED.  The reference reaches tiny-cuda-nn (NVlabs/tiny-cuda-nn, version not pinned: the reference
has no requirements/lock file, README.md:13-17) at projects/neuralangelo/utils/modules.py:42-50,84-86.  Its
source is not under /root/reference, so this file restates the *published* algorithm of tcnn's
``GridEncodingTemplated`` / ``kernel_grid`` (grid.h) from knowledge of the upstream project:

  scale_l   = exp2f(l * log2f(per_level_scale)) * base_resolution - 1            (float32 in tcnn; see level_table)
  res_l     = (uint32) ceilf(scale_l) + 1
  size_l    = min(next_multiple(res_l**3 (clamped to UINT32_MAX/2), 8), 2**log2_hashmap_size)
  pos       = fmaf(scale_l, x, 0.5f); cell = floorf(pos); w = pos - cell; cell -> (uint32)(int)cell
  index     = sum_d cell_d * stride_d while stride <= size_l (stride *= res_l);
              if size_l < stride: index = cell_x*1 ^ cell_y*2654435761 ^ cell_z*805459861   (uint32 wrap)
              index %= size_l
  out[l*F+f] = sum_{corner c = 0..7 (bit d of c selects +1 along dim d)} prod_d (w_d or 1-w_d) * table[off_l + index_c, f]
  params    : flat fp32, level-major, entry, feature-minor; init U(-1e-4, 1e-4)

Real tcnn stores params/outputs in fp16 on the GPU; the oracle (and the product's fp32 mode) are fp32, as the
task's north star specifies.
"""
import math

import numpy as np
import torch

PRIME_Y = 2654435761
PRIME_Z = 805459861


def level_table(n_levels, log2_hashmap_size, base_resolution, per_level_scale):
    """Per-level (scale, res, size, offset, hashed) following tcnn's constructor + grid_scale/grid_resolution."""
    # tcnn: exp2f(l * log2f(s)) * base - 1 in float32 on the device (last-ulp rounding is platform dependent);
    # here: the same formula in float64 on the float32-rounded per_level_scale, rounded once to float32.
    log2_pls = math.log2(float(np.float32(per_level_scale)))
    levels, offset = [], 0
    for lv in range(n_levels):
        scale = np.float32(math.exp2(lv * log2_pls) * base_resolution - 1.0)
        res = int(math.ceil(float(scale))) + 1
        max_params = (2 ** 32 - 1) // 2
        dense = res ** 3 if float(res) ** 3 <= float(max_params) else max_params
        size = min(((dense + 7) // 8) * 8, 1 << log2_hashmap_size)
        # replicate grid_index(): the dense stride walk stops once stride > size
        stride, hashed = 1, False
        for _ in range(3):
            if stride <= size:
                stride *= res
        hashed = size < stride
        levels.append(dict(scale=float(scale), res=res, size=size, offset=offset, hashed=hashed))
        offset += size
    return levels, offset


def corner_indices(x01, lv):
    """int64 [M,8] table row (without level offset) and fp32 [M,8] trilinear weight for one level."""
    scale = torch.tensor(lv["scale"], dtype=torch.float32)
    # fmaf(scale, x, 0.5): emulate the single rounding through float64 (exact product of two fp32 fits fp64)
    pos = (x01.double() * scale.double() + 0.5).float()
    cell_f = torch.floor(pos)
    w = pos - cell_f
    cell = cell_f.to(torch.int32).to(torch.int64) & 0xFFFFFFFF  # (uint32)(int)floor
    res, size = lv["res"], lv["size"]
    idx_all, wt_all = [], []
    for c in range(8):
        weight = torch.ones_like(w[:, 0])
        g = []
        for d in range(3):
            if (c >> d) & 1:
                weight = weight * w[:, d]
                g.append((cell[:, d] + 1) & 0xFFFFFFFF)
            else:
                weight = weight * (1 - w[:, d])
                g.append(cell[:, d])
        if lv["hashed"]:
            idx = (g[0] ^ ((g[1] * PRIME_Y) & 0xFFFFFFFF) ^ ((g[2] * PRIME_Z) & 0xFFFFFFFF)) & 0xFFFFFFFF
        else:
            idx, stride = torch.zeros_like(g[0]), 1
            for d in range(3):
                if stride <= size:
                    idx = (idx + g[d] * stride) & 0xFFFFFFFF
                    stride *= res
        idx_all.append(idx % size)
        wt_all.append(weight)
    return torch.stack(idx_all, 1), torch.stack(wt_all, 1)


class _HashEncode(torch.autograd.Function):
    """out[:, l*F:(l+1)*F] = sum_c w_c * table[off_l + idx_c]  with a backward that scatter-adds straight into ONE dense
    gradient (index_add_ per level and corner), the way tcnn's backward kernel does.  Plain advanced indexing gives the
    same numbers but makes autograd allocate and accumulate a full table-sized zero tensor per indexing op
    (16 levels x 8 corners x 1.46 GB at T = 2^22), which is minutes of pure memset on a CPU."""

    @staticmethod
    def forward(ctx, params, x01, levels, feat):
        table = params.detach().view(-1, feat)
        x01 = x01.detach().float()
        outs = []
        for lv in levels:
            idx, wt = corner_indices(x01, lv)
            acc = torch.zeros(x01.shape[0], feat, dtype=torch.float32)
            for c in range(8):  # same corner order as tcnn's fma chain
                acc = acc + wt[:, c:c + 1] * table[lv["offset"] + idx[:, c]]
            outs.append(acc)
        ctx.save_for_backward(x01)
        ctx.levels, ctx.feat, ctx.n = levels, feat, params.numel()
        return torch.cat(outs, dim=-1)

    @staticmethod
    def backward(ctx, d_out):
        (x01,) = ctx.saved_tensors
        feat = ctx.feat
        grad = torch.zeros(ctx.n // feat, feat, dtype=torch.float32)
        d_out = d_out.float()
        for l, lv in enumerate(ctx.levels):
            idx, wt = corner_indices(x01, lv)
            d = d_out[:, l * feat:(l + 1) * feat]
            for c in range(8):
                grad.index_add_(0, lv["offset"] + idx[:, c], wt[:, c:c + 1] * d)
        return grad.view(-1), None, None, None


def hash_encode(params, x01, levels, feat):
    """Differentiable (w.r.t. ``params``) hash-grid encoding of x01 [M,3] -> [M, L*F]."""
    return _HashEncode.apply(params, x01, levels, feat)


class TorchHashGrid(torch.nn.Module):
    """Drop-in for tcnn.Encoding (forward(x[M,3] in [0,1]) -> [M, L*F]); differentiable w.r.t. ``params``."""

    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=torch.float32):
        super().__init__()
        assert n_input_dims == 3 and encoding_config["otype"] == "HashGrid"
        self.L = int(encoding_config["n_levels"])
        self.F = int(encoding_config["n_features_per_level"])
        self.levels, n_entries = level_table(self.L, int(encoding_config["log2_hashmap_size"]),
                                             int(encoding_config["base_resolution"]),
                                             float(encoding_config["per_level_scale"]))
        self.n_output_dims = self.L * self.F
        gen = torch.Generator().manual_seed(seed)
        init = (torch.rand(n_entries * self.F, generator=gen, dtype=torch.float32) * 2 - 1) * 1e-4
        self.params = torch.nn.Parameter(init)

    def forward(self, x01):
        return hash_encode(self.params, x01, self.levels, self.F)
