"""``Encoding``: the tcnn.Encoding(3, {"otype": "HashGrid", ...}) surface the reference binds
(/root/reference/projects/neuralangelo/utils/modules.py:42-50,84-86), backed by libmli_b200's sm_100a kernels.

Same constructor arguments, same ``params`` layout (flat fp32, level-major / entry / feature-minor, U(-1e-4,1e-4) init),
``forward(x[M,3] in [0,1]) -> [M, n_levels*n_features]`` and gradient w.r.t. ``params``.  Outputs are fp32 (real tcnn
returns fp16; the north star asks for an fp32 oracle-parity mode).
"""
import torch

from . import _lib


class _EncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x01, params, enc):
        x01 = x01.detach().contiguous().float()
        M = x01.shape[0]
        out = torch.empty(M, enc.n_output_dims, dtype=torch.float32, device=x01.device)
        _lib.call("mli_hashgrid_fwd", enc.grid, params, x01, M, out, enc.n_output_dims)
        ctx.save_for_backward(x01)
        ctx.enc = enc
        ctx.n_params = params.numel()
        return out

    @staticmethod
    def backward(ctx, d_out):
        (x01,) = ctx.saved_tensors
        enc = ctx.enc
        grad = torch.zeros(ctx.n_params, dtype=torch.float32, device=x01.device)
        d_out = d_out.contiguous()
        _lib.call("mli_hashgrid_bwd", enc.grid, x01, x01.shape[0], d_out, d_out.shape[1], grad)
        return None, grad, None


class Encoding(torch.nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=torch.float32):
        super().__init__()
        if n_input_dims != 3 or encoding_config.get("otype") != "HashGrid":
            raise NotImplementedError("only the 3-D HashGrid encoding is on the hot path")
        self.n_input_dims = 3
        self.encoding_config = dict(encoding_config)
        self.n_levels = int(encoding_config["n_levels"])
        self.n_features = int(encoding_config["n_features_per_level"])
        self.n_output_dims = self.n_levels * self.n_features
        self.grid = _lib.make_grid(self.n_levels, self.n_features, int(encoding_config["log2_hashmap_size"]),
                                   int(encoding_config["base_resolution"]), float(encoding_config["per_level_scale"]))
        gen = torch.Generator().manual_seed(seed)
        n = int(self.grid.n_entries) * self.n_features
        self.params = torch.nn.Parameter((torch.rand(n, generator=gen, dtype=torch.float32) * 2 - 1) * 1e-4)

    def level_table(self):
        return [dict(scale=l.scale, res=l.res, size=l.size, offset=l.offset, hashed=bool(l.hashed))
                for l in list(self.grid.level)[:self.n_levels]]

    def forward(self, x):
        return _EncodeFn.apply(x, self.params, self)
