"""Host-side orchestration of the per-ray render path over the C ABI (libmli_b200.so).

This file is plumbing: it owns the HBM layout (which buffer each kernel reads/writes), issues the C-ABI calls in
order on the current CUDA stream and hands torch-allocated device pointers across the boundary.  No arithmetic of the
path happens here and there is no CPU fallback.

Reference call stack being replaced (relative to /root/reference/):
  projects/NeuralLumen/model.py:232-336 render_rays_lumen, :338-403 render_rays_object_lumen
  projects/neuralangelo/model.py:420-515 bounds / sampling / NeuS alphas
  projects/neuralangelo/utils/modules.py:68-178 NeuralSDF forward + numerical gradients
  projects/NeuralLumen/utils/modules.py:106-174 LumenRGB forward
  + the autograd backward of all of it (projects/NeuralLumen/trainer.py:189-206).

HBM layout, fp32 mode (row-major; M = R*N samples, P = 1+taps stencil planes)
  X0   [P*M, 144]  = [hash encoding 0:128 | xyz 128:131 | 0-pad]          input of SDF layer 0 (K padded to 16)
  H0   [P*M, 256]  = softplus100(X0 W0^T + b0)                            (plane 0 = centre rows)
  sdf  [P*M]       = H0 w_sdf + b_sdf
  XH   [M, 304]    = [feat 0:256 | xyz | SH(view) | normal | SH(light) | 0-pad]   input of the fused head layer 0
  A1..A4 [M, nh*256] hidden activations of the nh heads side by side (batched block-diagonal layers 1..3)
  S    [M, 8]      per-sample head outputs (rgb3|o_r3|o_s1 for rgb_r_s)

bf16 (tensor-core) mode: every matrix above lives in bf16 "TCL" [rows/128][cols/8][128][8] (csrc/gemm_tcgen05.cu); the
trunk input is the split-bf16 delta-basis matrix Xd [P*M, 2*144] (plane 0 = centre rows, plane i = tap - centre; hi | lo
halves), sdf planes >= 1 hold sdf_tap - sdf_centre, S0 = sigmoid(100 z0) (fp32) and DZ = W0 (x_tap - x_centre) (bf16) are
kept for the backward pass, Am[l] are the relu sign bits of the head activations.  Backward: data-gradient chain first
(main stream), weight-gradient GEMMs deferred to a second stream, table gradient scattered per level group.
"""
import math
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SOFTPLUS100, call

K0_PAD = 144      # 128 + 3 -> multiple of 16
K0_CH = K0_PAD // 8   # chunks of one bf16 half of a trunk input row (the row holds [hi | lo] = 2*K0_CH chunks)
KH_PAD = 304      # 256 + 38 -> multiple of 16
XH_OFF = 256      # first non-feature column of XH
HID = 256


@dataclass
class PathCfg:
    """Resolved hot-path hyper-parameters (what the reference reads from cfg.model / cfg.data / cfg.trainer)."""
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
    network_mode: Optional[str] = "rgb_r_s"
    white_background: bool = True
    anneal_end: float = 0.1
    outside_val: float = 1000.0
    bounding: str = "unit_sphere"
    aabb: Optional[Tuple[float, ...]] = None
    c2f_enabled: bool = False
    precision: int = _lib.PREC_FP32

    @property
    def n_samples(self):
        return self.coarse + self.fine * self.hierarchy

    @property
    def growth_rate(self):
        return float(math.exp((math.log(2 ** self.max_logres) - math.log(2 ** self.min_logres)) / (self.n_levels - 1)))

    def resolutions(self):
        return [int(math.floor(2 ** self.min_logres * self.growth_rate ** lv)) + 1 for lv in range(self.n_levels)]


def head_layout(mode):
    """(state-dict name, input kind, out_dim, sigmoid) per head, in the channel order of the composite kernel."""
    table = {
        "rgb_r_s": [("mlp", "full", 3, True), ("mlp_r", "geo", 3, True), ("mlp_s", "geo_l", 1, True)],
        "rgb_r": [("mlp", "full", 3, True), ("mlp_r", "geo", 3, True)],
        "r_s": [("mlp_r", "geo", 3, True), ("mlp_s", "full", 3, False)],
        "r_s_re": [("mlp_r", "geo", 3, True), ("mlp_s", "geo_l", 3, True), ("mlp_re", "full", 3, True)],
        "rgb": [("mlp", "full", 3, True)],
        None: [("mlp", "full", 3, True)],
    }
    if mode not in table:
        raise NotImplementedError(f"unknown network_mode {mode}")
    return table[mode]


def _col_map(kind):
    pts, view, nrm = [256, 257, 258], list(range(259, 275)), [275, 276, 277]
    feat, light = list(range(0, 256)), list(range(278, 294))
    return {"full": pts + view + nrm + feat + light, "geo": pts + nrm + feat, "geo_l": pts + nrm + feat + light}[kind]


class RenderEngine:
    def __init__(self, cfg: PathCfg, device="cuda"):
        if cfg.hidden != HID or cfg.sh_levels != 3 or cfg.n_levels * cfg.feat_per_level != 128:
            raise NotImplementedError("engine is specialised for hidden=256, SH level 3 and a 128-wide hash encoding")
        _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        self.grid = _lib.make_grid(cfg.n_levels, cfg.feat_per_level, cfg.log2_hashmap_size, 2 ** cfg.min_logres,
                                   cfg.growth_rate)
        self.heads = head_layout(cfg.network_mode)
        self.nh = len(self.heads)
        self.mode = _lib.MODE_BY_NAME[cfg.network_mode]
        self.J = sum(h[2] for h in self.heads)
        self.col_off = [hi * HID for hi, h in enumerate(self.heads) for _ in range(h[2])]
        self.act_mask = 0
        j = 0
        for h in self.heads:
            for _ in range(h[2]):
                self.act_mask |= (1 << j) if h[3] else 0
                j += 1
        self.n_out = {0: 3, 1: 10, 2: 9, 3: 9, 4: 12}[self.mode]
        self.lds = 8 if self.J <= 8 else 12  # row length of the per-sample head outputs S / their gradients
        i32 = dict(dtype=torch.int32, device=self.device)
        self.map_sdf0 = torch.tensor([128, 129, 130] + list(range(128)), **i32)
        self.map_head = [torch.tensor(_col_map(h[1]), **i32) for h in self.heads]
        self.normal_eps = 1.0 / cfg.resolutions()[-1]
        self.active_levels = cfg.n_levels
        self.W = None
        # multi-GPU overlap: called as hook(table_grad_flat, start_elem, end_elem) right after the scatter kernel of a
        # level group has been launched (that slab of the hash-table gradient is final once the kernel completes)
        self.table_grad_hook = None
        self._side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._wg_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.overlap_wgrad = True  # weight-gradient GEMMs on a second stream, concurrent with trunk backward + scatter
        # multi-GPU (set by GradReducer.attach): keep every weight-gradient GEMM back until the hash-grid scatter has been
        # launched, so that the all-reduce of the 1.46 GB table gradient -- which can only start when the scatter, the last
        # link of the data-gradient chain, has produced a slab -- has ~0.9 ms of independent work to hide behind
        self.wgrad_after_scatter = False
        self._wg_keep = []
        self._wg_jobs = None
        self._tg_early = None
        # Schedule scalars that move every iteration -- the s_var anneal ratio (neuralangelo/model.py:492-499) and the loss
        # weights (curvature warm-up, neuralangelo/trainer.py:56-63) -- can live in a small device buffer
        # [anneal, w_render, w_eikonal, w_curvature, w_intrinsic, w_regularize_re] that the composite / loss kernels read:
        # the launches then carry no per-iteration value and ONE captured CUDA graph serves the whole schedule
        # (Model._graphed_train_step switches this on and refreshes the buffer before each replay).
        self.dynamic_scalars = False
        self._dyn = torch.zeros(8, dtype=torch.float32, device=self.device) if self.device.type == "cuda" else None
        self._dyn_last = None
        self.zero_fill_ctas = int(os.environ.get("MLI_ZERO_FILL_CTAS", "128"))  # 1.46 GB: 16 CTAs 1.6 ms, 64: 0.44, 128: 0.26 (step 4.689 / 4.671)
        # bf16 mode: the head stack as ONE on-chip kernel per direction (csrc/heads_fused.cu).  Measured at the bench shape
        # (tools/bench_heads.py, profiles/r02_heads_fused.md): data-gradient chain 435 us fused vs 564 us layer by layer;
        # forward WITHOUT stored activations (inference / no-grad) 493 us vs 579 us; forward that also has to write the
        # four activation matrices + relu masks for the backward pass 685 us vs 579 us -- so a training forward keeps the
        # layer-by-layer GEMMs (MLI_FUSE_HEADS_TRAIN_FWD=1 forces the fused kernel, MLI_FUSE_HEADS=0 disables all of it)
        # L2 persisting window over the DENSE levels of the table (levels whose grid fits the table: 0-5 at T = 2^22, 119 MB
        # fp32) on the stream the encode kernels run on: MLI_L2_PERSIST=<hit ratio in (0,1]>, 0 = off (default; measured
        # A/B in profiles/r02_summary.md)
        self.l2_persist = float(os.environ.get("MLI_L2_PERSIST", "0"))
        self._l2_key = None
        self.fuse_heads = os.environ.get("MLI_FUSE_HEADS", "1") == "1"
        self.fuse_heads_train_fwd = os.environ.get("MLI_FUSE_HEADS_TRAIN_FWD", "0") == "1"
        # persistent buffer the table gradient is accumulated in (multi-GPU: the IPC-shared buffer of PeerTableReducer);
        # None: a fresh zero-filled buffer per step
        self.table_grad_buffer = None

    # ------------------------------------------------------------------------------------------------------
    def n_table_params(self):
        return int(self.grid.n_entries) * self.cfg.feat_per_level

    def set_dynamic_scalars(self, progress, lcfg):
        """Refresh the device-resident schedule scalars (stream-ordered copy on the current stream)."""
        # a fresh pageable host tensor per call: the driver stages it before the copy call returns, so the host may run any
        # number of steps ahead of the GPU without a later step's values overtaking an earlier step's copy (a reused
        # pinned buffer would race exactly that way under graph replay)
        vals = (min(progress / self.cfg.anneal_end, 1.0), lcfg.w_render, lcfg.w_eikonal, lcfg.w_curvature,
                lcfg.w_intrinsic, lcfg.w_regularize_re, 0.0, 0.0)
        if vals == self._dyn_last:  # static part of training: nothing to send, replays stay back to back
            return
        self._dyn.copy_(torch.tensor(vals, dtype=torch.float32))
        self._dyn_last = vals

    def start_table_grad_zero(self):
        """Zero-fill of the 1.46 GB hash-table gradient buffer, issued on a side stream so that it overlaps the
        latency-bound sampling rounds of the forward pass instead of sitting in front of the scatter kernel.  The fused
        train step calls this before sampling; backward() picks the buffer up (and falls back to a plain zero-fill)."""
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            tg = self.table_grad_buffer if self.table_grad_buffer is not None else \
                torch.empty(self.n_table_params(), dtype=torch.float32, device=self.device)
            # a persistent grid of a few CTAs, not torch's full-grid fill: that one occupied every SM slot for its 0.2 ms
            # and the main stream's small sampling kernels queued behind it (bench: sample_coarse 4.7 us -> 160 us)
            call("mli_zero_fill_background", tg, tg.numel() * 4, self.zero_fill_ctas)
        tg.record_stream(cur)
        self._tg_early = tg

    def _flush_later(self, later, after_event=None):
        """Launch the deferred weight-gradient work.  Tensor-core mode: on the weight-gradient stream, ordered after
        `after_event` (recorded on the main stream where the operands became final), so that these HBM-bound GEMMs run
        concurrently with the latency-bound hash-grid scatter and the trunk backward on the main stream."""
        if not later:
            return
        if self.tc and self.overlap_wgrad and self._wg_stream is not None:
            if after_event is not None:
                self._wg_stream.wait_event(after_event)
            with torch.cuda.stream(self._wg_stream):
                for fn in later:
                    fn()
            # the closures hold the operand tensors: keep them alive until the streams have joined, or the caching
            # allocator could hand their memory to a main-stream allocation while the second stream still reads it
            self._wg_keep.extend(later)
        else:
            for fn in later:
                fn()
        later.clear()

    def _take_table_grad(self):
        if self._tg_early is not None:
            tg, self._tg_early = self._tg_early, None
            torch.cuda.current_stream().wait_stream(self._side)
            return tg
        if self.table_grad_buffer is not None:
            return self.table_grad_buffer.zero_()
        return self._z(self.n_table_params())

    def level_groups(self, max_groups=4):
        """Level ranges launched separately by the table-gradient scatter when a multi-GPU hook wants the slabs early.
        NCCL's all-reduce is markedly more efficient on large messages (measured at N = 2: 2.6 ms for the whole 1.46 GB,
        3.3 ms in 11 slabs), so the levels are grouped into at most `max_groups` slabs of similar size.
        -> [(level_begin, level_end, elem_begin, elem_end)]"""
        F, lv = self.cfg.feat_per_level, self.grid.level
        n = self.cfg.n_levels
        total = int(self.grid.n_entries)
        target = total / max_groups
        bounds, acc = [0], 0
        for l in range(n):
            acc += int(lv[l].size)
            if acc >= target * len(bounds) and l + 1 < n and len(bounds) < max_groups:
                bounds.append(l + 1)
        bounds.append(n)
        out = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            end = (int(lv[b].offset) if b < n else total) * F
            out.append((a, b, int(lv[a].offset) * F, int(end)))
        return out

    def set_active_levels(self, active):
        self.active_levels = int(active)
        self.grid.active_levels = int(active) if self.cfg.c2f_enabled else self.cfg.n_levels

    def _f(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    def _z(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32, device=self.device)

    def _tcl(self, rows, chunks, tile=128):
        """bf16 tile-chunk-layout buffer for a [rows, 8*chunks] matrix (see csrc/gemm_tcgen05.cu)."""
        return torch.empty((rows + tile - 1) // tile, chunks, tile, 8, dtype=torch.bfloat16, device=self.device)

    def _to_tcl(self, src, ld, rows, cols, dst, tile=128, chunk0=0, n_chunks=None):
        call("mli_tc_to_tcl", src, ld, rows, cols, dst, tile, dst.shape[1], chunk0,
             n_chunks if n_chunks is not None else (cols + 7) // 8)
        return dst

    @property
    def tc(self):
        return self.cfg.precision == _lib.PREC_BF16

    @property
    def tap_eps(self):  # modules.py:133,158
        return self.normal_eps / math.sqrt(3) if self.cfg.taps == 4 else self.normal_eps

    # ------------------------------------------------------------------------------------------------------
    def pack_weights(self, p):
        """weight_norm reparameterisation into the padded / fused layouts the layer kernels read (per step)."""
        if self.tc:
            return self._pack_weights_tc(p)
        nh, W = self.nh, {}
        W["W0"], W["W0t"] = self._z(HID, K0_PAD), self._z(K0_PAD, HID)
        call("mli_weightnorm_pack", p["neural_sdf.mlp.linears.0.weight_v"], p["neural_sdf.mlp.linears.0.weight_g"], HID,
             131, self.map_sdf0, W["W0"], K0_PAD, W["W0t"], HID, 0)
        W["W1"], W["W1t"] = self._f(HID, HID), self._f(HID, HID)
        call("mli_weightnorm_pack", p["neural_sdf.mlp.linears.1.weight_v"], p["neural_sdf.mlp.linears.1.weight_g"], HID,
             HID, None, W["W1"], HID, W["W1t"], HID, 0)
        W["b0"], W["b1"] = p["neural_sdf.mlp.linears.0.bias"], p["neural_sdf.mlp.linears.1.bias"]
        W["w_sdf"], W["b_sdf"] = p["neural_sdf.mlp.linear_sdf.weight"], p["neural_sdf.mlp.linear_sdf.bias"]
        W["Wh0"], W["Wh0t"] = self._z(nh * HID, KH_PAD), self._z(KH_PAD, nh * HID)
        W["Whl"] = [self._f(nh, HID, HID) for _ in range(3)]
        W["Whlt"] = [self._f(nh, HID, HID) for _ in range(3)]
        W["Wout"] = self._f(self.J, HID)
        j0 = 0
        for hi, (name, kind, odim, _) in enumerate(self.heads):
            pre = f"neural_rgb.{name}.linears."
            k_in = len(_col_map(kind))
            call("mli_weightnorm_pack", p[pre + "0.weight_v"], p[pre + "0.weight_g"], HID, k_in, self.map_head[hi],
                 W["Wh0"], KH_PAD, W["Wh0t"], nh * HID, hi * HID)
            for l in range(3):
                call("mli_weightnorm_pack", p[pre + f"{l + 1}.weight_v"], p[pre + f"{l + 1}.weight_g"], HID, HID, None,
                     W["Whl"][l][hi], HID, W["Whlt"][l][hi], HID, 0)
            call("mli_weightnorm_pack", p[pre + "4.weight_v"], p[pre + "4.weight_g"], odim, HID, None, W["Wout"], HID,
                 None, 0, j0)
            j0 += odim
        W["bh"] = [torch.cat([p[f"neural_rgb.{h[0]}.linears.{l}.bias"] for h in self.heads]) for l in range(4)]
        W["bout"] = torch.cat([p[f"neural_rgb.{h[0]}.linears.4.bias"] for h in self.heads])
        self.W = W
        return W

    def _wn_specs(self):
        """(v name, N, K, col_map tensor, row_off) of every weight-normalised matrix, in a fixed order."""
        specs = [("neural_sdf.mlp.linears.0", HID, 131, self.map_sdf0, 0, "sdf0"),
                 ("neural_sdf.mlp.linears.1", HID, HID, None, 0, "sdf1")]
        j0 = 0
        for hi, (name, kind, odim, _) in enumerate(self.heads):
            pre = f"neural_rgb.{name}.linears."
            specs.append((pre + "0", HID, len(_col_map(kind)), self.map_head[hi], hi * HID, ("h0", hi)))
            for l in range(3):
                specs.append((pre + f"{l + 1}", HID, HID, None, hi * HID, ("hl", l, hi)))
            specs.append((pre + "4", odim, HID, None, j0, ("out", hi)))
            j0 += odim
        return specs

    def _pack_weights_tc(self, p):
        """Tensor-core mode: ONE launch writes every W = g v/||v|| straight into the bf16 TCL layouts (and fp32 W_out);
        one zero-fill provides the K padding.  Replaces 17 pack + 15 layout-conversion launches."""
        nh, W, T = self.nh, {}, {}
        bf = torch.bfloat16
        sizes = dict(W0s=HID * 2 * K0_PAD, W0t_enc=128 * HID, W1=HID * HID, W1t=HID * HID, Wh0=nh * HID * KH_PAD,
                     Wh0t_feat=HID * nh * HID, Wh0t_x=(KH_PAD - XH_OFF) * nh * HID)
        for l in range(3):
            sizes[f"Whl{l}"] = nh * HID * HID
            sizes[f"Whlt{l}"] = nh * HID * HID
        if self.fuse_heads:  # hidden-layer weights (and transposes) in 128-row tiles: [head][N-half][32][128][8]
            for l in range(3):
                sizes[f"Whl128_{l}"] = nh * HID * HID
                sizes[f"Whlt128_{l}"] = nh * HID * HID
        flat = torch.zeros(sum(sizes.values()), dtype=bf, device=self.device)
        off, buf = 0, {}
        for k, n in sizes.items():
            buf[k] = flat[off:off + n]
            off += n
        T["W0s"] = buf["W0s"].view(1, 2 * K0_CH, 256, 8)
        T["W0t_enc"] = buf["W0t_enc"].view(1, 32, 128, 8)
        T["W1"], T["W1t"] = buf["W1"].view(1, 32, 256, 8), buf["W1t"].view(1, 32, 256, 8)
        T["Wh0"] = buf["Wh0"].view(nh, KH_PAD // 8, 256, 8)
        T["Wh0t_feat"] = buf["Wh0t_feat"].view(1, nh * 32, 256, 8)
        T["Wh0t_x"] = buf["Wh0t_x"].view(1, nh * 32, KH_PAD - XH_OFF, 8)
        T["Whl"] = [buf[f"Whl{l}"].view(nh, 32, 256, 8) for l in range(3)]
        T["Whlt"] = [buf[f"Whlt{l}"].view(nh, 32, 256, 8) for l in range(3)]
        if self.fuse_heads:
            T["Whl128"] = [buf[f"Whl128_{l}"].view(nh * 2, 32, 128, 8) for l in range(3)]
            T["Whlt128"] = [buf[f"Whlt128_{l}"].view(nh * 2, 32, 128, 8) for l in range(3)]  # rows = input unit
        W["Wout"] = self._f(self.J, HID)
        descs = (_lib.WnDesc * 24)()
        specs = self._wn_specs()
        for i, (name, N, K, cmap, row_off, tag) in enumerate(specs):
            d = descs[i]
            d.v, d.g = p[name + ".weight_v"].data_ptr(), p[name + ".weight_g"].data_ptr()
            d.col_map = cmap.data_ptr() if cmap is not None else None
            d.N, d.K, d.row_off, d.tcl_lo = N, K, row_off, -1
            if tag == "sdf0":
                d.tcl, d.tcl_tile, d.tcl_chunks, d.tcl_lo = T["W0s"].data_ptr(), 256, 2 * K0_CH, K0_CH
                d.tclt[0] = T["W0t_enc"].data_ptr()
                d.tclt_c0[0], d.tclt_c1[0], d.tclt_tile[0], d.tclt_chunks[0] = 0, 128, 128, 32
            elif tag == "sdf1":
                d.tcl, d.tcl_tile, d.tcl_chunks = T["W1"].data_ptr(), 256, 32
                d.tclt[0] = T["W1t"].data_ptr()
                d.tclt_c0[0], d.tclt_c1[0], d.tclt_tile[0], d.tclt_chunks[0] = 0, HID, 256, 32
            elif tag[0] == "h0":
                hi = tag[1]
                d.tcl, d.tcl_tile, d.tcl_chunks = T["Wh0"].data_ptr(), 256, KH_PAD // 8
                d.tclt[0] = T["Wh0t_feat"].data_ptr()
                d.tclt_c0[0], d.tclt_c1[0], d.tclt_tile[0], d.tclt_chunks[0] = 0, HID, 256, nh * 32
                d.tclt_col_off[0] = hi * HID
                d.tclt[1] = T["Wh0t_x"].data_ptr()
                d.tclt_c0[1], d.tclt_c1[1], d.tclt_tile[1], d.tclt_chunks[1] = XH_OFF, KH_PAD, KH_PAD - XH_OFF, nh * 32
                d.tclt_col_off[1] = hi * HID
            elif tag[0] == "hl":
                _, l, hi = tag
                d.tcl, d.tcl_tile, d.tcl_chunks = T["Whl"][l].data_ptr(), 256, 32
                d.tclt[0] = T["Whlt"][l].data_ptr()
                d.tclt_c0[0], d.tclt_c1[0], d.tclt_tile[0], d.tclt_chunks[0] = 0, HID, 256, 32
                d.tclt_row_off[0] = hi * HID
                if self.fuse_heads:
                    d.tcl2, d.tcl2_tile, d.tcl2_chunks = T["Whl128"][l].data_ptr(), 128, 32
                    d.tclt[1] = T["Whlt128"][l].data_ptr()
                    d.tclt_c0[1], d.tclt_c1[1], d.tclt_tile[1], d.tclt_chunks[1] = 0, HID, 128, 32
                    d.tclt_row_off[1] = hi * HID
            else:  # output layer: fp32 rows of W_out
                d.Wp, d.ldw = W["Wout"].data_ptr(), HID
        call("mli_weightnorm_pack_batch", _lib.C.addressof(descs), len(specs))
        W["b0"], W["b1"] = p["neural_sdf.mlp.linears.0.bias"], p["neural_sdf.mlp.linears.1.bias"]
        W["w_sdf"], W["b_sdf"] = p["neural_sdf.mlp.linear_sdf.weight"], p["neural_sdf.mlp.linear_sdf.bias"]
        # every head bias of the step in ONE concatenation (layer-major, heads side by side; output-layer biases last)
        nb = nh * HID
        bflat = torch.cat([p[f"neural_rgb.{h[0]}.linears.{l}.bias"] for l in range(4) for h in self.heads] +
                          [p[f"neural_rgb.{h[0]}.linears.4.bias"] for h in self.heads])
        W["bh"] = [bflat[l * nb:(l + 1) * nb] for l in range(4)]
        W["bout"] = bflat[4 * nb:]
        W["T"], W["_keep"] = T, flat
        self.W = W
        return W

    def _unpack_grads_tc(self, p, dW0, dW1, dWh0, dWh, dWout, need_heads, train_mlp):
        """One launch for the weight_norm backward of every matrix: dW (packed fp32) -> dv, dg."""
        specs = [s for s in self._wn_specs() if (s[5] in ("sdf0", "sdf1") and train_mlp) or
                 (s[5] not in ("sdf0", "sdf1") and need_heads)]
        total = sum(N * K + N for _, N, K, _, _, _ in specs)
        flat = self._f(total)
        descs = (_lib.WnDesc * 24)()
        out, off = {}, 0
        for i, (name, N, K, cmap, row_off, tag) in enumerate(specs):
            d = descs[i]
            dv, dg = flat[off:off + N * K].view(N, K), flat[off + N * K:off + N * K + N].view(N, 1)
            off += N * K + N
            d.v, d.g = p[name + ".weight_v"].data_ptr(), p[name + ".weight_g"].data_ptr()
            d.col_map = cmap.data_ptr() if cmap is not None else None
            d.N, d.K, d.tcl_lo = N, K, -1
            d.dv, d.dg = dv.data_ptr(), dg.data_ptr()
            if tag == "sdf0":
                src, d.ldw, d.row_off = dW0, K0_PAD, 0
            elif tag == "sdf1":
                src, d.ldw, d.row_off = dW1, HID, 0
            elif tag[0] == "h0":
                src, d.ldw, d.row_off = dWh0, KH_PAD, row_off
            elif tag[0] == "hl":
                src, d.ldw, d.row_off = dWh[tag[1]][tag[2]], HID, 0
            else:
                src, d.ldw, d.row_off = dWout, HID, row_off
            d.dWp = src.data_ptr()
            out[name + ".weight_v"], out[name + ".weight_g"] = dv, dg
        call("mli_weightnorm_unpack_grad_batch", _lib.C.addressof(descs), len(specs))
        return out
    # ---- tensor-core (bf16 TCL) building blocks -------------------------------------------------------------------
    def _tc_linear(self, A, a_chunk0, a_bchunks, B, b_belems, K, N, BN, bias, bias_b, aux, aux_chunk0, aux_bchunks, act,
                   out, out_f32, out_chunk0, out_bchunks, ldo, M, batch, epi, mask=None, mask_bchunks=0):
        """mask: relu sign bits [tiles, chunks32, 128] int32 -- written by a relu forward (epi 0), read instead of `aux`
        by a relu data gradient (epi 1)."""
        call("mli_tc_linear", A, A.shape[1], a_chunk0, a_bchunks, B, b_belems, K, N, BN, bias, bias_b, aux,
             aux.shape[1] if aux is not None else 0, aux_chunk0, aux_bchunks, act, out, int(out_f32),
             0 if out_f32 else out.shape[1], out_chunk0, out_bchunks, ldo, 0, 0, M, batch, epi, mask,
             mask.shape[1] if mask is not None else 0, 0, mask_bchunks)

    def _mask(self, rows, cols):
        return torch.empty((rows + 127) // 128, cols // 32, 128, dtype=torch.int32, device=self.device)

    def _tc_wgrad(self, L, l_chunk0, l_b, R, r_chunk0, r_b, M, rows, cols, batch, out, ldo, bstride, transpose=0,
                  db=None, defer=True):
        """out[b] = L[b]^T R[b]; db (optional, [batch*rows]) = column sums of L from the same pass (bias gradient).
        defer=False: the caller reads `out` right away (no batching of the split-K reduction)."""
        ws = torch.empty(_lib.load().mli_tc_wgrad_ws_bytes(M, rows, cols, batch), dtype=torch.uint8, device=self.device)
        if defer and self._wg_jobs is not None:  # inside backward(): the split-K reductions of all GEMMs run as one launch later
            job = _lib.TnReduceJob()
            call("mli_tc_wgrad_defer", L, L.shape[1], l_chunk0, l_b, R, R.shape[1], r_chunk0, r_b, M, rows, cols, batch, out,
                 ldo, bstride, transpose, db, rows, ws, job)
            self._wg_jobs.append((job, ws, out, db))  # the partials (and outputs) must outlive the batched reduction
            return
        call("mli_tc_wgrad", L, L.shape[1], l_chunk0, l_b, R, R.shape[1], r_chunk0, r_b, M, rows, cols, batch, out, ldo,
             bstride, transpose, db, rows, ws)

    def _tc_wgrad_flush(self):
        """One launch for the deferred split-K reductions (weight + bias gradients) of every weight-gradient GEMM issued
        since backward() started; must run on the stream those GEMMs were launched on, before anything reads dW / db."""
        jobs = self._wg_jobs
        if not jobs:
            return
        arr = (_lib.TnReduceJob * len(jobs))(*[j[0] for j in jobs])
        call("mli_tc_wgrad_reduce_batch", _lib.C.addressof(arr), len(jobs))
        self._wg_keep.append(list(jobs))  # released at the stream join
        jobs.clear()

    def _tc_colsum(self, X, chunk0, n_chunks, M, out=None):
        out = self._f(n_chunks * 8) if out is None else out
        ws = torch.empty(_lib.load().mli_tc_colsum_ws_bytes(M, n_chunks), dtype=torch.uint8, device=self.device)
        call("mli_tc_colsum", X, X.shape[1], chunk0, n_chunks, M, out, ws)
        return out

    # ------------------------------------------------------------------------------------------------------
    def bounds(self, center, ray_unit):
        R = center.shape[0]
        near, far = self._f(R), self._f(R)
        outside = torch.empty(R, dtype=torch.uint8, device=self.device)
        aabb = list(self.cfg.aabb) if self.cfg.bounding == "box" else None
        call("mli_dist_bounds", center, ray_unit, R, aabb, near, far, outside)
        return near, far, outside

    def sdf_query(self, table, center, ray_unit, dists, ld, n):
        """SDF-only network query at n samples per ray (NeuralSDF.sdf, modules.py:73-74)."""
        R, W = center.shape[0], self.W
        if self.tc:  # encode -> split-bf16 TCL -> tcgen05 trunk with the SDF head fused into the epilogue
            X = self._tcl(R * n, 2 * K0_CH)
            call("mli_encode_rays_tcl", self.grid, table, center, ray_unit, dists, ld, R, n, 0, 0.0,
                 self.cfg.vol_range[0], self.cfg.vol_range[1], X, 2 * K0_CH, K0_CH)
            sdf = self._f(R * n)
            call("mli_tc_sdf_trunk_fwd", X, 2 * K0_CH, K0_PAD, W["T"]["W0s"], W["b0"], W["w_sdf"], W["b_sdf"], R * n, 2,
                 0, None, None, sdf)
            return sdf
        X = self._f(R * n, K0_PAD)
        call("mli_encode_rays", self.grid, table, center, ray_unit, dists, ld, R, n, 0, 0.0, self.cfg.vol_range[0],
             self.cfg.vol_range[1], X, K0_PAD)
        H = self._f(R * n, HID)
        call("mli_linear_fwd", X, K0_PAD, 0, W["W0"], K0_PAD, 0, W["b0"], 0, H, HID, 0, R * n, HID, K0_PAD,
             ACT_SOFTPLUS100, 1, _lib.PREC_FP32)
        sdf = self._f(R * n)
        call("mli_rowdot_fwd", H, HID, R * n, W["w_sdf"], W["b_sdf"], [0], 1, HID, ACT_NONE, 0, sdf, 1)
        return sdf

    def dense_table_bytes(self):
        """Bytes of the leading, densely indexed levels of the table (the part worth keeping in L2)."""
        lv, n = self.grid.level, self.cfg.n_levels
        first_hashed = next((l for l in range(n) if lv[l].hashed), n)
        end = int(lv[first_hashed].offset) if first_hashed < n else int(self.grid.n_entries)
        return end * self.cfg.feat_per_level * 4

    def apply_l2_window(self, table):
        """(Re)install the access-policy window when the table pointer / stream changes; no-op when MLI_L2_PERSIST is 0."""
        if self.l2_persist <= 0.0 or torch.cuda.is_current_stream_capturing():
            return  # (stream attributes cannot change during capture: Model._graphed_train_step installs the window before)
        key = (table.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if key != self._l2_key:
            call("mli_set_l2_window", table, self.dense_table_bytes(), self.l2_persist)
            self._l2_key = key

    def sample(self, table, center, ray_unit, near, far, rands=None):
        """Model.sample_dists_all (neuralangelo/model.py:449-465): coarse + hierarchical importance sampling."""
        cfg, R, N = self.cfg, center.shape[0], self.cfg.n_samples
        self.apply_l2_window(table)
        dists, sdfs = self._f(R, N), self._f(R, N)
        call("mli_sample_coarse", near, far, rands, R, cfg.coarse, dists, N)
        n = cfg.coarse
        if cfg.hierarchy > 0:
            s = self.sdf_query(table, center, ray_unit, dists, N, n)
            sdfs[:, :n] = s.view(R, n)
        # round h: importance-sample n_fine distances from (dists, sdfs), query their SDFs, merge.  The merge of round h
        # and the sampling of round h + 1 run as ONE launch (mli_sample_merge_fine); the last merge needs no SDFs.
        fine, fine_next = self._f(R, cfg.fine), self._f(R, cfg.fine)
        for h in range(cfg.hierarchy):
            if h == 0:
                call("mli_sample_fine", dists, sdfs, N, R, n, cfg.fine, float(64 * 2 ** h), fine, None, None, None, None)
            if h != cfg.hierarchy - 1:
                sf = self.sdf_query(table, center, ray_unit, fine, cfg.fine, cfg.fine)
                call("mli_sample_merge_fine", dists, sdfs, N, R, n, fine, sf, cfg.fine, float(64 * 2 ** (h + 1)), fine_next)
                fine, fine_next = fine_next, fine
            else:
                call("mli_sample_merge", dists, None, N, R, n, fine, None, cfg.fine)
            n += cfg.fine
        return dists

    # ------------------------------------------------------------------------------------------------------
    def sphere_trace(self, table, center, ray_unit, near, far, dist_start=None, iters=20):
        """Model.sphere_tracing_intersection (neuralangelo/model.py:298-325): [R] tensors -> (dist, mask uint8)."""
        R = center.shape[0]
        dist = (dist_start if dist_start is not None else near).clone().contiguous()
        mask = torch.ones(R, dtype=torch.uint8, device=self.device)
        for it in range(iters):
            sdf = self.sdf_query(table, center, ray_unit, dist.view(R, 1), 1, 1)
            call("mli_sphere_trace_step", dist, mask, sdf, near, far, R, int(it == iters - 1))
        return dist, mask

    def light_visibility(self, table, center, ray_unit, pts_light, near, far, blend_dist, gradient, camera_ray_type,
                         radius, aabb=None):
        """Model.get_light_visibility with type 'sphere_tracing' (NeuralLumen/model.py:133-184), eval mode.
        -> visibility uint8 [R], normal_x_light [R], inter_dist [R], inter_mask uint8 [R]"""
        R = center.shape[0]
        if camera_ray_type == "blend_z_sphere_tracing":
            inter_dist, inter_mask = self.sphere_trace(table, center, ray_unit, near, far, dist_start=blend_dist)
        elif camera_ray_type == "blend_z":
            inter_dist, inter_mask = blend_dist.contiguous(), (blend_dist > 0.0).to(torch.uint8)
        elif camera_ray_type == "sphere_tracing":
            inter_dist, inter_mask = self.sphere_trace(table, center, ray_unit, near, far)
        else:
            raise NotImplementedError(f"camera_ray_type {camera_ray_type}")
        light_unit, near_l, far_tr = self._f(R, 3), self._f(R), self._f(R)
        inside = torch.empty(R, dtype=torch.uint8, device=self.device)
        call("mli_light_rays", center, ray_unit, inter_dist, pts_light, R, float(radius), aabb, light_unit, near_l, far_tr,
             inside)
        _, mask_l = self.sphere_trace(table, pts_light, light_unit, near_l, far_tr)
        vis, nxl = torch.empty(R, dtype=torch.uint8, device=self.device), self._f(R)
        call("mli_light_finish", mask_l, inside, gradient.contiguous(), light_unit, R, vis, nxl)
        return vis, nxl, inter_dist, inter_mask

    # ------------------------------------------------------------------------------------------------------
    def forward(self, p, center, ray_unit, pts_light, dists, near, far, outside, training, progress, keep_dz=True):
        """render_rays_object_lumen + compositing for R rays with given sample distances.  Returns (out, ctx)."""
        cfg, W, nh = self.cfg, self.W, self.nh
        R, N = center.shape[0], cfg.n_samples
        M, P = R * N, 1 + cfg.taps
        table = p["neural_sdf.tcnn_encoding.params"]
        prec = _lib.PREC_FP32  # the CUDA-core entry points; tensor-core layers go through _tc_*
        sdf = self._f(P * M)
        X0 = H0 = Xd = S0 = DZ = H0c = None
        if not self.tc:
            X0 = self._f(P * M, K0_PAD)
            call("mli_encode_rays", self.grid, table, center, ray_unit, dists, N, R, N, cfg.taps, self.tap_eps,
                 cfg.vol_range[0], cfg.vol_range[1], X0, K0_PAD)
            H0 = self._f(P * M, HID)
            call("mli_linear_fwd", X0, K0_PAD, 0, W["W0"], K0_PAD, 0, W["b0"], 0, H0, HID, 0, P * M, HID, K0_PAD,
                 ACT_SOFTPLUS100, 1, prec)
            call("mli_rowdot_fwd", H0, HID, P * M, W["w_sdf"], W["b_sdf"], [0], 1, HID, ACT_NONE, 0, sdf, 1)
        else:
            # SDF trunk on the tensor cores in the delta basis: plane 0 = centre rows, planes 1.. = tap - centre
            # (formed in fp32 by the encode kernel); the tap planes' hidden activations never reach HBM
            T = W["T"]
            Xd = self._tcl(P * M, 2 * K0_CH)
            call("mli_encode_rays_tcl", self.grid, table, center, ray_unit, dists, N, R, N, cfg.taps, self.tap_eps,
                 cfg.vol_range[0], cfg.vol_range[1], Xd, 2 * K0_CH, K0_CH)
            # sigma0 / dz are only written for the backward pass; inside the kernel sigma0 lives in TMEM
            S0 = torch.empty(M // 128, 64, 128, 4, dtype=torch.float32, device=self.device) if keep_dz else None
            H0c = self._tcl(M, 32)
            DZ = self._tcl(cfg.taps * M, 32) if keep_dz else None
            call("mli_tc_sdf_trunk_fused", Xd, 2 * K0_CH, K0_PAD, T["W0s"], W["b0"], W["w_sdf"], W["b_sdf"], M, cfg.taps,
                 S0, H0c, DZ, sdf)
        gradients = self._f(M, 3)
        hessians = self._f(M, 3) if training else None
        S = self._f(M, self.lds)
        if not self.tc:
            XH = self._f(M, KH_PAD)
            call("mli_linear_fwd", H0, HID, 0, W["W1"], HID, 0, W["b1"], 0, XH, KH_PAD, 0, M, HID, HID, ACT_SOFTPLUS100, 1,
                 prec)
            call("mli_geometry_fwd", sdf, M, N, cfg.taps, self.tap_eps, outside, cfg.outside_val, center, ray_unit,
                 pts_light, dists, N, gradients, hessians, XH, KH_PAD, XH_OFF, 0, None, 0, 0)
            A = [self._f(M, nh * HID) for _ in range(4)]
            call("mli_linear_fwd", XH, KH_PAD, 0, W["Wh0"], KH_PAD, 0, W["bh"][0], 0, A[0], nh * HID, 0, M, nh * HID,
                 KH_PAD, ACT_RELU, 1, prec)
            for l in range(3):
                call("mli_linear_fwd", A[l], nh * HID, HID, W["Whl"][l], HID, HID * HID, W["bh"][l + 1], HID, A[l + 1],
                     nh * HID, HID, M, HID, HID, ACT_RELU, nh, prec)
            call("mli_rowdot_fwd", A[3], nh * HID, M, W["Wout"], W["bout"], self.col_off, self.J, HID, ACT_SIGMOID,
                 self.act_mask, S, self.lds)
            H0c = None
        else:
            # layer 1 + all head layers on the tensor cores; activations in bf16 TCL (never leave that layout)
            XH = self._tcl(M, KH_PAD // 8)
            self._tc_linear(H0c, 0, 0, T["W1"], 0, HID, HID, 256, W["b1"], 0, None, 0, 0, ACT_SOFTPLUS100, XH, False, 0, 0,
                            0, M, 1, 0)
            # gradients / Hessians + the non-feature head inputs, written straight into XH's last 6 TCL chunks
            call("mli_geometry_fwd", sdf, M, N, cfg.taps, self.tap_eps, outside, cfg.outside_val, center, ray_unit,
                 pts_light, dists, N, gradients, hessians, None, 0, 0, 1, XH, KH_PAD // 8, XH_OFF // 8)
            j0s, njs, j = [], [], 0
            for h in self.heads:
                j0s.append(j)
                njs.append(h[2])
                j += h[2]
            fused = self.fuse_heads and "Whl128" in T and (not keep_dz or self.fuse_heads_train_fwd)
            # hidden activations: kept (bf16 TCL) only when a backward pass follows -- the fused kernel needs no HBM copy
            A = [self._tcl(M, nh * 32) if (keep_dz or not fused) else None for _ in range(4)]
            # relu sign bits of every hidden activation (only when a backward pass follows): the data-gradient
            # epilogues read these 4 bytes per 32 columns instead of the 64-byte bf16 activations
            Am = [self._mask(M, nh * HID) if keep_dz else None for _ in range(4)]
        if self.tc and fused:
            # the whole head stack (layer 0 .. output layers) in ONE persistent kernel: activations stay in shared memory
            call("mli_tc_heads_fwd", XH, M, nh, KH_PAD, int(keep_dz), KH_PAD // 8, T["Wh0"], T["Whl128"][0],
                 T["Whl128"][1], T["Whl128"][2], W["bh"][0], W["bh"][1], W["bh"][2], W["bh"][3], W["Wout"], W["bout"], j0s,
                 njs, ACT_SIGMOID, self.act_mask, A[0], A[1], A[2], A[3], Am[0], Am[1], Am[2], Am[3], S, self.lds)
        elif self.tc:
            self._tc_linear(XH, 0, 0, T["Wh0"], 0, KH_PAD, nh * HID, 256, W["bh"][0], 0, None, 0, 0, ACT_RELU, A[0], False,
                            0, 0, 0, M, 1, 0, mask=Am[0])
            for l in range(2):
                self._tc_linear(A[l], 0, 32, T["Whl"][l], HID * HID, HID, HID, 256, W["bh"][l + 1], HID, None, 0, 0,
                                ACT_RELU, A[l + 1], False, 0, 32, 0, M, nh, 0, mask=Am[l + 1], mask_bchunks=8)
            # last hidden layer with the 256 -> 3/3/1 output layers + sigmoid fused into its epilogue
            call("mli_tc_linear_dot", A[2], nh * 32, 0, 32, T["Whl"][2], HID * HID, HID, W["bh"][3], HID, A[3], nh * 32, 0,
                 32, M, nh, W["Wout"], W["bout"], j0s, njs, ACT_SIGMOID, self.act_mask, S, self.lds, Am[3],
                 Am[3].shape[1] if Am[3] is not None else 0, 0, 8)
        ccfg = _lib.CompositeCfg(N, self.mode, int(cfg.white_background), int(not training),
                                 min(progress / cfg.anneal_end, 1.0),
                                 self._dyn.data_ptr() if self.dynamic_scalars else None)
        weights, out = self._f(R, N), self._f(R, self.n_out)
        extras = self._f(R, 5) if not training else None
        call("mli_composite_fwd", ccfg, p["s_var"], sdf, gradients, ray_unit, dists, N, far, S, self.lds, R, None, weights, out,
             extras)
        ctx = dict(R=R, M=M, P=P, X0=X0, H0=H0, H0c=H0c, Xd=Xd, S0=S0, DZ=DZ, Am=Am if self.tc else None, sdf=sdf, XH=XH, gradients=gradients, A=A, S=S, weights=weights,
                   ccfg=ccfg, center=center, ray_unit=ray_unit, dists=dists, far=far, outside=outside)
        res = dict(out=out, weights=weights, gradients=gradients, hessians=hessians, extras=extras, S=S, sdf=sdf)
        return res, ctx

    # ------------------------------------------------------------------------------------------------------
    def backward(self, p, ctx, d_out, d_gradients=None, d_hessians=None, d_weights=None, need=("heads", "sdf", "table",
                                                                                               "s_var")):
        """Hand-written backward of forward().  Returns a dict {state-dict name: gradient}."""
        cfg, W, nh = self.cfg, self.W, self.nh
        R, M, P, N = ctx["R"], ctx["M"], ctx["P"], cfg.n_samples
        prec = _lib.PREC_FP32  # the CUDA-core entry points; tensor-core layers go through _tc_*
        need_sdf = ("sdf" in need) or ("table" in need)
        grads = {}
        self._wg_jobs = [] if self.tc else None  # deferred split-K reductions of this pass (see _tc_wgrad_flush)
        d_grad = d_gradients.contiguous().clone() if d_gradients is not None else self._z(M, 3)
        dS, d_sdf_c = self._f(M, self.lds), self._f(M)
        d_svar = self._z(1) if "s_var" in need else None
        call("mli_composite_bwd", ctx["ccfg"], p["s_var"], ctx["sdf"], ctx["gradients"], ctx["ray_unit"], ctx["dists"], N,
             ctx["far"], ctx["S"], self.lds, R, ctx["weights"], d_out, d_weights, self.act_mask, dS, d_sdf_c, d_grad, d_svar, 0,
             self._f(R) if d_svar is not None else None)
        if d_svar is not None:
            grads["s_var"] = d_svar.view(())
        A = ctx["A"]
        need_heads = "heads" in need
        if not (need_heads or need_sdf):
            self._wg_jobs = None
            return grads
        train_mlp = "sdf" in need
        H0 = ctx["H0"]
        dWh, dbh = [None] * 3, [None] * 4
        dWout, dbout = self._f(self.J, HID), self._f(self.J)
        dWh0 = self._f(nh * HID, KH_PAD) if need_heads else None
        dZ0 = self._f(P * M, HID) if (need_sdf and not self.tc) else None  # rows [0, M) first receive dL/dH0 (centre)
        dH0 = self._tcl(M, 32) if (need_sdf and self.tc) else None            # same quantity, bf16 TCL
        dXx = self._f(M, KH_PAD - XH_OFF) if need_sdf else None
        dW1, db1 = (self._f(HID, HID), None) if (need_sdf and train_mlp) else (None, None)
        later = []  # deferred weight-gradient launches (tensor-core mode)
        if not self.tc:
            # ---- heads, fp32 CUDA-core path -------------------------------------------------------------------------
            dZ = self._f(M, nh * HID)
            ws = torch.empty(_lib.load().mli_rowdot_bwd_ws_bytes(M, self.J, HID), dtype=torch.uint8, device=self.device)
            call("mli_rowdot_bwd", dS, self.lds, A[3], nh * HID, M, W["Wout"], self.col_off, self.J, HID, ACT_RELU, dZ, nh * HID,
                 nh * HID, 0, dWout if need_heads else None, dbout, ws)
            for l in (2, 1, 0):
                if need_heads:
                    dWh[l], dbh[l + 1] = self._f(nh, HID, HID), self._f(nh * HID)
                    wsb = _lib.load().mli_linear_wgrad_ws_bytes(M, HID, HID, nh)
                    call("mli_linear_wgrad", dZ, nh * HID, HID, A[l], nh * HID, HID, dWh[l], HID, HID * HID, dbh[l + 1],
                         HID, M, HID, HID, nh, prec, torch.empty(wsb, dtype=torch.uint8, device=self.device))
                dZp = self._f(M, nh * HID)
                call("mli_linear_dgrad", dZ, nh * HID, HID, W["Whlt"][l], HID, HID * HID, A[l], nh * HID, HID, dZp,
                     nh * HID, HID, M, HID, HID, ACT_RELU, 0, nh, prec)
                dZ = dZp
            XH = ctx["XH"]
            if need_heads:
                dbh[0] = self._f(nh * HID)
                wsb = _lib.load().mli_linear_wgrad_ws_bytes(M, nh * HID, KH_PAD, 1)
                call("mli_linear_wgrad", dZ, nh * HID, 0, XH, KH_PAD, 0, dWh0, KH_PAD, 0, dbh[0], 0, M, nh * HID, KH_PAD,
                     1, prec, torch.empty(wsb, dtype=torch.uint8, device=self.device))
            if need_sdf:
                dZ1 = self._f(M, HID)  # d pre-activation of SDF layer 1 = (dZ_head0 . Wh0[:, feat]) * softplus'(feat)
                call("mli_linear_dgrad", dZ, nh * HID, 0, W["Wh0t"], nh * HID, 0, XH, KH_PAD, 0, dZ1, HID, 0, M, nh * HID,
                     HID, ACT_SOFTPLUS100, 0, 1, prec)
                call("mli_linear_dgrad", dZ, nh * HID, 0, W["Wh0t"][XH_OFF:], nh * HID, 0, None, 0, 0, dXx, KH_PAD - XH_OFF,
                     0, M, nh * HID, KH_PAD - XH_OFF, ACT_NONE, 0, 1, prec)
                if train_mlp:
                    db1 = self._f(HID)
                    wsb = _lib.load().mli_linear_wgrad_ws_bytes(M, HID, HID, 1)
                    call("mli_linear_wgrad", dZ1, HID, 0, H0, HID, 0, dW1, HID, 0, db1, 0, M, HID, HID, 1, prec,
                         torch.empty(wsb, dtype=torch.uint8, device=self.device))
                call("mli_linear_dgrad", dZ1, HID, 0, W["W1t"], HID, 0, None, 0, 0, dZ0, HID, 0, M, HID, HID, ACT_NONE, 0,
                     1, prec)
        else:
            # ---- heads + SDF layer 1 on the tensor cores (bf16 TCL operands, fp32 accumulation in TMEM) ---------------
            # Data-gradient chain first, weight gradients deferred (`later`): the hash-table gradient -- the 1.46 GB that
            # a multi-GPU step has to all-reduce -- only depends on the former, so it is launched as early as possible
            # and its exchange overlaps the weight-gradient GEMMs.
            T, XH = W["T"], ctx["XH"]
            Am = ctx["Am"] if ctx["Am"][0] is not None else [None] * 4
            fused_bwd = self.fuse_heads and "Whlt128" in T and Am[0] is not None
            if fused_bwd:
                # the whole data-gradient chain (output layers + three hidden layers) in ONE on-chip kernel: every
                # pre-activation gradient is written once, for the weight-gradient GEMMs and the layer-0 data gradient
                dZs = [self._tcl(M, nh * 32) for _ in range(4)]
                j0s, njs, j = [], [], 0
                for h in self.heads:
                    j0s.append(j)
                    njs.append(h[2])
                    j += h[2]
                call("mli_tc_heads_bwd", dS, self.lds, M, nh, T["Whlt128"][2], T["Whlt128"][1], T["Whlt128"][0], W["Wout"],
                     j0s, njs, Am[0], Am[1], Am[2], Am[3], dZs[0], dZs[1], dZs[2], dZs[3])
                dZ = dZs[3]
            else:
                dZ = self._tcl(M, nh * 32)
                call("mli_tc_rowdot_bwd_data", dS, self.lds, A[3], nh * 32, M, W["Wout"], self.col_off, self.J, HID,
                     ACT_RELU, dZ, Am[3])
            if need_heads:
                def _out_layer():
                    dSt = self._to_tcl(dS, self.lds, M, self.lds, self._tcl(M, 2), 128, 0, 2)
                    outT = self._f(nh, 16, HID)  # [head][j][k] = sum_m dS[m, j] A4[m, head*256 + k]
                    self._tc_wgrad(A[3], 0, 32, dSt, 0, 0, M, HID, 16, nh, outT, HID, 16 * HID, transpose=1, defer=False)
                    dWout.copy_(torch.stack([outT[self.col_off[j] // HID, j] for j in range(self.J)]))
                    dbout.copy_(self._tc_colsum(dSt, 0, 2, M)[:self.J])
                later.append(_out_layer)
            for l in (2, 1, 0):
                if need_heads:
                    dWh[l], dbh[l + 1] = self._f(nh, HID, HID), self._f(nh * HID)
                    later.append(lambda dZ=dZ, l=l: self._tc_wgrad(dZ, 0, 32, A[l], 0, 32, M, HID, HID, nh, dWh[l], HID,
                                                                   HID * HID, db=dbh[l + 1]))
                if fused_bwd:
                    dZ = dZs[l]
                    continue
                dZp = self._tcl(M, nh * 32)
                self._tc_linear(dZ, 0, 32, T["Whlt"][l], HID * HID, HID, HID, 256, None, 0,
                                A[l] if Am[l] is None else None, 0, 32, ACT_RELU, dZp, False, 0, 32, 0, M, nh, 1,
                                mask=Am[l], mask_bchunks=8)
                dZ = dZp
            if need_heads:
                dbh[0] = self._f(nh * HID)
                later.append(lambda dZ=dZ: self._tc_wgrad(dZ, 0, 0, XH, 0, 0, M, nh * HID, 256, 1, dWh0, KH_PAD, 0, db=dbh[0]))
                later.append(lambda dZ=dZ: self._tc_wgrad(dZ, 0, 0, XH, XH_OFF // 8, 0, M, nh * HID, KH_PAD - XH_OFF, 1,
                                                          dWh0[:, XH_OFF:], KH_PAD, 0))
            if need_sdf:
                dZ1 = self._tcl(M, 32)
                self._tc_linear(dZ, 0, 0, T["Wh0t_feat"], 0, nh * HID, HID, 256, None, 0, XH, 0, 0, ACT_SOFTPLUS100, dZ1,
                                False, 0, 0, 0, M, 1, 1)
                self._tc_linear(dZ, 0, 0, T["Wh0t_x"], 0, nh * HID, KH_PAD - XH_OFF, KH_PAD - XH_OFF, None, 0, None, 0, 0,
                                ACT_NONE, dXx, True, 0, 0, KH_PAD - XH_OFF, M, 1, 1)
                if train_mlp:
                    db1 = self._f(HID)
                    later.append(lambda: self._tc_wgrad(dZ1, 0, 0, ctx["H0c"], 0, 0, M, HID, HID, 1, dW1, HID, 0, db=db1))
                self._tc_linear(dZ1, 0, 0, T["W1t"], 0, HID, HID, 256, None, 0, None, 0, 0, ACT_NONE, dH0, False, 0, 0, 0,
                                M, 1, 1)
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record()  # every operand of the head / layer-1 weight gradients is final here
        if need_sdf:
            grads.update(self._backward_sdf(p, ctx, need, train_mlp, d_grad, d_hessians, dXx, d_sdf_c, dZ0, dH0, dW1, db1,
                                            later, ev))
        self._flush_later(later, ev)  # whatever has not been started yet (heads-only training, fp32 mode)
        if self.tc:
            # weight_norm backward of every matrix in one launch; biases are plain column sums (already computed)
            dW0 = grads.pop("_dW0", None)
            if need_heads or dW0 is not None:
                unp = {}

                def _reduce_and_unpack():
                    self._tc_wgrad_flush()  # dW / db of every layer: one launch for the 13 split-K reductions
                    unp.update(self._unpack_grads_tc(p, dW0, dW1, dWh0, dWh, dWout, need_heads, dW0 is not None))
                self._flush_later([_reduce_and_unpack])
                grads.update(unp)
            else:
                self._flush_later([self._tc_wgrad_flush])
            self._wg_jobs = None
            if self.overlap_wgrad and self._wg_stream is not None:
                cur.wait_stream(self._wg_stream)  # join: all gradients are complete for whoever reads them next
            self._wg_keep.clear()
            if need_heads:
                j0 = 0
                for hi, (name, kind, odim, _) in enumerate(self.heads):
                    pre = f"neural_rgb.{name}.linears."
                    for l in range(4):
                        grads[pre + f"{l}.bias"] = dbh[l][hi * HID:(hi + 1) * HID]
                    grads[pre + "4.bias"] = dbout[j0:j0 + odim]
                    j0 += odim
        elif need_heads:
            j0 = 0
            for hi, (name, kind, odim, _) in enumerate(self.heads):
                pre = f"neural_rgb.{name}.linears."
                k_in = len(_col_map(kind))
                dv, dg = self._f(HID, k_in), self._f(HID, 1)
                call("mli_weightnorm_unpack_grad", p[pre + "0.weight_v"], p[pre + "0.weight_g"], dWh0, KH_PAD, HID, k_in,
                     self.map_head[hi], hi * HID, dv, dg)
                grads[pre + "0.weight_v"], grads[pre + "0.weight_g"] = dv, dg
                grads[pre + "0.bias"] = dbh[0][hi * HID:(hi + 1) * HID]
                for l in range(3):
                    dv, dg = self._f(HID, HID), self._f(HID, 1)
                    call("mli_weightnorm_unpack_grad", p[pre + f"{l + 1}.weight_v"], p[pre + f"{l + 1}.weight_g"],
                         dWh[l][hi], HID, HID, HID, None, 0, dv, dg)
                    grads[pre + f"{l + 1}.weight_v"], grads[pre + f"{l + 1}.weight_g"] = dv, dg
                    grads[pre + f"{l + 1}.bias"] = dbh[l + 1][hi * HID:(hi + 1) * HID]
                dv, dg = self._f(odim, HID), self._f(odim, 1)
                call("mli_weightnorm_unpack_grad", p[pre + "4.weight_v"], p[pre + "4.weight_g"], dWout, HID, odim, HID,
                     None, j0, dv, dg)
                grads[pre + "4.weight_v"], grads[pre + "4.weight_g"] = dv, dg
                grads[pre + "4.bias"] = dbout[j0:j0 + odim]
                j0 += odim
        return grads

    def _backward_sdf(self, p, ctx, need, train_mlp, d_grad, d_hessians, dXx, d_sdf_c, dZ0, dH0, dW1, db1, later,
                      ev_heads=None):
        """SDF stencil -> SDF trunk -> hash table part of backward().  The table gradient is launched here, level group
        by level group; the trunk's weight gradients are appended to `later`."""
        cfg, W = self.cfg, self.W
        R, M, P, N = ctx["R"], ctx["M"], ctx["P"], cfg.n_samples
        prec = _lib.PREC_FP32
        H0 = ctx["H0"]
        grads = {}
        d_sdf = self._f(P * M)
        call("mli_geometry_bwd", ctx["gradients"], M, N, cfg.taps, self.tap_eps, ctx["outside"], d_grad, d_hessians, dXx,
             KH_PAD - XH_OFF, 0, d_sdf_c, d_sdf)
        dw_sdf, db_sdf = (self._f(1, HID), self._f(1)) if train_mlp else (None, None)
        dW0, db0 = (self._f(HID, K0_PAD), None) if train_mlp else (None, None)
        if self.tc:
            # ---- SDF trunk backward on the tensor cores, delta basis (csrc/sdf_trunk.cu) -------------------------------
            T = W["T"]
            if ctx["DZ"] is None:
                raise _lib.MliError("forward() was run with keep_dz=False: the SDF trunk cannot be differentiated")
            Ed = self._tcl(P * M, 32)
            ws = torch.empty(_lib.load().mli_tc_sdf_trunk_bwd_ws_bytes(M), dtype=torch.uint8, device=self.device)
            call("mli_tc_sdf_trunk_bwd", d_sdf, M, cfg.taps, ctx["S0"], ctx["DZ"], dH0, ctx["H0c"], W["w_sdf"], Ed, dw_sdf,
                 db_sdf, ws)
            hold = self.wgrad_after_scatter and "table" in need
            if not hold:
                self._flush_later(later, ev_heads)  # head / layer-1 weight gradients: start them on the second stream now
            if train_mlp:
                db0 = self._f(HID)
                ev_ed = torch.cuda.Event()
                ev_ed.record()  # E_d is final
                w0 = [lambda: self._tc_wgrad(Ed, 0, 0, ctx["Xd"], 0, 0, P * M, HID, K0_PAD, 1, dW0, K0_PAD, 0),
                      lambda: self._tc_colsum(Ed, 0, 32, M, out=db0)]  # plane 0 = E = sum over the planes
                if hold:
                    later.extend(w0)
                else:
                    self._flush_later(w0, ev_ed)
            if "table" in need:
                dX0 = self._tcl(P * M, 16)  # bf16 TCL: chunk l = level l, read back coalesced by the scatter kernel
                self._tc_linear(Ed, 0, 0, T["W0t_enc"], 0, HID, 128, 128, None, 0, None, 0, 0, ACT_NONE, dX0, False, 0, 0,
                                0, P * M, 1, 1)
        else:
            # ---- SDF network layer 0 + SDF head in fp32 on the CUDA cores (the rtol-1e-3 parity mode) -------------------
            ws = torch.empty(_lib.load().mli_rowdot_bwd_ws_bytes(P * M, 1, HID), dtype=torch.uint8, device=self.device)
            # centre plane: accumulate onto the layer-1 path; tap planes: SDF head only
            call("mli_rowdot_bwd", d_sdf, 1, H0, HID, M, W["w_sdf"], [0], 1, HID, ACT_SOFTPLUS100, dZ0, HID, HID, 1, None,
                 None, ws)
            call("mli_rowdot_bwd", d_sdf[M:], 1, H0[M:], HID, (P - 1) * M, W["w_sdf"], [0], 1, HID, ACT_SOFTPLUS100,
                 dZ0[M:], HID, HID, 0, None, None, ws)
            if train_mlp:
                call("mli_rowdot_bwd", d_sdf, 1, H0, HID, P * M, W["w_sdf"], [0], 1, HID, ACT_NONE, None, 0, 0, 0, dw_sdf,
                     db_sdf, ws)
                db0 = self._f(HID)
                wsb = _lib.load().mli_linear_wgrad_ws_bytes(P * M, HID, K0_PAD, 1)
                call("mli_linear_wgrad", dZ0, HID, 0, ctx["X0"], K0_PAD, 0, dW0, K0_PAD, 0, db0, 0, P * M, HID, K0_PAD, 1,
                     prec, torch.empty(wsb, dtype=torch.uint8, device=self.device))
            if "table" in need:
                dX0 = self._f(P * M, 128)
                call("mli_linear_dgrad", dZ0, HID, 0, W["W0t"], HID, 0, None, 0, 0, dX0, 128, 0, P * M, HID, 128, ACT_NONE,
                     0, 1, prec)
        if "table" in need:
            tg = self._take_table_grad()
            if self.tc:
                groups = self.level_groups() if self.table_grad_hook is not None else \
                    [(0, cfg.n_levels, 0, self.n_table_params())]  # one launch when nobody waits for single slabs
                for lv0, lv1, e0, e1 in groups:
                    call("mli_encode_rays_bwd_tcl", self.grid, ctx["center"], ctx["ray_unit"], ctx["dists"], N, R, N,
                         cfg.taps, self.tap_eps, cfg.vol_range[0], cfg.vol_range[1], dX0, 16, tg, lv0, lv1)
                    if self.table_grad_hook is not None:
                        self.table_grad_hook(tg, e0, e1)
                if later and self.wgrad_after_scatter:
                    ev_sc = torch.cuda.Event()
                    ev_sc.record()
                    self._flush_later(later, ev_sc)
            else:
                call("mli_encode_rays_bwd", self.grid, ctx["center"], ctx["ray_unit"], ctx["dists"], N, R, N, cfg.taps,
                     self.tap_eps, cfg.vol_range[0], cfg.vol_range[1], dX0, 128, tg, 0)
            grads["neural_sdf.tcnn_encoding.params"] = tg
        if train_mlp and self.tc:  # weight_v / weight_g come from the batched unpack in backward()
            grads["_dW0"] = dW0
            grads["neural_sdf.mlp.linears.0.bias"] = db0
            grads["neural_sdf.mlp.linears.1.bias"] = db1
            grads["neural_sdf.mlp.linear_sdf.weight"], grads["neural_sdf.mlp.linear_sdf.bias"] = dw_sdf, db_sdf
        elif train_mlp:
            dv0, dg0, dv1, dg1 = self._f(HID, 131), self._f(HID, 1), self._f(HID, HID), self._f(HID, 1)

            def _unpack():
                call("mli_weightnorm_unpack_grad", p["neural_sdf.mlp.linears.0.weight_v"],
                     p["neural_sdf.mlp.linears.0.weight_g"], dW0, K0_PAD, HID, 131, self.map_sdf0, 0, dv0, dg0)
                call("mli_weightnorm_unpack_grad", p["neural_sdf.mlp.linears.1.weight_v"],
                     p["neural_sdf.mlp.linears.1.weight_g"], dW1, HID, HID, HID, None, 0, dv1, dg1)
            later.append(_unpack)  # after the (possibly deferred) weight-gradient GEMMs
            grads["neural_sdf.mlp.linears.0.weight_v"], grads["neural_sdf.mlp.linears.0.weight_g"] = dv0, dg0
            grads["neural_sdf.mlp.linears.0.bias"] = db0
            grads["neural_sdf.mlp.linears.1.weight_v"], grads["neural_sdf.mlp.linears.1.weight_g"] = dv1, dg1
            grads["neural_sdf.mlp.linears.1.bias"] = db1
            grads["neural_sdf.mlp.linear_sdf.weight"], grads["neural_sdf.mlp.linear_sdf.bias"] = dw_sdf, db_sdf
        return grads

    # ------------------------------------------------------------------------------------------------------
    def losses(self, lcfg, out, gradients, hessians, outside, targets):
        """In-kernel losses + gradient seeds (NeuralLumen/trainer.py:133-149)."""
        R, N = out.shape[0], self.cfg.n_samples
        M = R * N
        losses = self._f(8)
        d_out, d_grad = self._f(R, self.n_out), self._f(M, 3)
        d_hess = self._f(M, 3) if hessians is not None else None
        ws = torch.empty(_lib.load().mli_losses_ws_bytes(R, M), dtype=torch.uint8, device=self.device)
        if self.dynamic_scalars:  # the kernels read the five weights from the device buffer (see set_dynamic_scalars)
            lcfg = _lib.LossCfg.from_buffer_copy(lcfg)
            lcfg.weights_dev = self._dyn.data_ptr() + 4
        call("mli_losses_fwd_bwd", lcfg, self.mode, out, gradients, hessians, outside, R, N, targets["image_sampled"],
             targets.get("pseudo_ref_sampled"), targets.get("pseudo_sha_sampled"),
             targets.get("pseudo_visibility_certainty_sampled"), losses, d_out, d_grad, d_hess, ws)
        return losses, d_out, d_grad, d_hess
