#!/usr/bin/env python
"""bench.py -- train rays/s (forward + backward) of the NeuralLumen render hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W                 (N>1: launched through torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W   CPU reference arm (oracle port, host cores)

Headline workload (BASELINE.json configs[1], SURVEY.md section 8d "C1"): syn_hotdog_b shape -- Neuralangelo hash grid
(16 levels x 8 features, 2^22 entries/level), 2048 rays x 128 samples per step and GPU, 4 gradient taps, rgb_r_s heads,
all five losses, FULL-GRAD (every parameter incl. the 1.46 GB hash table receives a gradient), synthetic rays from a
pinhole camera at distance 3 (f = 711 px, 512x512), random-init weights (reference init).  One "step" = ray generation
+ bounds + hierarchical sampling + forward + in-kernel losses + full backward (optimizer excluded, as in the metric);
with N>1 ranks it also includes the exchange of all parameter gradients (mli_nerf_b200/dist.py; see `config.exchange`).

One JSON line on rank 0.  Keys beyond the contract's:
  roofline.kernels[]     per C-ABI entry point of the step: algorithmic flops / bytes of the calls (from their shapes),
                         CUDA-event microseconds per step, achieved TFLOP/s and GB/s, fraction of the nearer roof
  render                 rays/s of Model.inference on an 800x800 view (chunks of 20 000), with and without the per-sample
                         debug tensors; at N>1 the frame's rays are partitioned over the ranks (BASELINE configs[4])
  autograd_dropin        rays/s of the true drop-in path model(data) -> torch losses -> backward()
  workloads              the other shapes of SURVEY 8d, each timed for a few steps: NRHints_Pikachu_b (C2),
                         rene_savannah_b (C3: AABB bounds, rays from the reference's own ReNe frames, DistributedSampler
                         frame assignment), the 128+64-sample variant, C0 (T = 2^14, 4096 rays; GPU and CPU)
  with_optimizer         the same step followed by the AdamW update (not the metric)
`--workload NAME` makes one of those shapes the measured one (value / e2e / roofline); `--no-extras` skips the extra
sections.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 128
METRIC = "train_rays_per_sec_fwd_bwd"
UNIT = "rays/s"
# algorithmic MLP flops per ray, fwd+bwd, 4 taps, full-grad (unpadded layer shapes)
MLP_FLOP_PER_RAY = 806.0e6

# name -> (experiment, log2 table size, rays per step and GPU, coarse samples, description)
WORKLOADS = {
    "syn_hotdog_b": ("syn_hotdog_b", 22, 2048, 64,
                     "syn_hotdog_b shape: hash grid 16x8 T=2^22, 2048 rays x 128 samples/GPU, 4 taps, rgb_r_s, 5 losses, "
                     "full-grad"),
    "NRHints_Pikachu_b": ("NRHints_Pikachu_b", 22, 2048, 64,
                          "NRHints_Pikachu_b shape (C2): as syn_hotdog_b, black background, multi-light pseudo-intrinsic "
                          "losses + eikonal/curvature"),
    "rene_savannah_b": ("rene_savannah_b", 22, 2048, 64,
                        "rene_savannah_b shape (C3): 270x360 frames of dataset_rene/savannah/train_transforms.json "
                        "(44 cameras x 37 lights, committed as tests/golden/rene_savannah_frames.npz), AABB bounds, rank r "
                        "takes frame perm[it*W + r] (DistributedSampler), ray_idx = randperm(97200)[:2048]"),
    "syn_hotdog_b_192": ("syn_hotdog_b", 22, 2048, 128,
                         "syn_hotdog_b shape with 128 coarse + 4x16 fine = 192 samples per ray (north star '128+64')"),
    "c0": ("syn_hotdog_b", 14, 4096, 64,
           "C0 (BASELINE configs[0]): 4096 rays x 128 samples, T=2^14 (every level hashed), rgb_r_s, 5 losses, full-grad"),
}


def synthetic_batch(R, seed, H=512, W=512):
    """One frame worth of training data in the reference's `data` dict format."""
    import torch
    g = torch.Generator().manual_seed(seed)
    # camera on a radius-3 sphere looking at the origin (world->camera [R|t]), point light at radius 4
    d = torch.nn.functional.normalize(torch.randn(3, generator=g), dim=0)
    eye = 3.0 * d
    z = -d
    up = torch.tensor([0.0, 0.0, 1.0]) if abs(float(d[2])) < 0.9 else torch.tensor([0.0, 1.0, 0.0])
    x = torch.nn.functional.normalize(torch.linalg.cross(z, up), dim=0)
    y = torch.linalg.cross(z, x)
    Rm = torch.stack([x, y, z])  # rows = camera axes in world coordinates
    pose = torch.cat([Rm, (-Rm @ eye)[:, None]], dim=1)[None]
    lp = 4.0 * torch.nn.functional.normalize(torch.randn(3, generator=g), dim=0)
    pose_light = torch.cat([torch.eye(3), (-lp)[:, None]], dim=1)[None]
    intr = torch.tensor([[[711.0, 0.0, W / 2], [0.0, 711.0, H / 2], [0.0, 0.0, 1.0]]])
    ray_idx = torch.randperm(H * W, generator=g)[:R][None]
    return dict(pose=pose.float(), intr=intr, pose_light=pose_light.float(), ray_idx=ray_idx,
                idx=torch.zeros(1, dtype=torch.long),
                image_sampled=torch.rand(1, R, 3, generator=g), pseudo_ref_sampled=torch.rand(1, R, 3, generator=g),
                pseudo_sha_sampled=torch.rand(1, R, 1, generator=g),
                pseudo_visibility_certainty_sampled=torch.rand(1, R, 1, generator=g))


_RENE = {}


def rene_batch(R, it, rank, world, seed=0):
    """Training sample of the rene_savannah_b workload: the reference's own frames (camera + light poses through its pose
    conventions, oracle/gen_rene_frames.py), DistributedSampler frame assignment, random pixels, random targets."""
    import numpy as np
    import torch
    if not _RENE:
        d = np.load(os.path.join(ROOT, "tests", "golden", "rene_savannah_frames.npz"))
        _RENE.update({k: torch.from_numpy(d[k].astype(np.float32)) if d[k].dtype.kind == "f" else d[k] for k in d.files})
        _RENE["perm"] = torch.randperm(_RENE["pose"].shape[0], generator=torch.Generator().manual_seed(seed))
    n = _RENE["pose"].shape[0]
    f = int(_RENE["perm"][(it * world + rank) % n])
    H, W = (int(v) for v in _RENE["image_size"])
    g = torch.Generator().manual_seed(1_000_003 * seed + 7919 * it + rank)
    return dict(pose=_RENE["pose"][f][None], intr=_RENE["intr"][None], pose_light=_RENE["pose_light"][f][None],
                ray_idx=torch.randperm(H * W, generator=g)[:R][None], idx=torch.tensor([f]),
                image_sampled=torch.rand(1, R, 3, generator=g), pseudo_ref_sampled=torch.rand(1, R, 3, generator=g),
                pseudo_sha_sampled=torch.rand(1, R, 1, generator=g),
                pseudo_visibility_certainty_sampled=torch.rand(1, R, 1, generator=g))


def workload_batch(name, i, rank, world):
    R = WORKLOADS[name][2]
    if name == "rene_savannah_b":
        return rene_batch(R, i, rank, world)
    return synthetic_batch(R, 1000 * rank + i)


def workload_cfg(name, precision):
    from mli_nerf_b200 import config
    exp, dict_size, rays, coarse, _ = WORKLOADS[name]
    cfg = config.experiment(exp, dict_size=dict_size, rand_rays=rays)
    cfg.model.render.num_samples.coarse = coarse
    cfg.model.mli_precision = precision
    return cfg


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of THIS rank's GPU during the timed region (B200_PROFILING.md recipe), read in-process
    through NVML (nvidia_ml_py).  Round 2 lesson: the nvidia-smi subprocess this used to spawn five times a second on
    every rank takes driver-wide locks while it enumerates all GPUs -- at N = 8 forty of them per second stalled every
    rank's kernel launches (host enqueue 9.4 ms per step during the sampled loop, 4 ms without the sampler).  Falls
    back to one nvidia-smi query per second when NVML cannot be imported."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._halt = index, [], set(), threading.Event()
        self.sm_max, self.source = None, "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml, self.source = None, "nvidia-smi"

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        for name, bit in self.REASONS:
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        f = [v.strip() for v in out.split(",")]
        self.samples.append(float(f[0]))
        self.sm_max = float(f[1])
        for (name, _), v in zip(self.REASONS, f[2:]):
            if v.lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self._halt.is_set():
            try:
                self._sample_nvml() if self._nvml is not None else self._sample_smi()
            except Exception:
                pass
            self._halt.wait(0.02 if self._nvml is not None else 1.0)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s), "source": self.source}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# algorithmic work of one C-ABI call, from its shape arguments (include/mli_b200.h): (flops, HBM bytes)
# ------------------------------------------------------------------------------------------------------------------
def kernel_work(name, a):
    f32 = 4
    if name == "mli_tc_linear":
        K, N, M, batch = a[6], a[7], a[24], a[25]
        byt = M * batch * (K * 2 + N * (f32 if a[17] else 2)) + K * N * 2 * batch
        if a[11] is not None:
            byt += M * N * 2 * batch          # previous layer's output (activation derivative)
        if a[27] is not None:
            byt += M * N // 8 * batch         # relu sign bits
        return 2.0 * M * K * N * batch, float(byt)
    if name == "mli_tc_linear_dot":
        K, M, batch = a[6], a[13], a[14]
        return 2.0 * M * K * 256 * batch + 2.0 * M * 256 * 7, float(M * batch * (K + 256) * 2 + M * 8 * f32)
    if name == "mli_tc_heads_fwd":
        M, nh, kh, store = a[1], a[2], a[3], a[4]
        return 2.0 * M * nh * 256 * (kh + 3 * 256) + 2.0 * M * 256 * 7, \
            float(M * kh * 2 + (M * nh * 256 * 2 * 4 if store else 0) + M * 8 * f32)
    if name == "mli_tc_heads_bwd":
        M, nh = a[2], a[3]
        return 2.0 * M * nh * 256 * 3 * 256 + 2.0 * M * 256 * 7, float(M * 8 * f32 + M * nh * 256 // 8 * 4 + M * nh * 256 * 2 * 4)
    if name in ("mli_tc_wgrad", "mli_tc_wgrad_defer"):
        M, rows, cols, batch = a[8], a[9], a[10], a[11]
        return 2.0 * M * rows * cols * batch, float(M * batch * (rows + cols) * 2)
    if name == "mli_tc_sdf_trunk_fused":
        xch, K, M, taps = a[1], a[2], a[7], a[8]
        P = 1 + taps
        byt = M * P * xch * 16 + M * 256 * 2 + P * M * f32
        byt += (M * 256 * f32 if a[9] is not None else 0) + (taps * M * 256 * 2 if a[11] is not None else 0)
        return 2.0 * M * P * K * 256 + 2.0 * M * P * 256, float(byt)
    if name == "mli_tc_sdf_trunk_fwd":
        xch, K, rows = a[1], a[2], a[7]
        return 2.0 * rows * K * 256 + 2.0 * rows * 256, float(rows * xch * 16 + rows * f32)
    if name == "mli_tc_sdf_trunk_bwd":
        M, taps = a[1], a[2]
        P = 1 + taps
        return 8.0 * M * P * 256, float(M * (P * f32 + 256 * f32 + taps * 256 * 2 + 2 * 256 * 2) + P * M * 256 * 2)
    if name == "mli_tc_rowdot_bwd_data":
        M, J = a[4], a[7]
        nh = len(set(int(v) for v in a[6]))
        return 2.0 * M * J * 256, float(M * (8 * f32 + nh * 256 * 2 + (nh * 256 // 8 if a[11] is not None else nh * 256 * 2)))
    if name == "mli_encode_rays_tcl":
        q = a[6] * a[7] * (1 + a[8])
        return 0.0, float(q * (16 * 8 * 32) + q * a[13] * 16)   # one 32-byte table entry per corner and level + X_d row
    if name == "mli_encode_rays_bwd_tcl":
        q, levels = a[5] * a[6] * (1 + a[7]), a[15] - a[14]
        return 0.0, float(q * levels * (8 * 32 * 2) + q * levels * 16)  # fp32 read-modify-write per corner + dX slice
    if name in ("mli_encode_rays", "mli_encode_rays_bwd"):
        q = a[6] * a[7] * (1 + a[8]) if name == "mli_encode_rays" else a[5] * a[6] * (1 + a[7])
        return 0.0, float(q * 16 * 8 * 32 * (1 if name == "mli_encode_rays" else 2) + q * 128 * f32)
    if name == "mli_composite_fwd":
        R, N = a[10], a[0].N
        return 0.0, float(R * N * (f32 * (1 + 3 + 1 + 8) + f32) + R * 64)
    if name == "mli_composite_bwd":
        R, N = a[10], a[0].N
        return 0.0, float(R * N * f32 * (1 + 3 + 1 + 8 + 1 + 8 + 1 + 3) + R * 64)
    if name == "mli_geometry_fwd":
        M, taps = a[1], a[3]
        return 0.0, float(M * ((1 + taps) * f32 + 6 * f32 + 48 * 2))
    if name == "mli_geometry_bwd":
        M, taps = a[1], a[3]
        return 0.0, float(M * ((1 + taps) * f32 + 9 * f32 + 48 * f32))
    if name == "mli_losses_fwd_bwd":
        R, N = a[6], a[7]
        return 0.0, float(R * N * f32 * 12)
    return None


def roofline_kernels(prof, work, steps, hbm_peak, tf_peak):
    """Per entry point: us/step, algorithmic TFLOP/s and GB/s, fraction of the nearer roof."""
    rows = []
    for k, (calls, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        us = 1e3 * ms / steps
        row = {"entry": k, "calls_per_step": calls / steps, "us_per_step": round(us, 2)}
        w = work.get(k)
        if w is not None and ms > 0:
            tf = w[0] / (ms * 1e-3) / 1e12
            gbs = w[1] / (ms * 1e-3) / 1e9
            ft, fh = tf / tf_peak, gbs / hbm_peak
            if k.startswith("mli_encode_rays"):
                # hash-grid gather / scatter: the algorithmic bytes are the 32-byte sectors REQUESTED (8 corners x
                # levels per sample); most are served by L1 / L2 (ncu: 1.2 GB / 0.9 GB of DRAM traffic per step for
                # 7.2 GB / 11.1 GB requested), so the roof is the L1TEX/L2 sector rate, not HBM: no HBM fraction claimed
                row.update(bytes_per_step=w[1] / steps, gather_gbs=round(gbs, 1), bound="l2_gather", frac=None,
                           note="requested sectors per second; DRAM traffic is 6-12x lower (profiles/r02_summary.md)")
                rows.append(row)
                continue
            row.update(flops_per_step=w[0] / steps, bytes_per_step=w[1] / steps, tflops=round(tf, 1), gbs=round(gbs, 1),
                       bound="tensor" if ft >= fh else "hbm", frac=round(max(ft, fh), 4),
                       frac_tensor=round(ft, 4), frac_hbm=round(fh, 4))
        else:
            row["bound"] = "latency"
        rows.append(row)
    return rows


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(n_rays, threads, seed=0, workload="syn_hotdog_b", state={}):
    """One fwd+bwd of the oracle port (torch CPU, all host threads) on n_rays rays of the bench workload."""
    import torch
    from oracle import port
    torch.set_num_threads(threads)
    exp, dict_size, _, coarse, _ = WORKLOADS[workload]
    if state.get("key") != workload:
        state.clear()
        kw = dict(log2_hashmap_size=dict_size, coarse=coarse)
        if exp.startswith("rene"):
            kw.update(bounding="box", aabb=(-0.66, -0.516, -0.18, 0.66, 0.42, 0.3), white_background=False)
        elif exp.startswith("NRHints"):
            kw.update(white_background=False)
        cfg = port.PathConfig(**kw)
        state.update(key=workload, cfg=cfg,
                     p={k: v.requires_grad_(True) for k, v in port.init_params(cfg, seed=0, generic=False).items()})
    cfg, p = state["cfg"], state["p"]
    if workload == "rene_savannah_b":
        b = rene_batch(n_rays, seed, 0, 1)
        size = (270, 360)
    else:
        b = synthetic_batch(n_rays, seed)
        size = (512, 512)
    c, ray, l = port.rays_from_pose(b["pose"], b["intr"], b["pose_light"], size, b["ray_idx"])
    t0 = time.perf_counter()
    for v in p.values():
        v.grad = None
    out = port.render_rays(p, cfg, c, torch.nn.functional.normalize(ray, dim=-1), l,
                           rands=torch.rand(1, n_rays, cfg.coarse, 1), training=True, progress=0.5)
    total, _, _ = port.total_loss(cfg, out, b)
    total.backward()
    return time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the reference's own algorithm on the host cores (oracle port: the reference is Python that only
    exists in the build container; port.py is pinned against it by tests/test_oracle_vs_reference.py + tests/golden).
    Runs the FULL per-step ray count of the workload whenever warmup+steps then fit in ~4 minutes; otherwise a bounded
    sample (a power-of-two ray count) and `config.rays_per_step_note` says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    rays = WORKLOADS[wl][2]
    threads = os.cpu_count() or 1
    # A step costs t(n) = a + b*n (a = the dense table gradient, independent of the ray count): fit a, b on two probes.
    cpu_reference_step(32, threads, seed=100, workload=wl)  # also builds the table
    t32 = cpu_reference_step(32, threads, seed=101, workload=wl)
    t128 = cpu_reference_step(128, threads, seed=102, workload=wl)
    b = max((t128 - t32) / 96.0, 1e-4)
    a = max(t32 - 32 * b, 0.0)
    budget = 240.0 / max(1, args.steps + args.warmup)
    n_fit = (budget - a) / b
    if n_fit >= rays:
        n_rays, note = rays, "full per-step ray count of the workload"
    else:
        n_rays = int(max(32, n_fit))
        n_rays = 1 << (n_rays.bit_length() - 1)
        note = (f"bounded sample: {n_rays} of the {rays} rays of one step -- the full count would need "
                f"{(a + b * rays) * (args.steps + args.warmup):.0f} s for {args.steps}+{args.warmup} steps on {threads} cores")
    for i in range(args.warmup):
        cpu_reference_step(n_rays, threads, seed=200 + i, workload=wl)
    times = [cpu_reference_step(n_rays, threads, seed=300 + i, workload=wl) for i in range(args.steps)]
    total = sum(times)
    value = n_rays * args.steps / total
    ns = WORKLOADS[wl][3] + 64
    sample = f"{n_rays} of the {rays} rays of one step (x{ns} samples), fwd+bwd, all 5 losses, fp32 torch CPU"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[wl][4], "rays_per_step": n_rays, "rays_per_step_note": note},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def trainer_losses_torch(cfg_trainer, out, data):
    """Bench-only stand-in for the reference trainer's torch-side loss (projects/NeuralLumen/trainer.py:133-149 +
    imaginaire/trainers/base.py:534-544), which lives in the reference and is what the drop-in path `model(data)` ->
    `_compute_loss` -> `backward()` runs; used to time that path, not to check it (tests do that against the oracle)."""
    import torch
    import torch.nn.functional as F
    w = cfg_trainer.loss_weight
    inside = (~out["outside"]).float()
    total = w.render * F.l1_loss(out["rgb"], data["image_sampled"]) * 3
    g = out["gradients"]
    total = total + w.eikonal * (((g.norm(dim=-1) - 1.0) ** 2).nan_to_num(0.0, 0.0, 0.0) * inside).mean()
    total = total + w.curvature * (out["hessians"].sum(dim=-1).abs().nan_to_num(0.0, 0.0, 0.0) * inside).mean()
    if hasattr(w, "intrinsic"):
        pi = cfg_trainer.para_intrinsic_loss

        def mm(x, r):
            x = x.detach()
            return r[0] + (x - x.min()) / torch.clamp(x.max() - x.min(), min=1e-6) * (r[1] - r[0])
        w_sha = mm(data["pseudo_sha_sampled"], pi.weight_map_range_shading)
        w_ref = torch.minimum(mm(data["pseudo_visibility_certainty_sampled"], pi.weight_map_range_visibility), w_sha)
        total = total + w.intrinsic * (((out["o_r"] - data["pseudo_ref_sampled"]).abs() * w_ref).mean() * pi.factor_ref
                                       + ((out["o_s"] - data["pseudo_sha_sampled"]).abs() * w_sha).mean() * pi.factor_sha)
    if hasattr(w, "regularize_re"):
        pr = cfg_trainer.para_regularize_re_loss
        o = out["o_re"]
        z = torch.zeros((), device=o.device)
        total = total + w.regularize_re * (torch.where(o < 0, o, z).abs().mean() * pr.factor_negative
                                           + torch.where(o >= 0, o, z).pow(pr.exponent_positive).mean() * pr.factor_positive)
    return total


def timed(fn, steps, barrier):
    """Device time of `steps` calls of fn(i), barrier + synchronize on both sides.  -> (ms, host enqueue ms)"""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    h0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    host_ms = (time.perf_counter() - h0) * 1e3
    e1.record()
    barrier()
    return e0.elapsed_time(e1), host_ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MLI_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--grad", default="full", choices=["full", "heads"], help="full-grad (primary) or stage-b as shipped")
    ap.add_argument("--workload", default="syn_hotdog_b", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the render / drop-in / other-workload sections")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from mli_nerf_b200 import _lib
    from mli_nerf_b200.dist import make_reducer
    from mli_nerf_b200.losses import loss_cfg_from_trainer
    from mli_nerf_b200.model import Model
    from mli_nerf_b200.optim import FusedAdamW

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        print(f"[bench] --warmup {args.warmup} < 3: the timing rules ask for at least 3 warm-up steps", file=sys.stderr)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert _lib.device_ok(), "bench needs a B200 (no CPU fallback)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build(name, grad):
        cfg = workload_cfg(name, args.precision)
        torch.manual_seed(0)
        model = Model(cfg.model, cfg.data).cuda().train()
        model.progress = 0.5
        if grad == "heads":
            for n, p in model.named_parameters():
                p.requires_grad_("neural_rgb" in n)
        return cfg, model, loss_cfg_from_trainer(cfg.trainer)

    n_batches = 8
    RAYS = WORKLOADS[args.workload][2]
    cfg, model, lcfg = build(args.workload, args.grad)
    reducer = make_reducer(model, world) if world > 1 else None
    if reducer is not None:
        reducer.attach(model.engine)  # hash-table gradient slabs are exchanged while the backward is still running
        reducer.warm_up()             # communicator start-up (collectives only, no steps): see dist.py
    host = [{k: v.pin_memory() for k, v in workload_batch(args.workload, i, rank, world).items()} for i in range(n_batches)]
    dev = [{k: v.cuda(non_blocking=True) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    # N > 1 launches eagerly: capturing the side-stream NCCL collectives inside the step's CUDA graph did not complete on
    # the test box (MLI_GRAPH_MULTI=1 re-enables the attempt)
    use_graph = (not args.no_graph) and (world == 1 or os.environ.get("MLI_GRAPH_MULTI", "0") == "1")
    hook = reducer.exchange_grads if reducer is not None else None

    def step(batch, graph=use_graph):
        return model.fused_train_step(batch, lcfg, use_graph=graph, after_backward=hook)

    # N > 1: the first few dozen steps of a process group run well below the steady state (lazy NCCL connections per
    # message size and protocol, allocator pools of eight processes still growing: 13.4 -> 9.5 ms per step over the first
    # 40 steps in round 1), so at least 40 untimed steps are run there; the line reports the warm-up actually done.
    n_warm = args.warmup if world == 1 else max(args.warmup, 40)
    for i in range(n_warm):
        step(dev[i % n_batches])
    barrier()

    # ---- device-resident timing (value) + clocks ---------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    ms, host_ms = timed(lambda i: step(dev[i % n_batches]), args.steps, barrier)
    # per-entry-point CUDA-event profile of the same steps, launched eagerly (events bracket every C-ABI call on the
    # launching stream); also counts our kernel launches per step and the algorithmic work of every call
    # (the real step runs the weight-gradient GEMMs on a second stream next to the trunk backward and the scatter; an
    # event pair around a launch that shares the GPU measures the pair, not the kernel, so the profile pass keeps
    # everything on one stream and every launch is timed alone)
    _lib.LAUNCH_COUNT = 0
    overlap, model.engine.overlap_wgrad = model.engine.overlap_wgrad, False
    _lib.profile_begin(kernel_work)
    for i in range(args.steps):
        step(dev[i % n_batches], graph=False)
    prof = _lib.profile_end()
    model.engine.overlap_wgrad = overlap
    work = _lib.profile_work()
    launches = _lib.LAUNCH_COUNT
    clocks = sampler.stop()

    # ---- end-to-end timing: pinned host inputs -> H2D, step, loss -> D2H, every step -------------------------------
    d2h = [0]

    def e2e_step(i):
        b = {k: v.cuda(non_blocking=True) for k, v in host[i % n_batches].items()}
        lv = step(b).cpu()
        d2h[0] = lv.numel() * lv.element_size()
    ms_e2e, _ = timed(e2e_step, args.steps, barrier)

    # ---- informational: the same device-resident step followed by the AdamW update of every parameter -------------
    # (the metric is fwd+bwd, SURVEY.md 8f rank 1 keeps the optimizer out of it; this line shows what it adds.  N > 1:
    # the reducer's optimizer -- rank-owned AdamW shard of the table + parameter all-gather when the exchange is a
    # reduce-scatter)
    opt = reducer.make_optimizer(lr=1e-3, weight_decay=1e-2) if reducer is not None else \
        FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=1e-2)

    def opt_step(i):
        step(dev[i % n_batches])
        opt.step()
    for i in range(3):
        opt_step(i)
    ms_opt, _ = timed(opt_step, args.steps, barrier)

    if reducer is not None:  # the extra sections below run without a gradient exchange
        torch.cuda.synchronize()
        exchange = reducer.describe()
        reducer.close()
        del opt
    else:
        exchange = None
    extras = {}
    if not args.no_extras:
        extras = run_extras(args, world, rank, model, cfg, lcfg, dev, barrier, build, step)
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_opt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_opt = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * RAYS * args.steps / (ms * 1e-3)
    e2e = world * RAYS * args.steps / (ms_e2e * 1e-3)
    hbm_peak, tf_peak, peak_src = peaks()
    # Dominant kernel class = the dense layers (the fused-MLP work, 806.0 MFLOP per ray fwd+bwd at 4 taps,
    # full-grad; 631.3 heads-only), i.e. every tcgen05 GEMM entry point (+ the CUDA-core ones in fp32 mode).  Their
    # summed CUDA-event time over the timed steps is the denominator of `achieved`.
    dense_keys = ("mli_linear", "mli_rowdot", "mli_tc_linear", "mli_tc_wgrad", "mli_tc_sdf_trunk", "mli_tc_rowdot",
                  "mli_tc_heads")
    dense = [k for k in prof if k.startswith(dense_keys)]
    dense_ms = sum(prof[k][1] for k in dense)
    prof_ms = sum(v[1] for v in prof.values())
    frac_flops = 1.0 if args.grad == "full" else 631.3 / 806.0
    n_s = WORKLOADS[args.workload][3] + 64
    # SURVEY 8d formula: fwd+bwd = 6 * N * (99 328 + taps * 33 792 + 805 120) + 2 * Q_s * 33 792 (Q_s = N - 16 sampling
    # queries per ray): 806.0 MFLOP at N = 128, 4 taps
    flop_per_ray = frac_flops * (6.0 * n_s * 1039616 + 2.0 * (n_s - 16) * 33792)
    assert abs(6.0 * 128 * 1039616 + 2.0 * 112 * 33792 - MLP_FLOP_PER_RAY) < 0.1e6
    achieved_tf = flop_per_ray * RAYS * args.steps / (dense_ms * 1e-3) / 1e12 if dense_ms > 0 else 0.0
    # DRAM traffic of the same kernels for one step, from the committed ncu capture of this command (profiles/)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("precision") == args.precision and t.get("grad") == args.grad and args.workload == "syn_hotdog_b":
            traffic = t.get("dense_layers_dram_bytes_per_step")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms / args.steps, "host_enqueue_ms_per_step": host_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][4], "grad": args.grad, "precision": args.precision,
                   "rays_per_gpu": RAYS, "samples_per_ray": n_s, "cuda_graph": use_graph, "exchange": exchange,
                   "warmup_requested": args.warmup,
                   "l2": "inputs larger than L2: 1.46 GB hash table + 1.46 GB gradient "
                   "buffer streamed every step (L2 = 126 MB), 8 rotating ray batches"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h[0]},
        "with_optimizer": {"value": world * RAYS * args.steps / (ms_opt * 1e-3), "unit": UNIT,
                           "ms_per_step": ms_opt / args.steps,
                           "note": "informational: step + AdamW over every trainable parameter (not the metric)"},
        "gpu_launches": launches,  # our kernels per `steps` steps (counted on the eager pass; the graph replays the same)
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "dense layers: " + ", ".join(sorted(dense)),
                     "achieved": achieved_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved_tf / tf_peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "timing": "CUDA events around every launch of an eager single-stream pass over the same steps "
                               "(each kernel timed alone; the graph-replayed step overlaps the weight-gradient GEMMs "
                               "with the scatter on a second stream)",
                     "share_of_step": dense_ms / prof_ms if prof_ms > 0 else None,
                     "hbm_peak_gbs": hbm_peak,
                     "kernels": roofline_kernels(prof, work, args.steps, hbm_peak, tf_peak)},
    }
    line.update(extras)
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = 512  # bounded sample: ~10-20 s of host time
        cpu_reference_step(16, threads, seed=1, workload=args.workload)  # warm-up (allocates the table)
        t = cpu_reference_step(n, threads, seed=2, workload=args.workload)
        line["cpu_baseline"] = {"value": n / t, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n} of the {RAYS} rays of one step (x{n_s} samples), fwd+bwd, all 5 "
                                          f"losses, oracle port (fp32 torch CPU), {t:.1f} s"}
        if not args.no_extras and args.workload == "syn_hotdog_b" and "workloads" in line:
            cpu_reference_step(16, threads, seed=1, workload="c0")
            t0 = cpu_reference_step(1024, threads, seed=2, workload="c0")
            line["workloads"]["c0"]["cpu_port"] = {"value": 1024 / t0, "unit": UNIT, "cores": threads,
                                                   "sample": f"1024 of the 4096 rays, {t0:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, world, rank, model, cfg, lcfg, dev, barrier, build, step):
    """Informational sections measured in the same run (all ranks take part; a few steps each)."""
    import torch
    import torch.distributed as dist
    from mli_nerf_b200.model import Model
    out = {}
    k = max(5, min(args.steps, 10))
    RAYS = WORKLOADS[args.workload][2]

    def agg(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return ms

    # ---- the true drop-in path: model(data) -> torch-side losses -> backward() (autograd node = our kernels) --------
    # (single-GPU figure: at N > 1 the reference wraps this path in DDP, which is not what this section measures)
    def dropin(i):
        for p in model.parameters():
            p.grad = None
        o = model(dev[i % len(dev)])
        trainer_losses_torch(cfg.trainer, o, dev[i % len(dev)]).backward()
    if world == 1:
        for i in range(3):
            dropin(i)
        ms, _ = timed(dropin, k, barrier)
        out["autograd_dropin"] = {"value": RAYS * k / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / k,
                                  "note": "model(data) -> torch losses -> backward(), eager launches"}

    # ---- render: 800x800 view, chunks of 20 000 rays, frame's rays partitioned over the ranks ----------------------
    H = W = 800
    model.image_size_val = [H, W]
    pose = dev[0]["pose"]
    intr = torch.tensor([[[1111.0, 0.0, W / 2], [0.0, 1111.0, H / 2], [0.0, 0.0, 1.0]]], device="cuda")
    view = dict(pose=pose, intr=intr, pose_light=dev[0]["pose_light"], idx=torch.zeros(1, dtype=torch.long))
    shard = (rank, world) if world > 1 else None
    render = {"shape": f"{H}x{W} view, chunks of {model.rand_rays_val} rays, eval outputs + maps"
                       + (f", rays partitioned over {world} ranks + all-gather" if world > 1 else ""), "unit": UNIT}
    for key, per_sample in (("rays_per_sec", False), ("rays_per_sec_with_per_sample_tensors", True)):
        if per_sample and world > 1:
            continue  # 64 k rays x 128 samples x 7 floats per rank: gathered nowhere in the reference either
        model.inference(view, per_sample=per_sample, shard=shard)
        n_views = 2
        ms, _ = timed(lambda i: model.inference(view, per_sample=per_sample, shard=shard), n_views, barrier)
        render[key] = H * W * n_views / (agg(ms) * 1e-3)
    model.train()
    out["render"] = render

    # ---- the other workload shapes of SURVEY 8d (a few fused steps each, this N, no gradient exchange at N = 1) ---
    if args.workload == "syn_hotdog_b":
        wl = {}
        del model
        for name in ("NRHints_Pikachu_b", "rene_savannah_b", "syn_hotdog_b_192", "c0"):
            torch.cuda.empty_cache()
            cfg2, m2, lcfg2 = build(name, args.grad)
            batches = [{kk: v.cuda() for kk, v in workload_batch(name, i, rank, world).items()} for i in range(4)]
            use_graph = world == 1 and not args.no_graph
            for i in range(3):
                m2.fused_train_step(batches[i % 4], lcfg2, use_graph=use_graph)
            ms, _ = timed(lambda i: m2.fused_train_step(batches[i % 4], lcfg2, use_graph=use_graph), k, barrier)
            ms = agg(ms)
            r = WORKLOADS[name][2]
            wl[name] = {"value": world * r * k / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / k, "rays_per_gpu": r,
                        "samples_per_ray": WORKLOADS[name][3] + 64, "shape": WORKLOADS[name][4],
                        "note": "fused step, per-rank (no gradient exchange in this section)" if world > 1 else "fused step"}
            del m2
        out["workloads"] = wl
    return out


if __name__ == "__main__":
    main()
