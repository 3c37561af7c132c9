#!/usr/bin/env python
"""bench.py -- train rays/s (forward + backward) of the NeuralLumen render hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W                 (N>1: launched through torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W   CPU reference arm (oracle port, host cores)

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C1"): syn_hotdog_b shape -- Neuralangelo hash grid
(16 levels x 8 features, 2^22 entries/level), 2048 rays x 128 samples per step and GPU, 4 gradient taps, rgb_r_s heads,
all five losses, FULL-GRAD (every parameter incl. the 1.46 GB hash table receives a gradient), synthetic rays from a
pinhole camera at distance 3 (f = 711 px, 512x512), random-init weights (reference init).  One "step" = ray generation
+ bounds + hierarchical sampling + forward + in-kernel losses + full backward (optimizer excluded, as in the metric);
with N>1 ranks it also includes the all-reduce(mean) of all parameter gradients (table gradient: copy engines over NVLink
peer memory at N = 2, NCCL beyond -- MLI_TABLE_ALLREDUCE=peer|nccl; MLP gradients: one NCCL bucket).  One JSON line on
rank 0.  Extra keys next to the contract's: `host_enqueue_ms_per_step` (host time to enqueue a step, no sync),
`with_optimizer` (informational: the same step followed by FusedAdamW over every parameter), `profile_ms_per_step`
(CUDA-event time per C-ABI entry point on an eager pass).
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

RAYS = 2048
N_SAMPLES = 128
METRIC = "train_rays_per_sec_fwd_bwd"
UNIT = "rays/s"
# algorithmic MLP flops per ray, fwd+bwd, 4 taps, full-grad (unpadded layer shapes)
MLP_FLOP_PER_RAY = 806.0e6
WORKLOAD = "syn_hotdog_b shape: hash grid 16x8 T=2^22, 2048 rays x 128 samples/GPU, 4 taps, rgb_r_s, 5 losses, full-grad"


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


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._halt = index, [], set(), threading.Event()
        self.sm_max = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [v.strip() for v in out.split(",")]
                self.samples.append(float(f[0]))
                self.sm_max = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(n_rays, threads, seed=0, state={}):
    """One fwd+bwd of the oracle port (torch CPU, all host threads) on n_rays rays of the bench workload."""
    import torch
    from oracle import port
    torch.set_num_threads(threads)
    if "p" not in state:
        cfg = port.PathConfig(log2_hashmap_size=22)
        state["cfg"] = cfg
        state["p"] = {k: v.requires_grad_(True) for k, v in port.init_params(cfg, seed=0, generic=False).items()}
    cfg, p = state["cfg"], state["p"]
    b = synthetic_batch(n_rays, seed)
    c, ray, l = port.rays_from_pose(b["pose"], b["intr"], b["pose_light"], (512, 512), b["ray_idx"])
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
    exists in the build container; port.py is pinned against it by tests/test_oracle_vs_reference.py + tests/golden)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: size the per-step ray count so that warmup+steps finish in ~2.5 minutes.  A step costs
    # t(n) = a + b*n (a = the dense 1.46 GB table gradient, independent of the ray count): fit a, b on two probes.
    cpu_reference_step(32, threads, seed=100)  # also builds the 1.46 GB table
    t32 = cpu_reference_step(32, threads, seed=101)
    t128 = cpu_reference_step(128, threads, seed=102)
    b = max((t128 - t32) / 96.0, 1e-4)
    a = max(t32 - 32 * b, 0.0)
    budget = 150.0 / max(1, args.steps + args.warmup)
    n_rays = int(max(32, min(RAYS, (budget - a) / b)))
    n_rays = 1 << (n_rays.bit_length() - 1)
    for i in range(args.warmup):
        cpu_reference_step(n_rays, threads, seed=200 + i)
    times = [cpu_reference_step(n_rays, threads, seed=300 + i) for i in range(args.steps)]
    total = sum(times)
    value = n_rays * args.steps / total
    sample = f"{n_rays} of the {RAYS} rays of one step (x{N_SAMPLES} samples), fwd+bwd, all 5 losses, fp32 torch CPU"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step": n_rays},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MLI_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--grad", default="full", choices=["full", "heads"], help="full-grad (primary) or stage-b as shipped")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dict-size", type=int, default=22)
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from mli_nerf_b200 import _lib, config
    from mli_nerf_b200.dist import GradReducer
    from mli_nerf_b200.losses import loss_cfg_from_trainer
    from mli_nerf_b200.model import Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert max(1, args.warmup) >= 1
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert _lib.device_ok(), "bench needs a B200 (no CPU fallback)"

    cfg = config.experiment("syn_hotdog_b", dict_size=args.dict_size, rand_rays=RAYS)
    cfg.model.mli_precision = args.precision
    torch.manual_seed(0)
    model = Model(cfg.model, cfg.data).cuda().train()
    model.progress = 0.5
    if args.grad == "heads":
        for n, p in model.named_parameters():
            p.requires_grad_("neural_rgb" in n)
    lcfg = loss_cfg_from_trainer(cfg.trainer)
    reducer = GradReducer(model, world, comm_sms=int(os.environ.get("MLI_COMM_SMS", "32"))) if world > 1 else None
    if reducer is not None:
        reducer.attach(model.engine)  # hash-table gradient slabs are all-reduced while the backward is still running

    n_batches = 8
    host = [{k: v.pin_memory() for k, v in synthetic_batch(RAYS, 1000 * rank + i).items()} for i in range(n_batches)]
    dev = [{k: v.cuda(non_blocking=True) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    # N > 1 launches eagerly: capturing the side-stream NCCL all-reduces inside the step's CUDA graph did not complete on
    # the test box (MLI_GRAPH_MULTI=1 re-enables the attempt); eager costs ~0.3 ms of the step
    graph_multi = os.environ.get("MLI_GRAPH_MULTI", "0") == "1"

    def step(batch, graph=not args.no_graph):
        hook = reducer.allreduce_grads if reducer is not None else None
        return model.fused_train_step(batch, lcfg, use_graph=graph and (world == 1 or graph_multi), after_backward=hook)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the first few dozen large all-reduces run slower than the steady state (measured at N = 8: 13.4 ms per step
    # for steps 4-23, 11.3 ms for steps 13-32, 9.5 ms from step ~40 on), so the untimed warm-up is longer there
    n_warm = max(args.warmup, 3) if world == 1 else max(args.warmup, 40)
    for i in range(n_warm):
        step(dev[i % n_batches])
    barrier()

    # ---- device-resident timing (value) + per-entry-point CUDA-event profile + clocks ------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    h0 = time.perf_counter()
    for i in range(args.steps):
        step(dev[i % n_batches])
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps  # host time to enqueue one step (no sync inside the loop)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # per-entry-point CUDA-event profile of the same steps, launched eagerly (events bracket every C-ABI call on the
    # launching stream); also counts our kernel launches per step
    _lib.LAUNCH_COUNT = 0
    _lib.profile_begin()
    for i in range(args.steps):
        step(dev[i % n_batches], graph=False)
    prof = _lib.profile_end()
    launches = _lib.LAUNCH_COUNT
    clocks = sampler.stop()

    # ---- end-to-end timing: pinned host inputs -> H2D, step, loss -> D2H, every step ------------------------------
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    d2h = 0
    for i in range(args.steps):
        b = {k: v.cuda(non_blocking=True) for k, v in host[i % n_batches].items()}
        lv = step(b).cpu()
        d2h = lv.numel() * lv.element_size()
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    # ---- informational: the same device-resident step followed by the fused AdamW update of every parameter -----
    # (the metric is fwd+bwd, SURVEY.md 8f rank 1 keeps the optimizer out of it; this line shows what it adds)
    from mli_nerf_b200.optim import FusedAdamW
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=1e-2)
    for i in range(3):
        step(dev[i % n_batches])
        opt.step()
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for i in range(args.steps):
        step(dev[i % n_batches])
        opt.step()
    o1.record()
    barrier()
    ms_opt = o0.elapsed_time(o1)

    if reducer is not None:
        torch.cuda.synchronize()
        reducer.close()
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
    dense_keys = ("mli_linear", "mli_rowdot", "mli_tc_linear", "mli_tc_wgrad", "mli_tc_sdf_trunk", "mli_tc_rowdot")
    dense = [k for k in prof if k.startswith(dense_keys)]
    dense_ms = sum(prof[k][1] for k in dense)
    prof_ms = sum(v[1] for v in prof.values())
    frac_flops = 1.0 if args.grad == "full" else 631.3 / 806.0
    achieved_tf = MLP_FLOP_PER_RAY * frac_flops * RAYS * args.steps / (dense_ms * 1e-3) / 1e12 if dense_ms > 0 else 0.0
    # DRAM traffic of the same kernels for one step, from the committed ncu capture of this command (profiles/)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("precision") == args.precision and t.get("grad") == args.grad:
            traffic = t.get("dense_layers_dram_bytes_per_step")
    # per-kernel HBM view (algorithmic bytes are in DESIGN.md section 4): time share of every entry point
    shares = {k: round(v[1] / prof_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]} if prof_ms else {}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms / args.steps, "host_enqueue_ms_per_step": host_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "grad": args.grad, "precision": args.precision, "rays_per_gpu": RAYS,
                   "samples_per_ray": N_SAMPLES, "cuda_graph": (not args.no_graph) and (world == 1 or graph_multi),
                   "l2": "inputs larger than L2: 1.46 GB hash table + 1.46 GB gradient "
                   "buffer streamed every step (L2 = 126 MB), 8 rotating ray batches"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h},
        "with_optimizer": {"value": world * RAYS * args.steps / (ms_opt * 1e-3), "unit": UNIT,
                           "ms_per_step": ms_opt / args.steps,
                           "note": "informational: step + FusedAdamW over every trainable parameter (not the metric)"},
        "gpu_launches": launches,  # our kernels per `steps` steps (counted on the eager pass; the graph replays the same)
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "dense layers: " + ", ".join(sorted(dense)),
                     "achieved": achieved_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved_tf / tf_peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "share_of_step": dense_ms / prof_ms if prof_ms > 0 else None,
                     "hbm_view": {"peak_gbs": hbm_peak,
                                  "note": "unfused layer-by-layer GEMMs are HBM-bound: see profiles/ for GB/s per kernel"},
                     "time_shares": shares},
        "profile_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
    }
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = 512  # bounded sample: ~10-20 s of host time
        cpu_reference_step(16, threads, seed=1)  # warm-up (allocates the 1.46 GB table)
        t = cpu_reference_step(n, threads, seed=2)
        line["cpu_baseline"] = {"value": n / t, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n} of the {RAYS} rays of one step (x{N_SAMPLES} samples), fwd+bwd, all 5 "
                                          f"losses, oracle port (fp32 torch CPU), {t:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
