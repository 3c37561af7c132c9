"""Host-side pieces of bench.py (CPU): workload table, ReNe frame assignment, per-call work accounting, and the torch
stand-in for the reference trainer's loss that the drop-in-path timing uses (checked against the oracle)."""
import torch

import bench
from oracle import port


def test_workloads_build_and_batches_have_the_reference_data_keys():
    for name, (exp, dict_size, rays, coarse, _) in bench.WORKLOADS.items():
        cfg = bench.workload_cfg(name, "bf16")
        assert cfg.model.render.num_samples.coarse == coarse and cfg.model.render.rand_rays == rays
        assert cfg.model.object.sdf.encoding.hashgrid.dict_size == dict_size
        b = bench.workload_batch(name, 0, 0, 1)
        assert b["ray_idx"].shape == (1, rays) and b["pose"].shape == (1, 3, 4) and b["pose_light"].shape == (1, 3, 4)
        for k in ("image_sampled", "pseudo_ref_sampled"):
            assert b[k].shape == (1, rays, 3)
        H, W = cfg.data.train.image_size
        assert int(b["ray_idx"].max()) < H * W


def test_rene_frames_follow_distributed_sampler_assignment():
    W = 8
    frames = [[int(bench.rene_batch(16, it, r, W)["idx"]) for r in range(W)] for it in range(3)]
    flat = [f for row in frames for f in row]
    assert len(set(flat)) == len(flat)  # rank r takes perm[it * W + r]: no frame twice within the first epoch
    b = bench.rene_batch(16, 0, 3, W)
    # world->camera poses: rotation part orthonormal, intrinsics rescaled to 270 x 360
    Rm = b["pose"][0, :, :3]
    assert torch.allclose(Rm @ Rm.t(), torch.eye(3), atol=1e-5)
    assert abs(float(b["intr"][0, 0, 2]) - 868.2309 * 360 / 1440) < 1e-2 and abs(float(b["intr"][0, 1, 2]) - 454.0686 * 270 / 1080) < 1e-2


def test_kernel_work_accounting():
    M, K, N, batch = 262144, 256, 256, 3
    a = [None] * 32
    a[6], a[7], a[24], a[25], a[17] = K, N, M, batch, 0
    fl, by = bench.kernel_work("mli_tc_linear", a)
    assert fl == 2.0 * M * K * N * batch and by == M * batch * (K + N) * 2 + K * N * 2 * batch
    a = [None] * 20
    a[8], a[9], a[10], a[11] = M, 256, 256, 3
    assert bench.kernel_work("mli_tc_wgrad", a)[0] == 2.0 * M * 256 * 256 * 3
    assert bench.kernel_work("mli_some_unknown_entry", []) is None
    # 806.0 MFLOP per ray (128 samples, 4 taps, full-grad)
    assert abs(6.0 * 128 * 1039616 + 2.0 * 112 * 33792 - bench.MLP_FLOP_PER_RAY) < 0.1e6


def test_trainer_loss_stand_in_matches_the_oracle():
    from mli_nerf_b200 import config
    torch.manual_seed(0)
    R, N = 64, 16
    cfg = config.experiment("syn_hotdog_b")
    out = dict(rgb=torch.rand(1, R, 3), o_r=torch.rand(1, R, 3), o_s=torch.rand(1, R, 1), o_re=torch.randn(1, R, 3) * 0.1,
               gradients=torch.randn(1, R, N, 3), hessians=torch.randn(1, R, N, 3),
               outside=torch.rand(1, R, 1) > 0.8)
    data = port.synthetic_targets(R, seed=3)
    want, _, _ = port.total_loss(port.PathConfig(), out, data)
    got = bench.trainer_losses_torch(cfg.trainer, out, data)
    assert abs(float(got) - float(want)) < 1e-5 * abs(float(want))


def test_launch_count_indices_point_at_the_right_header_arguments():
    """_lib._n_launches decides from optional pointer arguments how many kernels a call launches (bench.py's
    gpu_launches claim); the positions are checked against include/mli_b200.h so they cannot drift from the ABI."""
    import re
    from mli_nerf_b200 import _lib
    text = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)

    def arg(name, idx):
        m = re.search(r"\b" + name + r"\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
        return " ".join(m.group(1).split()).split(",")[idx].split()[-1].lstrip("*")

    assert arg("mli_tc_wgrad", 16) == "colsum_L"
    assert arg("mli_tc_sdf_trunk_bwd", 9) == "dw_sdf"
    assert arg("mli_rowdot_bwd", 10) == "dA" and arg("mli_rowdot_bwd", 14) == "dw"
    assert arg("mli_composite_bwd", 18) == "d_s_var"
    none17 = [None] * 19
    assert _lib._n_launches("mli_tc_wgrad", none17) == 2
    none17[16] = 1
    assert _lib._n_launches("mli_tc_wgrad", none17) == 3


def test_committed_profile_tables_regenerate_from_the_committed_raw_pages(tmp_path):
    """profiles/traffic.json (read by bench.py for roofline.traffic) and the launch table must be what
    tools/launches_summary.py makes of the committed ncu launch list, and tools/ncu_summary.py must read the committed
    `--set full` raw page: the evidence under profiles/ stays consistent with the tools that produced it."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    prof = os.path.join(root, "profiles")
    out_md = tmp_path / "launches.md"
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "launches_summary.py"),
                        os.path.join(prof, "r02_launches_bf16.csv"), str(out_md), "--precision", "bf16", "--grad", "full"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    regenerated = json.load(open(tmp_path / "traffic.json"))
    committed = json.load(open(os.path.join(prof, "traffic.json")))
    for k in ("dense_layers_dram_bytes_per_step", "dense_layers_us_per_step_ncu", "step_us_ncu"):
        assert abs(regenerated[k] - committed[k]) <= 1e-6 * abs(committed[k]), k
    assert open(out_md).read() == open(os.path.join(prof, "r02_launches_bf16.md")).read()
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"),
                        os.path.join(prof, "r02_step_full_raw.csv"), "--md", "--merge"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for kernel in ("tc_gemm_nt_persist_kernel", "tc_gemm_tn_kernel", "tc_sdf_trunk_fused_kernel", "tc_heads_kernel",
                   "encode_rays_tcl_cached_kernel", "encode_rays_bwd_tcl_v2_kernel", "sdf_trunk_bwd_kernel"):
        assert kernel in r.stdout, kernel


def test_saved_tensor_aliases_break_identity_not_storage():
    """model._aliases (what the render node keeps on ctx): same storage, new tensor objects, containers rebuilt."""
    import torch
    from mli_nerf_b200.model import _aliases
    a, b = torch.arange(6.0), torch.ones(3)
    saved = {"x": a, "lst": [b, None], "tup": (a, 3), "n": 7, "s": "k"}
    out = _aliases(saved)
    assert out["x"] is not a and out["x"].data_ptr() == a.data_ptr()
    assert out["lst"][0] is not b and out["lst"][0].data_ptr() == b.data_ptr() and out["lst"][1] is None
    assert isinstance(out["tup"], tuple) and out["tup"][1] == 3 and out["n"] == 7 and out["s"] == "k"
