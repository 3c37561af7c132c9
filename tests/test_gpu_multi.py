"""Multi-GPU device paths (need >= 2 B200 on the box; skipped on a single-GPU box).  Host-side logic of the same code is
covered on CPU by tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_table_reducer_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29733", os.path.join(ROOT, "tests", "multi", "peer_allreduce_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "PEER_ALLREDUCE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("table", ["allreduce", "reduce_scatter"])
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_sharded_train_step_gradients(mode, table):
    """Ray-sharded fused step + gradient exchange (all-reduce, or reduce-scatter + rank-owned AdamW shard + parameter
    all-gather) over both transports against NCCL's all-reduce(mean) / torch.optim.AdamW of the per-rank gradients."""
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29734", os.path.join(ROOT, "tests", "multi", "train_step_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, MLI_TABLE_ALLREDUCE=mode, MLI_TABLE_EXCHANGE=table))
    assert out.returncode == 0 and "TRAIN_STEP_EXCHANGE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_inference_matches_single_rank():
    """Model.inference(shard=True): the frame's rays partitioned over the ranks + all-gather == the single-rank maps."""
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29735", os.path.join(ROOT, "tests", "multi", "inference_shard_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "INFERENCE_SHARD_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
