"""Ray-sharded multi-GPU training: one process per GPU, every rank renders its own rays against a full replica of the
hash table + MLPs, and the only exchange per step is the all-reduce(mean) of the parameter gradients -- the DDP
semantics of the reference (/root/reference/imaginaire/trainers/utils/get_trainer.py:80-88), owned explicitly here
because the reference toggles requires_grad after the DDP wrap (projects/NeuralLumen/trainer.py:44-54).

torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests) is the plumbing.  The 1.46 GB
hash-table gradient is reduced in per-level slices so that NCCL can start on the coarse levels while the fine ones are
still in flight on the compute stream; the ~3.6 MB of MLP gradients travel as one flat bucket.
"""
import os

import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, model, world_size=None, level_slices=None, side_stream=True, comm_sms=8):
        self.comm_sms = comm_sms
        self.model = model
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.level_slices = level_slices  # [(start, end)] element ranges of the hash table per level (optional)
        self.stream = None
        if side_stream and torch.cuda.is_available():
            self.stream = torch.cuda.Stream(priority=-1)  # communication kernels go first when SMs free up
        self._early = {}  # data_ptr of gradients whose slabs were already reduced through the engine hook

    # -- overlap of the hash-table gradient exchange with the rest of the backward pass ------------------------------
    def attach(self, engine):
        """Have `engine` hand over each level group's slab of the hash-table gradient as soon as its scatter kernel has
        been launched: the all-reduce of that slab then runs on the side stream (NCCL over NVLink) while the compute
        stream continues with the next level group and the deferred weight-gradient GEMMs."""
        engine.table_grad_hook = self._on_table_slab if self.world > 1 else None
        # the table gradient is produced last; holding the weight-gradient GEMMs back until its scatter is launched gives
        # the all-reduce independent work to overlap with (MLI_WGRAD_LAST=0: single-GPU schedule, for A/B runs)
        engine.wgrad_after_scatter = self.world > 1 and os.environ.get("MLI_WGRAD_LAST", "1") == "1"
        if self.world > 1 and torch.cuda.is_available():
            from . import _lib
            # leave SMs for the NCCL kernels: a persistent GEMM CTA owns its SM's whole shared memory, so without
            # this the all-reduce only advances in the gaps between kernels
            _lib.set_sm_limit(148 - self.comm_sms)

    def _avg(self, t):
        if dist.get_backend() == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG)  # averaged inside NCCL: no extra pass over the 1.46 GB
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.mul_(1.0 / self.world)

    def _on_table_slab(self, flat, a, b):
        if self.stream is None:
            self._avg(flat[a:b])
        else:
            ev = torch.cuda.Event()
            ev.record()  # the slab is final once everything enqueued so far on the compute stream has run
            self.stream.wait_event(ev)
            flat.record_stream(self.stream)
            with torch.cuda.stream(self.stream):
                self._avg(flat[a:b])
        self._early[flat.data_ptr()] = self._early.get(flat.data_ptr(), 0) + (b - a)

    def _buckets(self):
        """(big tensors reduced in place, list of small grads flattened into one bucket)."""
        big, small = [], []
        for n, p in self.model.named_parameters():
            if p.grad is None:
                continue
            if self._early.get(p.grad.data_ptr(), 0) == p.grad.numel():
                continue  # every slab of this gradient already went through the hook during backward
            (big if p.grad.numel() >= (1 << 22) else small).append(p.grad)
        self._early = {}
        return big, small

    def allreduce_grads(self):
        """all-reduce(mean) of every existing .grad; returns after enqueueing on the current stream (stream-ordered)."""
        if self.world <= 1:
            return
        big, small = self._buckets()
        cur = torch.cuda.current_stream() if self.stream is not None else None
        if self.stream is not None:
            self.stream.wait_stream(cur)
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _null()
        with ctx:
            for g in big:
                flat = g.view(-1)
                slices = self.level_slices or [(0, flat.numel())]
                for a, b in slices:
                    self._avg(flat[a:b])
            if small:
                bucket = torch.cat([g.reshape(-1) for g in small])
                self._avg(bucket)
                off = 0
                for g in small:
                    g.copy_(bucket[off:off + g.numel()].view_as(g))
                    off += g.numel()
        if self.stream is not None:
            cur.wait_stream(self.stream)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def shard_rays(n_rays, rank, world):
    """Contiguous partition of a frame's rays over ranks (inference: SURVEY.md section 8e)."""
    per = (n_rays + world - 1) // world
    return min(rank * per, n_rays), min((rank + 1) * per, n_rays)
