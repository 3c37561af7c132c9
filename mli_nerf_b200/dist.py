"""Ray-sharded multi-GPU training: one process per GPU, every rank renders its own rays against a full replica of the
hash table + MLPs, and the only exchange per step is the all-reduce(mean) of the parameter gradients -- the DDP
semantics of the reference (/root/reference/imaginaire/trainers/utils/get_trainer.py:80-88), owned explicitly here
because the reference toggles requires_grad after the DDP wrap (projects/NeuralLumen/trainer.py:44-54).

torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests) is the plumbing.  The 1.46 GB
hash-table gradient is reduced in per-level slices so that NCCL can start on the coarse levels while the fine ones are
still in flight on the compute stream; the ~3.6 MB of MLP gradients travel as one flat bucket.
"""
import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, model, world_size=None, level_slices=None, side_stream=True):
        self.model = model
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.level_slices = level_slices  # [(start, end)] element ranges of the hash table per level (optional)
        self.stream = None
        if side_stream and torch.cuda.is_available():
            self.stream = torch.cuda.Stream()

    def _buckets(self):
        """(big tensors reduced in place, list of small grads flattened into one bucket)."""
        big, small = [], []
        for n, p in self.model.named_parameters():
            if p.grad is None:
                continue
            (big if p.grad.numel() >= (1 << 22) else small).append(p.grad)
        return big, small

    def allreduce_grads(self):
        """all-reduce(mean) of every existing .grad; returns after enqueueing on the current stream (stream-ordered)."""
        if self.world <= 1:
            return
        big, small = self._buckets()
        scale = 1.0 / self.world
        cur = torch.cuda.current_stream() if self.stream is not None else None
        if self.stream is not None:
            self.stream.wait_stream(cur)
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _null()
        with ctx:
            for g in big:
                flat = g.view(-1)
                slices = self.level_slices or [(0, flat.numel())]
                for a, b in slices:
                    dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM)
                flat.mul_(scale)
            if small:
                bucket = torch.cat([g.reshape(-1) for g in small])
                dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
                bucket.mul_(scale)
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
