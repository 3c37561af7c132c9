"""Ray-sharded multi-GPU training: one process per GPU, every rank renders its own rays against a full replica of the
hash table + MLPs, and the only exchange per step is that of the parameter gradients.

Two exchanges of the 1.46 GB hash-table gradient (`GradReducer(table_mode=...)`, env MLI_TABLE_EXCHANGE):

* ``allreduce`` -- the DDP semantics of the reference (/root/reference/imaginaire/trainers/utils/get_trainer.py:80-88):
  every rank ends the step with the full mean gradient; any optimizer can follow.  2 (W-1)/W x 1.46 GB per link
  direction and step.
* ``reduce_scatter`` (default of `make_reducer`) -- rank r ends the step with the mean gradient of ITS 1/W shard of every
  level-group slab only ((W-1)/W x 1.46 GB per direction: half the bytes inside the fwd+bwd metric); the optimizer that
  follows is `ShardedTableAdamW`: AdamW on the owned shard (1/W of the 10 GB optimizer pass, 1/W of the moment
  memory) and an all-gather of the updated PARAMETERS.  Same arithmetic as all-reduce + dense AdamW: every table entry
  is updated exactly once, by its owner, from the same mean gradient.

The gradient slabs are exchanged as soon as the scatter kernel of a level group has been launched (engine hook), on a
side stream, while the weight-gradient GEMMs run; the ~3.6 MB of MLP gradients travel as one flat NCCL bucket.
torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests) is the plumbing.

Transport (`PeerTableReducer`, default for two ranks): NCCL needs 32 resident CTAs to move this payload at full rate and
therefore cannot overlap with the weight-gradient GEMMs that own the SMs at that point of the step.  Here every rank
maps its peers' gradient buffers (CUDA IPC) and moves the data with the copy engines instead: rank r owns shard r of
each slab, pulls that shard of every peer's buffer over NVLink, sums (`mli_reduce_slots`) -- zero SMs for the
transfers.  Cross-rank ordering is a one-element NCCL all-reduce enqueued in stream order.  With more ranks NCCL
(in-switch reduction) is faster and is the default -- `MLI_TABLE_ALLREDUCE=peer|nccl` overrides.
"""
import os
import sys

import torch
import torch.distributed as dist


def make_reducer(model, world, **kw):
    """The gradient exchange bench.py / a trainer uses: reduce-scatter + rank-owned optimizer shard unless
    MLI_TABLE_EXCHANGE=allreduce asks for the reference's DDP semantics."""
    mode = os.environ.get("MLI_TABLE_EXCHANGE", "reduce_scatter")
    return GradReducer(model, world, comm_sms=int(os.environ.get("MLI_COMM_SMS", "32")), table_mode=mode, **kw)


def num_sms():
    if torch.cuda.is_available():
        return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return 148


class GradReducer:
    def __init__(self, model, world_size=None, level_slices=None, side_stream=True, comm_sms=32, table_mode="allreduce"):
        if table_mode not in ("allreduce", "reduce_scatter"):
            raise ValueError(f"table_mode {table_mode!r}: expected 'allreduce' or 'reduce_scatter'")
        self.table_mode = table_mode
        self.comm_sms = comm_sms
        self.model = model
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.level_slices = level_slices  # [(start, end)] element ranges of the hash table per level (optional)
        self.stream = None
        if side_stream and torch.cuda.is_available():
            self.stream = torch.cuda.Stream(priority=-1)  # communication kernels go first when SMs free up
        self._early = {}  # data_ptr of gradients whose slabs were already reduced through the engine hook
        self.peer = None  # PeerTableReducer (set by attach)
        self._engine = None
        self._limit_sms = False
        self._limited = False
        self._hooked = None
        self._transport = "nccl"
        # reduce_scatter mode: after exchange_grads(), [(elem_begin, elem_end, mean gradient of that owned range)]
        self.table_grad_shards = []
        self._shard_buf = None
        self._shard_off = 0

    # -- overlap of the hash-table gradient exchange with the rest of the backward pass ------------------------------
    def attach(self, engine):
        """Have `engine` hand over each level group's slab of the hash-table gradient as soon as its scatter kernel has
        been launched: the exchange of that slab then runs on the side stream (NCCL / copy engines over NVLink) while
        the compute stream continues with the next level group and the deferred weight-gradient GEMMs."""
        engine.table_grad_hook = self._on_table_slab if self.world > 1 else None
        self._engine = engine
        # the table gradient is produced last; holding the weight-gradient GEMMs back until its scatter is launched gives
        # the exchange independent work to overlap with (MLI_WGRAD_LAST=0: single-GPU schedule, for A/B runs)
        engine.wgrad_after_scatter = self.world > 1 and os.environ.get("MLI_WGRAD_LAST", "1") == "1"
        if self.world > 1 and dist.is_initialized():
            # every rank must cut the table into the same slabs: the NCCL calls / tokens are matched by order
            n_slabs = len(engine.level_groups()) if hasattr(engine, "level_groups") else 0
            seen = [None] * self.world
            dist.all_gather_object(seen, n_slabs)
            if any(v != n_slabs for v in seen):
                raise RuntimeError(f"GradReducer.attach: ranks disagree on the number of table slabs: {seen}")
        if self.world > 1 and torch.cuda.is_available():
            from . import _lib
            # measured (bench.py, full-grad, all-reduce): N = 2  peer 7.05 ms / NCCL 7.42 ms per step;  N = 8  peer 10.26 ms
            # / NCCL 9.32 ms (NCCL reduces inside the NVSwitch there; the copy engines reach only 400-460 GB/s inbound
            # with seven sources) -> the peer transport is the default for two ranks, NCCL beyond
            mode = os.environ.get("MLI_TABLE_ALLREDUCE", "peer" if self.world == 2 else "nccl")
            backend = str(dist.get_backend())
            if dist.get_rank() == 0:
                print(f"[mli] table-gradient exchange: {self.table_mode} over {mode} (torch.distributed backend {backend})",
                      file=sys.stderr, flush=True)
            if mode == "peer" and "nccl" in backend:
                try:
                    self.peer = PeerTableReducer(engine.n_table_params(), engine.device,
                                                 max_slabs=len(engine.level_groups()))
                    engine.table_grad_buffer = self.peer.buf  # the scatter writes straight into the IPC-shared buffer
                    self._transport = "peer"
                except _lib.MliError as e:  # raised on every rank or on none: all ranks take the NCCL exchange together
                    print(f"[mli] rank {dist.get_rank()}: {e}; using the NCCL exchange", file=sys.stderr, flush=True)
                    self.peer = None
                    self._limit_sms = True
            elif mode in ("peer", "nccl"):
                # NCCL needs its 32 channels = 32 resident CTAs to move this payload at full rate (16 channels: 4.3 ms
                # instead of 2.55 ms at N = 2) and a persistent GEMM CTA owns its SM's whole shared memory: from the
                # first slab on, the persistent kernels of the step are launched on n_sms - comm_sms SMs (_on_table_slab)
                self._limit_sms = True
            else:
                raise _lib.MliError(f"MLI_TABLE_ALLREDUCE={mode}: expected peer or nccl")
        if self.table_mode == "reduce_scatter" and self.world > 1 and self.peer is None and hasattr(engine, "n_table_params"):
            n = engine.n_table_params()
            dev = engine.device if torch.cuda.is_available() else "cpu"
            self._shard_buf = torch.empty((n + self.world - 1) // self.world + 64, dtype=torch.float32, device=dev)

    def describe(self):
        return {"table": self.table_mode, "transport": self._transport, "mlp": "all-reduce (one flat bucket)",
                "world": self.world}

    def warm_up(self, iters=24):
        """Communicator start-up, NOT training steps: the first few dozen large NCCL collectives of a process run well below
        their steady-state rate (measured at N = 8: 13.4 ms per step for steps 4-23, 9.5 ms from step ~40 on).  Runs the
        exchange's own collective `iters` times on a scratch buffer of one slab's size."""
        if self.world <= 1 or not torch.cuda.is_available() or self.peer is not None or self._engine is None:
            return
        groups = self._engine.level_groups()
        n = max(e1 - e0 for _, _, e0, e1 in groups)
        n -= n % self.world
        x = torch.zeros(n, dtype=torch.float32, device=self._engine.device)
        y = torch.zeros(n // self.world, dtype=torch.float32, device=self._engine.device)
        for _ in range(iters):
            if self.table_mode == "reduce_scatter":
                dist.reduce_scatter_tensor(y, x, op=dist.ReduceOp.AVG)
                dist.all_gather_into_tensor(x, y)
            else:
                dist.all_reduce(x, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()

    def close(self):
        """Unmap / free the peer-shared table-gradient buffer (collective: every rank calls it)."""
        self._restore_sms()
        if self._engine is not None:  # detach: the engine goes back to the single-GPU schedule
            self._engine.table_grad_hook = None
            self._engine.wgrad_after_scatter = False
        self.table_grad_shards, self._last_shards = [], []
        if self.peer is None:
            return
        for p in self.model.parameters():
            if p.grad is not None and p.grad.data_ptr() == self.peer.buf.data_ptr():
                p.grad = None
        if self._engine is not None:
            self._engine.table_grad_buffer = None
        self.peer.close()
        self.peer = None

    def _restore_sms(self):
        if self._limited:
            from . import _lib
            _lib.set_sm_limit(num_sms())
            self._limited = False

    def _avg(self, t):
        if dist.get_backend() == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG)  # averaged inside NCCL: no extra pass over the 1.46 GB
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.mul_(1.0 / self.world)

    def _reduce_scatter_avg(self, out, t):
        """out [n/W] = mean over ranks of t[rank*n/W : (rank+1)*n/W]."""
        if dist.get_backend() == "nccl":
            dist.reduce_scatter_tensor(out, t, op=dist.ReduceOp.AVG)
        else:  # gloo (CPU tests) has no reduce-scatter: all-reduce a copy and keep the owned part
            tmp = t.clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
            n = out.numel()
            out.copy_(tmp[self.rank * n:(self.rank + 1) * n]).mul_(1.0 / self.world)

    def _on_table_slab(self, flat, a, b):
        rs = self.table_mode == "reduce_scatter"
        if rs and (b - a) % (4 * self.world) != 0:
            raise RuntimeError(f"reduce_scatter exchange: slab [{a}, {b}) is not divisible into {self.world} 16-byte shards")
        m0, m1 = shard_bounds(a, b, self.rank, self.world)
        if self.peer is not None:
            if flat.data_ptr() != self.peer.buf.data_ptr():
                raise RuntimeError("table gradient was not accumulated in the peer-shared buffer")
            self.peer.reduce_slab(a, b)  # leaves the mean of the owned shard in place (flat[m0:m1])
            if rs:
                self.table_grad_shards.append((m0, m1, flat[m0:m1]))
        else:
            out = None
            if rs:
                out = self._shard_buf[self._shard_off:self._shard_off + (m1 - m0)]
                self._shard_off += m1 - m0
                self.table_grad_shards.append((m0, m1, out))
            if self.stream is None:
                self._reduce_scatter_avg(out, flat[a:b]) if rs else self._avg(flat[a:b])
            else:
                if self._limit_sms and not self._limited:
                    from . import _lib
                    _lib.set_sm_limit(num_sms() - self.comm_sms)  # until exchange_grads() has joined
                    self._limited = True
                ev = torch.cuda.Event()
                ev.record()  # the slab is final once everything enqueued so far on the compute stream has run
                self.stream.wait_event(ev)
                flat.record_stream(self.stream)
                with torch.cuda.stream(self.stream):
                    self._reduce_scatter_avg(out, flat[a:b]) if rs else self._avg(flat[a:b])
        self._early[flat.data_ptr()] = self._early.get(flat.data_ptr(), 0) + (b - a)
        self._hooked = flat  # the exchange runs in place on this buffer: the parameter's .grad must alias it

    def _buckets(self):
        """(big tensors reduced in place, list of small grads flattened into one bucket)."""
        big, small = [], []
        hooked, self._hooked = self._hooked, None
        aliased = hooked is None
        for n, p in self.model.named_parameters():
            if p.grad is None:
                continue
            if self._early.get(p.grad.data_ptr(), 0) == p.grad.numel():
                aliased = aliased or p.grad.data_ptr() == hooked.data_ptr()
                continue  # every slab of this gradient already went through the hook during backward
            (big if p.grad.numel() >= (1 << 22) else small).append(p.grad)
        self._early = {}
        if not aliased:
            # e.g. backward() accumulated into an existing .grad (zero_grad(set_to_none=False), micro-batching): the
            # parameter then holds a copy taken while the in-place exchange of the hooked buffer was still in flight
            raise RuntimeError("the hash-table gradient exchanged during backward is not the parameter's .grad: clear the "
                               "gradients with set_to_none=True before every backward, or do not attach() the engine")
        return big, small

    def begin_step(self):
        """Forget the shard list of the previous step (called by exchange_grads' consumers implicitly: the hook of the
        next backward appends to a fresh list)."""
        self.table_grad_shards = []
        self._shard_off = 0

    def exchange_grads(self):
        """Complete the gradient exchange of the step; returns after enqueueing on the current stream (stream-ordered).
        allreduce mode: every existing .grad holds the mean over ranks.  reduce_scatter mode: the MLP / s_var gradients
        hold the mean; of the table gradient only `table_grad_shards` (this rank's shards) is meaningful."""
        if self.world <= 1:
            return
        try:
            big, small = self._buckets()
            rs = self.table_mode == "reduce_scatter"
            if big and rs:
                raise RuntimeError("reduce_scatter exchange needs the engine hook (GradReducer.attach): a large gradient "
                                   "reached exchange_grads() without having been exchanged slab by slab")
            if self.peer is not None and not rs:
                self.peer.gather()  # ahead of the MLP bucket in the NCCL queue: that one waits for the end of the backward
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
            if self.peer is not None:
                self.peer.finish(gather=not rs)
        finally:
            self._restore_sms()
        # the shard list stays valid until the next backward starts appending: hand it to the optimizer, then reset
        self._last_shards, self.table_grad_shards, self._shard_off = self.table_grad_shards, [], 0

    allreduce_grads = exchange_grads  # the all-reduce mode's historical name

    def make_optimizer(self, lr=1e-3, weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8):
        """The optimizer that matches the exchange: dense FusedAdamW after an all-reduce, ShardedTableAdamW after a
        reduce-scatter."""
        from .optim import FusedAdamW, ShardedTableAdamW
        params = [p for p in self.model.parameters() if p.requires_grad]
        if self.table_mode == "reduce_scatter" and self.world > 1:
            return ShardedTableAdamW(self, params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        return FusedAdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)


def shard_bounds(a, b, rank, world):
    """Element range of slab [a, b) owned by `rank`: equal shards, multiples of 4 elements (16-byte vector access)."""
    n = b - a
    sh = ((n + world - 1) // world + 3) // 4 * 4
    return a + min(rank * sh, n), a + min((rank + 1) * sh, n)


class PeerTableReducer:
    """all-reduce(mean) of one large fp32 buffer, slab by slab, over NVLink peer memory with the copy engines.

    Reduce-scatter: once every rank's scatter of a slab has finished [one-element NCCL all-reduce in stream order], rank r
    pulls shard r of every peer's buffer into staging (cudaMemcpyAsync on IPC-mapped pointers: DMA, no SMs) and
    `mli_reduce_slots` writes the mean into its own shard.  All-gather: once every rank has summed a slab [NCCL token],
    rank r pulls the finished shard p from every peer p.  All copies of a rank run back to back on ONE stream -- two
    concurrent inbound transfers share the link at a lower total rate (measured: 400-540 GB/s together, 735 GB/s alone;
    a push runs at 544 GB/s, hence pull for both phases) -- while the sums run beside them on a second stream.  Both link
    directions carry 2 (W-1)/W of the buffer per step, like a ring all-reduce, without occupying SMs."""

    def __init__(self, n_elems, device, max_slab=None, max_slabs=16):
        from . import _lib
        self._lib = _lib
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if not 2 <= self.world <= 8:
            raise _lib.MliError("PeerTableReducer: world size 2..8 (one NVSwitch domain)")
        self.device = torch.device(device)
        self.n = int(n_elems)
        # the persistent gradient buffer: its own cudaMalloc allocation (outside the caching allocator) + IPC handle.
        # Setup errors (no peer access, a rank that does not see its peers' GPUs, out of memory) are agreed on
        # collectively so that every rank raises -- or none.
        err, handle = None, None
        self.buf = self._own_ptr = None
        try:
            self.buf, handle, self._own_ptr = _lib.peer_alloc(self.n, self.device)
            self.buf.zero_()
            # staging for the pulled shards of every slab of one step (the sums run behind the copies)
            self.stage_cap = (self.world - 1) * ((self.n + self.world - 1) // self.world + 8 * max_slabs)
            self.stage = torch.empty(self.stage_cap, dtype=torch.float32, device=self.device)
        except Exception as e:  # noqa: BLE001
            err = e
        self._stage_off = 0
        self._slabs = []
        self.token = torch.zeros(1, dtype=torch.float32, device=self.device)
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.n, handle))
        self.peer_ptr = []
        try:
            for p, (n_p, h) in enumerate(handles):
                if n_p != self.n or h is None:
                    raise _lib.MliError(f"PeerTableReducer: rank {p} has no matching buffer")
                self.peer_ptr.append(self._own_ptr if p == self.rank else _lib.peer_open(h, self.device))
        except Exception as e:  # noqa: BLE001
            err = err or e
        ok = torch.tensor([0.0 if err is not None else 1.0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            self._release()
            raise _lib.MliError(f"PeerTableReducer: setup failed on at least one rank (this rank: {err})")
        self.s_sig = torch.cuda.Stream(device=self.device, priority=-1)
        self.s_copy = torch.cuda.Stream(device=self.device)
        self.s_add = torch.cuda.Stream(device=self.device, priority=-1)
        self._pending = False
        # MLI_PEER_TRACE=1: CUDA-event timeline of one exchange (rank 0 prints it once, after a few warm-up steps)
        self._trace = [] if os.environ.get("MLI_PEER_TRACE", "0") == "1" else None
        self._trace_step = 0
        torch.cuda.synchronize(self.device)
        dist.barrier()

    def _mark(self, label, stream=None):
        if self._trace is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream) if stream is not None else ev.record()
        self._trace.append((label, ev))

    def _others(self):
        return [(self.rank + k) % self.world for k in range(1, self.world)]  # rotated: spreads the load over the peers

    def _token(self, after):
        """Cross-rank ordering point: completes (in stream order) once every rank has reached it.  -> event"""
        self.s_sig.wait_event(after)
        with torch.cuda.stream(self.s_sig):
            dist.all_reduce(self.token)
            ev = torch.cuda.Event()
            ev.record()
        return ev

    def reduce_slab(self, a, b):
        """Enqueue the reduce-scatter of buf[a:b]; the slab is final once everything enqueued so far on the current stream
        has run.  Returns at once (stream-ordered); `gather()` + `finish()` complete the exchange."""
        call, r, W = self._lib.call, self.rank, self.world
        if not 0 <= a <= b <= self.n:
            raise self._lib.MliError("PeerTableReducer: bad slab")
        ev = torch.cuda.Event()
        ev.record()
        self._mark(f"slab {a}: scatter launched (compute stream position)")
        ready = self._token(ev)  # every rank's scatter of this slab has completed
        self._mark(f"slab {a}: all ranks scattered", self.s_sig)
        m0, m1 = shard_bounds(a, b, r, W)
        n = m1 - m0
        slot = (n + 3) // 4 * 4
        off = self._stage_off
        if off + (W - 1) * slot > self.stage_cap:
            raise self._lib.MliError("PeerTableReducer: staging buffer exhausted (too many slabs in one step)")
        self._stage_off += (W - 1) * slot
        self.s_copy.wait_event(ready)
        with torch.cuda.stream(self.s_copy):
            if n > 0:
                for k, p in enumerate(self._others()):
                    call("mli_copy_async", self.stage.data_ptr() + 4 * (off + k * slot), self.peer_ptr[p] + 4 * m0, 4 * n)
            pulled = torch.cuda.Event()
            pulled.record()
            self._mark(f"slab {a}: own shard pulled")
        self.s_add.wait_event(pulled)
        with torch.cuda.stream(self.s_add):
            if n > 0:
                call("mli_reduce_slots", self._own_ptr + 4 * m0, self.stage.data_ptr() + 4 * off, W - 1, slot, n, 1.0 / W)
            summed = torch.cuda.Event()
            summed.record()
            self._mark(f"slab {a}: own shard summed")
        self._slabs.append((a, b, summed))
        self._pending = True

    def gather(self):
        """Enqueue the all-gather of every slab reduced since the last call (after the last `reduce_slab` of the step)."""
        call, W = self._lib.call, self.world
        for a, b, summed in self._slabs:
            all_summed = self._token(summed)  # every rank's shard of this slab holds the mean
            self._mark(f"slab {a}: all ranks summed", self.s_sig)
            self.s_copy.wait_event(all_summed)
            with torch.cuda.stream(self.s_copy):
                for p in self._others():
                    p0, p1 = shard_bounds(a, b, p, W)
                    if p1 > p0:
                        call("mli_copy_async", self._own_ptr + 4 * p0, self.peer_ptr[p] + 4 * p0, 4 * (p1 - p0))
                self._mark(f"slab {a}: gathered")
        self._slabs = []

    def finish(self, gather=True):
        """Current stream waits until the whole buffer holds the mean on every rank (`gather=False`: only this rank's
        shards do -- reduce-scatter) and nobody reads this rank's buffer any more (it may be zeroed for the next step)."""
        if not self._pending:
            return
        if gather:
            self.gather()
        else:
            self._slabs = []
        self._mark("finish(): compute stream position")
        ev = torch.cuda.Event()
        with torch.cuda.stream(self.s_copy):
            ev.record()
        done = self._token(ev)
        cur = torch.cuda.current_stream()
        cur.wait_event(done)
        cur.wait_stream(self.s_add)
        self._mark("finish(): joined")
        self._pending = False
        self._stage_off = 0
        if self._trace is not None:
            self._trace_step += 1
            if self._trace_step == 50 and self.rank == 0:
                torch.cuda.synchronize(self.device)
                t0 = self._trace[0][1]
                print("\n".join(f"[peer trace] {t0.elapsed_time(e):8.3f} ms  {lbl}" for lbl, e in self._trace), flush=True)
            self._trace = []

    def _release(self):
        for p, ptr in enumerate(self.peer_ptr):
            if p != self.rank and ptr is not None:
                self._lib.call("mli_peer_close", ptr)
        self.peer_ptr = []
        dist.barrier()  # nobody maps this rank's buffer any more
        self.buf = None
        if self._own_ptr is not None:
            self._lib.call("mli_peer_free", self._own_ptr)
            self._own_ptr = None

    def close(self):
        if not self.peer_ptr:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier()
        self._release()

class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def shard_rays(n_rays, rank, world):
    """Contiguous partition of a frame's rays over ranks (inference: SURVEY.md section 8e)."""
    per = (n_rays + world - 1) // world
    return min(rank * per, n_rays), min((rank + 1) * per, n_rays)
