"""One-process-per-GPU plumbing for the inversion path.

The forward path (E forward + frozen G forward) shards by sample with NO data-path collective: every rank holds full
replicas of G and E and works on a contiguous slice of the global batch.  The reference seeds every process
identically (`set_seed(iteration % 30000)`, training_utils.py:46-51), so the GLOBAL latent batch is generated from
the seed and then sliced -- otherwise all ranks would invert the same samples.  The one real exchange step of the
training loop is the all-reduce of E's gradients after each backward (E_align_s2.py:205,220); it is done on ONE flat
bucket (97 MB for BE(16,9)): NVSwitch collectives are latency- not link-bound, so a single large all-reduce beats
per-tensor calls.  Backends: "nccl" on GPUs, "gloo" in the CPU tests (tests/test_dist_cpu.py, world_size 2).
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / MASTER_*); no-op for 1 process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend)
    return rank(), world_size()


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_bounds(n_global, rank_, world):
    """Contiguous [lo, hi) slice of a global batch for one rank (remainder spread over the first ranks)."""
    if world <= 0 or not (0 <= rank_ < world):
        raise ValueError(f"bad rank/world: {rank_}/{world}")
    base, rem = divmod(n_global, world)
    lo = rank_ * base + min(rank_, rem)
    return lo, lo + base + (1 if rank_ < rem else 0)


def shard_batch(t, rank_=None, world=None):
    """This rank's contiguous slice of a global batch tensor (dim 0)."""
    rank_ = rank() if rank_ is None else rank_
    world = world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank_, world)
    return t[lo:hi]


def global_latents(seed, n_global, dim=512, rank_=None, world=None, device="cpu"):
    """z for this rank: the global batch is drawn from `seed` on the CPU (the reference's RNG placement), then sliced."""
    g = torch.Generator().manual_seed(int(seed))
    z = torch.randn(n_global, dim, generator=g)
    return shard_batch(z, rank_, world).to(device)


def max_over_ranks(value, device="cpu"):
    """Device-timed milliseconds -> max over ranks (the number bench.py reports)."""
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_grads_(params, average=True):
    """Sum (or average) the gradients of `params` across ranks through ONE flat bucket, in place.
    Parameters without a gradient contribute zeros (a rank whose slice did not touch them)."""
    params = [p for p in params if p.requires_grad]
    if world_size() == 1 or not params:
        return 0
    dev, dt = params[0].device, params[0].dtype
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), dtype=dt, device=dev)
    off = 0
    for p, n in zip(params, sizes):
        if p.grad is not None:
            flat[off:off + n].copy_(p.grad.reshape(-1))
        off += n
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat.div_(world_size())
    off = 0
    for p, n in zip(params, sizes):
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel() * flat.element_size()


class GradBucket:
    """The encoder's gradients as views of ONE persistent flat buffer, all-reduced while the backward is still running.

    `p.grad` of every parameter is a view into `flat`, so there is no gather / scatter pass and the optimiser's
    multi-tensor tables never change.  The flat buffer is cut into `n_chunks` contiguous chunks in PARAMETER order.  A
    post-accumulate-grad hook counts down the parameters of a chunk; when the last one has its gradient the chunk's
    all-reduce is enqueued (async: NCCL's stream waits for the backward kernels issued so far and then runs beside the
    rest of the backward).  Backward walks the encoder from the last block to the first: the 512-channel blocks -- ~95 %
    of the 97 MB -- are ready first and their exchange hides behind the high-resolution blocks' data / weight gradients
    (~8 ms at 1024^2); only the last small chunk is exposed.  `finish()` (LREQAdam.step calls it) enqueues whatever was
    not sent, makes the current stream wait for all of it and re-arms the counters for the next backward
    (E_align_s2.py:205,220 runs two backward / step pairs per iteration)."""

    def __init__(self, params, n_chunks=4, average=True):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "GradBucket: no trainable parameters"
        dev, dt = self.params[0].device, self.params[0].dtype
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), dtype=dt, device=dev)
        self.average = average
        self.views, off = [], 0
        for p, n in zip(self.params, sizes):
            v = self.flat[off:off + n].view_as(p)
            if p.grad is not None:
                v.copy_(p.grad)
            p.grad = v
            self.views.append(v)
            off += n
        # chunk boundaries on parameter boundaries, ~equal bytes
        total, target = off, max(1, -(-off // max(1, n_chunks)))
        self.chunks, start, acc, first = [], 0, 0, 0
        for i, n in enumerate(sizes):
            acc += n
            if acc - start >= target or i == len(sizes) - 1:
                self.chunks.append({"lo": start, "hi": acc, "params": list(range(first, i + 1))})
                start, first = acc, i + 1
        self.chunk_of, self.index_of = {}, {}
        for ci, c in enumerate(self.chunks):
            for i in c["params"]:
                self.chunk_of[id(self.params[i])] = ci
                self.index_of[id(self.params[i])] = i
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.bytes_per_exchange = total * self.flat.element_size()
        self.enabled = True
        self._arm()

    def _arm(self):
        self._pending = [len(c["params"]) for c in self.chunks]
        self._work = [None] * len(self.chunks)

    def _launch(self, ci):
        if self._work[ci] is not None or world_size() == 1 or not self.enabled:
            return
        c = self.chunks[ci]
        buf = self.flat[c["lo"]:c["hi"]]
        if self.average and dist.get_backend() == "nccl":
            self._work[ci] = (dist.all_reduce(buf, op=dist.ReduceOp.AVG, async_op=True), None)
        else:
            self._work[ci] = (dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True), buf if self.average else None)

    def _on_grad(self, p):
        # autograd wrote / accumulated p.grad; if a fresh tensor was installed (p.grad was None), move it into the bucket
        i = self.index_of[id(p)]
        if p.grad is not self.views[i]:
            if p.grad is not None:
                self.views[i].copy_(p.grad)
            p.grad = self.views[i]
        ci = self.chunk_of[id(p)]
        self._pending[ci] -= 1
        if self._pending[ci] == 0:
            self._launch(ci)

    def zero(self):
        """zero_grad that keeps the views alive (set_to_none would detach the parameters from the bucket)."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def finish(self):
        """All gradients of this backward are exchanged when this returns (in stream order).  -> bytes exchanged."""
        sent = 0
        if world_size() > 1 and self.enabled:
            for ci in range(len(self.chunks)):
                self._launch(ci)
            for work, div in self._work:
                work.wait()
                if div is not None:
                    div.div_(world_size())
            sent = self.bytes_per_exchange
        self._arm()
        return sent

    def close(self):
        for h in self._handles:
            h.remove()


def broadcast_buffers_(module, src=0):
    """Keep per-rank stateful buffers identical (StyleGAN2 `w_avg`, spectral-norm `u`: SURVEY 8e caveat 2)."""
    if world_size() == 1:
        return
    for b in module.buffers():
        dist.broadcast(b.data, src=src)
