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


def broadcast_buffers_(module, src=0):
    """Keep per-rank stateful buffers identical (StyleGAN2 `w_avg`, spectral-norm `u`: SURVEY 8e caveat 2)."""
    if world_size() == 1:
        return
    for b in module.buffers():
        dist.broadcast(b.data, src=src)
