"""N>1 host logic on CPU: two gloo ranks.  The forward path shards by sample with no data-path collective;
the one exchange step of the training loop is the flat-bucket all-reduce of the encoder gradients."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deep-gan-encoders_b200")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from dge_b200 import dist as ddist
    from model.utils.custom_adam import LREQAdam  # noqa: F401  (import check: host-side mirror loads without a GPU)
    r, w = ddist.init("gloo")
    assert (r, w) == (rank, world)
    # 1. sharding: ranks see disjoint, contiguous, exhaustive slices of the SAME seeded global batch
    z = ddist.global_latents(1234, 10, 16)
    lo, hi = ddist.shard_bounds(10, rank, world)
    zg = torch.randn(10, 16, generator=torch.Generator().manual_seed(1234))
    assert torch.equal(z, zg[lo:hi])
    # 2. the exchange step: per-rank gradients of a replicated encoder-like parameter set -> average
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.Linear(8, 4))
    frozen = torch.nn.Linear(4, 4)
    for p in frozen.parameters():
        p.requires_grad_(False)
    loss = frozen(lin(z)).pow(2).sum()          # sum over this rank's samples
    loss.backward()
    local = [p.grad.clone() for p in lin.parameters()]
    nbytes = ddist.allreduce_grads_(list(lin.parameters()) + list(frozen.parameters()), average=False)
    assert nbytes == sum(p.numel() for p in lin.parameters()) * 4
    # reference: the gradient of the loss over the WHOLE batch, computed on one process
    lin2 = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.Linear(8, 4))
    lin2.load_state_dict(lin.state_dict())
    frozen(lin2(zg)).pow(2).sum().backward()
    for p, q in zip(lin.parameters(), lin2.parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6)
    # 2b. the same exchange through the persistent bucket the optimiser owns: gradients are views of one flat buffer,
    #     chunks are all-reduced as soon as their last gradient lands, `finish()` waits; two backward passes per
    #     iteration (E_align_s2.py:205,220) and zero_grad keep the views
    lin3 = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.Linear(8, 4))
    lin3.load_state_dict(lin2.state_dict())
    opt = LREQAdam(lin3.parameters(), lr=1e-3, betas=(0.0, 0.99))
    bucket = opt._bucket
    assert bucket is not None and len(bucket.chunks) >= 2
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    for it in range(2):
        opt.zero_grad()
        assert all(p.grad is v for p, v in zip(bucket.params, bucket.views)) and float(bucket.flat.abs().sum()) == 0.0
        (frozen(lin3(z)).pow(2).sum() * (it + 1)).backward()
        assert any(w is not None for w in bucket._work)            # exchange already enqueued by the hooks
        sent = bucket.finish()
        assert sent == bucket.flat.numel() * 4
        for p, q in zip(lin3.parameters(), lin2.parameters()):     # average over ranks of per-rank sums = whole-batch / world
            assert torch.allclose(p.grad, q.grad * (it + 1) / world, rtol=1e-5, atol=1e-6)
    # 3. timing reduction and buffer broadcast
    assert ddist.max_over_ranks(1.0 + rank) == float(world)
    bn = torch.nn.BatchNorm1d(4)
    bn.running_mean.fill_(float(rank + 1))
    ddist.broadcast_buffers_(bn)
    assert float(bn.running_mean[0]) == 1.0
    out.put((rank, float(local[0].abs().sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert [g[0] for g in got] == [0, 1]


def test_shard_bounds_cover_ragged_batches():
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from dge_b200 import dist as ddist
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [ddist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
