"""Encoder-train images/s (BASELINE.json north_star, second target): one process per GPU, every rank trains a full
replica of BE(startf=16, layer_count=9) on its own slice of the global batch (batch 8 per GPU at 1024^2, weak scaling),
the encoder gradients are averaged with ONE flat-bucket all-reduce (dge_b200.dist.allreduce_grads_) and LREQAdam steps.
Loss = the latent-space terms of the iteration (MSE on w and const against fixed targets) so that the measured work is
the encoder's forward + backward + exchange + step and nothing else.  Device-timed, max over ranks, one JSON line.

    python tools/bench_train.py [--steps 5] [--warmup 3] [--batch 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_train.py --steps 5 --warmup 3

Not the bench.py contract line (that one is the E+G forward metric); context for DESIGN.md section 7.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch

from dge_b200 import dist as ddist
from dge_b200 import ops
from model.E.E import BE
from model.utils.custom_adam import LREQAdam


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    rank, world = ddist.init()
    dev = torch.device("cuda", local)

    torch.manual_seed(0)                      # identical replicas on every rank
    E = BE(startf=16, maxf=512, layer_count=9).to(dev)
    E.set_noise_mode("device")
    ddist.broadcast_buffers_(E)
    opt = LREQAdam(E.parameters(), lr=0.0015, betas=(0.0, 0.99), weight_decay=0)
    g = torch.Generator().manual_seed(1)
    n_global = a.batch * world
    lo, hi = ddist.shard_bounds(n_global, rank, world)
    # synthetic data of the benchmark's shape; the global batch is drawn once and sliced (dist.global_latents idea)
    t_w = torch.randn(n_global, 18, 512, generator=g)[lo:hi].to(dev)
    img = torch.randn(a.batch, 3, 1024, 1024, generator=torch.Generator().manual_seed(100 + rank)).to(dev)

    def step():
        const, w = E(img)
        loss = ((w - t_w) ** 2).mean() + (const ** 2).mean()
        opt.zero_grad()
        loss.backward()
        nbytes = ddist.allreduce_grads_(list(E.parameters()))
        opt.step()
        return nbytes

    for _ in range(a.warmup):
        nbytes = step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ops.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ms = ddist.max_over_ranks(e0.elapsed_time(e1) / a.steps, device=dev)
    if rank == 0:
        print(json.dumps({"metric": "encoder-train images/sec (BE(16,9) fwd+bwd+allreduce+LREQAdam, 1024^2, bs=8/GPU)",
                          "value": n_global / ms * 1e3, "unit": "images/s", "n_gpus": world, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "dtype": "bf16x3 split precision convs (fp32-equivalent), fp32 point-wise", "data": "synthetic",
                          "allreduce_bytes_per_step": int(nbytes), "gpu_launches": ops.launch_count(),
                          "config": {"workload": "BE(startf=16,layer_count=9) training step", "batch_per_gpu": a.batch,
                                     "parallelism": f"dp{world}"}}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
