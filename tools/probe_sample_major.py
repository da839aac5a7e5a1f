"""Experiment: run the first encoder blocks SAMPLE BY SAMPLE so that a sample's 67 MB / 33 MB intermediate maps are still in
the 126 MB L2 when the next kernel reads them (batch-wise, every pass over the 537 MB / 268 MB batch tensors goes to HBM).
Times E forward (BE(16,9), 1024^2, batch 8, device noise) batch-major vs sample-major for the first k blocks, eager and as a
CUDA graph.  usage: python tools/probe_sample_major.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch

from dge_b200 import ops
from model.E.E import BE

torch.manual_seed(0)
E = BE(16, 512, 9, 512, 3).cuda().eval()
E.set_noise_mode("device")
img = torch.randn(8, 3, 1024, 1024, device="cuda").clamp_(-1, 1)


def fwd_sample_major(x, k):
    n = x.shape[0]
    outs, ws = [], []
    for s in range(n):
        f, st, mr = E.FromRGB.run_with_stats(x[s:s + 1], 1e-8)
        wl = []
        for i in range(k):
            f, w1, w2 = E.decode_block[i].run(f, stats=(st, mr) if i == 0 else None)
            wl.append((w1, w2))
        outs.append(f)
        ws.append(wl)
    f = ops.F32B.wrap(torch.cat([o.t for o in outs], dim=0), n, outs[0].c, outs[0].h, outs[0].w)
    w = None
    for i in range(k):
        w1 = torch.cat([ws[s][i][0] for s in range(n)], dim=0)
        w2 = torch.cat([ws[s][i][1] for s in range(n)], dim=0)
        w_ = torch.cat((w2.view(n, 1, 512), w1.view(n, 1, 512)), dim=1)
        w = w_ if w is None else torch.cat((w_, w), dim=1)
    for i in range(k, E.layer_count):
        f, w1, w2 = E.decode_block[i].run(f)
        w_ = torch.cat((w2.view(n, 1, 512), w1.view(n, 1, 512)), dim=1)
        w = torch.cat((w_, w), dim=1)
    return f.to_nchw(), w


def timeit(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


out = {}
with torch.no_grad():
    ref_c, ref_w = E(img)
    out["batch_major_eager_ms"] = timeit(lambda: E(img))
    out["batch_major_graph_ms"] = timeit(graphed(lambda: E(img)))
    for k in (1, 2, 3):
        c, w = fwd_sample_major(img, k)
        out[f"sample_major_{k}_shape_ok"] = bool(c.shape == ref_c.shape and w.shape == ref_w.shape)
        out[f"sample_major_{k}_eager_ms"] = timeit(lambda: fwd_sample_major(img, k))
        out[f"sample_major_{k}_graph_ms"] = timeit(graphed(lambda: fwd_sample_major(img, k)))
print(json.dumps(out, indent=1))
