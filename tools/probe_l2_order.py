"""Does running the encoder sample-by-sample (per-sample tensors fit the 126 MB L2) beat one batch-8 pass?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch
import bench

dev = torch.device("cuda")
G, E = bench.build_ours(dev)
with torch.no_grad():
    z = torch.randn(8, 512, device=dev)
    imgs = G(z, trunc_psi=0.7, trunc_layers=8)["image"].contiguous()

    def t(fn, it=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it

    for bs in (8, 4, 2, 1):
        chunks = [imgs[i:i + bs].contiguous() for i in range(0, 8, bs)]
        # capture in a graph to exclude launch overhead
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for c in chunks:
                E(c)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            for c in chunks:
                E(c)
        print(f"E forward, 8 images as {8 // bs} x batch {bs}: {t(g.replay):.3f} ms")
    w = E(imgs)[1]
    for bs in (8, 4, 2, 1):
        chunks = [w[i:i + bs].contiguous() for i in range(0, 8, bs)]
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for c in chunks:
                G.synthesis(c)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            for c in chunks:
                G.synthesis(c)
        print(f"G synthesis, 8 images as {8 // bs} x batch {bs}: {t(g.replay):.3f} ms")
