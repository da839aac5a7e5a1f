"""One E + G forward at the bench configuration after a warm-up pass (for ncu: `--launch-skip` past the warm-up)."""
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
    imgs1 = G(z, trunc_psi=0.7, trunc_layers=8)["image"]
    for it in range(2):
        c, w = E(imgs1)
        G.synthesis(w)
        torch.cuda.synchronize()
        print("pass", it, "done", flush=True)
