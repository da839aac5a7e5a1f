"""Per-kernel CUDA-event breakdown of one E+G step at the bench configuration (prints every launch class)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch
import bench
from dge_b200 import ops

dev = torch.device("cuda")
G, E = bench.build_ours(dev)
with torch.no_grad():
    z = torch.randn(8, 512, device=dev)
    imgs1 = G(z, trunc_psi=0.7, trunc_layers=8)["image"]
    for _ in range(3):
        c, w = E(imgs1)
        G.synthesis(w)
    torch.cuda.synchronize()
    for name, fn in (("E", lambda: E(imgs1)), ("G", lambda: G.synthesis(w))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms / forward")
        with ops.profile() as rec:
            for _ in range(3):
                fn()
        prof = rec.summary()
        rows = sorted(prof.items(), key=lambda kv: -kv[1][1])
        tot = sum(v[1] for v in prof.values()) / 3
        print(f"  sum of timed launches: {tot:.3f} ms")
        for (k, key), (cnt, ms) in rows:
            print(f"  {k:16s} {str(key):38s} x{cnt // 3:2d}  {ms / cnt:.3f} ms  total {ms / 3:.3f}")
