"""One fused training pass of the benchmark pair (BE(16,9) + StyleGAN2-1024, batch 8) for ncu: E forward -> G.synthesis
forward -> image MSE -> backward, so that every backward kernel of csrc/train_bwd.cu and the data / weight gradient convs
launch once at their real sizes.  usage (under ncu): python tools/probe_train_kernels.py [batch=8]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch

import model.E.E as EM
from model.stylegan2_generator import StyleGAN2Generator

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
E = EM.BE(16, 512, 9, 512, 3).cuda()
E.set_noise_mode("device")
G = StyleGAN2Generator(1024).cuda().eval()
z = torch.randn(batch, 512, device="cuda")
with torch.no_grad():
    imgs1 = G(z, trunc_psi=0.7, trunc_layers=8)["image"]
for it in range(2):          # pass 0 warms caches (packed weights, tensor maps); ncu skips it with --launch-skip
    const2, w2 = E(imgs1)
    imgs2 = G.synthesis(w2)["image"]
    ((imgs1 - imgs2) ** 2).mean().backward()
    torch.cuda.synchronize()
    print("pass", it, "done", flush=True)
