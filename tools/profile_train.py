"""Where the time of one E_align_s2.py training iteration goes (StyleGAN2-1024 + BE(16,9), batch 8): CUDA-event time of
each phase of the loop (:140-221) and a kernel-level table from torch.profiler.  Not a bench line.
usage: python tools/profile_train.py [batch=8] [iters=3] [lpips=1]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch

import model.E.E as EM
import training_utils as tu
from model.stylegan2_generator import StyleGAN2Generator
from model.utils.custom_adam import LREQAdam

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
use_lpips = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.manual_seed(0)
E = EM.BE(startf=16, maxf=512, layer_count=9).cuda()
E.set_noise_mode("device")
G = StyleGAN2Generator(resolution=1024).cuda().eval()
opt = LREQAdam(E.parameters(), lr=0.0015, betas=(0.0, 0.99), weight_decay=0)
if use_lpips:
    import lpips
    lp = lpips.LPIPS(net="vgg", pretrained=False, pnet_rand=True, verbose=False).cuda()
else:
    lp = lambda a, b: ((a - b) ** 2).mean(dim=(1, 2, 3), keepdim=True)
z = torch.randn(batch, 512, device="cuda")

phases = {}


class phase:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e0.record()

    def __exit__(self, *exc):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        phases.setdefault(self.name, []).append((self.e0, e1))


def crop(x, m):
    return x[:, :, m:-m, m:-m] if m else x


def iteration():
    with phase("G(z) no_grad"):
        with torch.no_grad():
            r1 = G(z, trunc_psi=0.7, trunc_layers=8)
        imgs1, w1 = r1["image"], r1["wp"]
    with phase("E fwd"):
        const2, w2 = E(imgs1)
    with phase("G.synthesis fwd"):
        imgs2 = G.synthesis(w2)["image"]
    with phase("losses fwd"):
        l0, _ = tu.space_loss(imgs1, imgs2, lpips_model=lp)
        m = imgs1.shape[3] // 8
        l1, _ = tu.space_loss(imgs1[:, :, :, m:-m], imgs2[:, :, :, m:-m], lpips_model=lp)
        m2 = m + imgs1.shape[2] // 32
        l2, _ = tu.space_loss(crop(imgs1, m2), crop(imgs2, m2), lpips_model=lp)
    with phase("backward 1"):
        opt.zero_grad()
        (l0 + 5 * l1 + 9 * l2).backward(retain_graph=True)
    with phase("step 1"):
        opt.step()
    with phase("latent loss"):
        lw, _ = tu.space_loss(w1, w2, image_space=False)
    with phase("backward 2"):
        opt.zero_grad()
        (lw * 0.01).backward()
    with phase("step 2"):
        opt.step()


for _ in range(2):
    iteration()
torch.cuda.synchronize()
phases.clear()
import time
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    iteration()
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / iters * 1e3
out = {"batch": batch, "lpips": bool(use_lpips), "ms_per_iteration": e0.elapsed_time(e1) / iters, "wall_ms": wall,
       "peak_gib": torch.cuda.max_memory_allocated() / 2 ** 30,
       "phases_ms": {k: sum(a.elapsed_time(b) for a, b in v) / iters for k, v in phases.items()}}
print(json.dumps(out, indent=1))

from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    iteration()
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = sys.argv[4] if len(sys.argv) > 4 else "r2"
with open(os.path.join(ROOT, "gpurun_out", f"{tag}_train_profile.txt"), "w") as f:
    f.write(json.dumps(out, indent=1) + "\n" + tab)
print(tab[-6000:])
