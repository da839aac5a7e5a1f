"""Encoder training iteration at the benchmark shape (BE(startf=16, layer_count=9), 1024^2, batch 8): forward +
backward + LREQAdam.step through the mirrored module, with the convs on the tcgen05 kernels (product training path)
and, for context, with the same graph's convs on ATen/cuDNN fp32 (TF32 off).  Not a bench line.
usage: python tools/probe_train.py [batch=8] [iters=3] [full]
`full` adds the whole E_align_s2.py iteration: imgs1 = G(z) (no_grad), E(imgs1), G.synthesis(w2), the three image-space
`space_loss` calls (stand-in LPIPS) + the latent one, two backward / LREQAdam steps (:152-233)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import model.E.E as EM
from model.utils.custom_adam import LREQAdam

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
E = EM.BE(startf=16, maxf=512, layer_count=9).cuda()
E.set_noise_mode("device")
opt = LREQAdam(E.parameters(), lr=0.0015, betas=(0.0, 0.99), weight_decay=0)
img = torch.randn(batch, 3, 1024, 1024, device="cuda")
t_w = torch.randn(batch, 18, 512, device="cuda")


def iteration():
    const, w = E(img)
    loss = ((w - t_w) ** 2).mean() + (const ** 2).mean()
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


def timed():
    for _ in range(2):
        iteration()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        iteration()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms_tc = timed()
grads_tc = {k: p.grad.clone() for k, p in E.named_parameters() if p.grad is not None}
peak = torch.cuda.max_memory_allocated() / 2 ** 30
tc_conv = EM.tc.conv2d
EM.tc.conv2d = lambda x, w, planes=2: F.conv2d(x, w, padding=w.shape[-1] // 2)
ms_aten = timed()
EM.tc.conv2d = tc_conv
out = {"batch": batch, "ms_per_iteration_tcgen05_convs": ms_tc, "ms_per_iteration_aten_fp32_convs": ms_aten,
       "images_per_s_tcgen05_convs": batch / ms_tc * 1e3, "images_per_s_aten_fp32_convs": batch / ms_aten * 1e3,
       "peak_gib": peak}
if len(sys.argv) > 3 and sys.argv[3] == "full":
    import training_utils as tu
    from model.stylegan2_generator import StyleGAN2Generator
    G = StyleGAN2Generator(resolution=1024).cuda().eval()
    lp = lambda a, b: ((a - b) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    z = torch.randn(batch, 512, device="cuda")

    def crop(x, m):
        return x[:, :, m:-m, m:-m] if m else x

    def full_iteration():
        with torch.no_grad():
            r1 = G(z, trunc_psi=0.7, trunc_layers=8)
        imgs1, w1 = r1["image"], r1["wp"]
        const2, w2 = E(imgs1)
        imgs2 = G.synthesis(w2)["image"]
        l0, _ = tu.space_loss(imgs1, imgs2, lpips_model=lp)
        m = imgs1.shape[3] // 8
        l1, _ = tu.space_loss(imgs1[:, :, :, m:-m], imgs2[:, :, :, m:-m], lpips_model=lp)
        m2 = m + imgs1.shape[2] // 32
        l2, _ = tu.space_loss(crop(imgs1, m2), crop(imgs2, m2), lpips_model=lp)
        opt.zero_grad()
        (l0 + 5 * l1 + 9 * l2).backward(retain_graph=True)
        opt.step()
        lw, _ = tu.space_loss(w1, w2, image_space=False)
        opt.zero_grad()
        lw.backward()
        opt.step()

    iteration = full_iteration
    torch.cuda.reset_peak_memory_stats()
    ms_full = timed()
    out.update({"ms_per_full_iteration": ms_full, "images_per_s_full_iteration": batch / ms_full * 1e3,
                "peak_gib_full": torch.cuda.max_memory_allocated() / 2 ** 30})
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "train_probe.json"), "w"), indent=1)
