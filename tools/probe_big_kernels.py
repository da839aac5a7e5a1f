"""One fused training pass of the BigGAN pair (E_BIG(64, 7) + BigGAN-deep-256, BASELINE configs[3]) for ncu / the sanitizer:
G(z) -> E_BIG -> G(w2) -> image MSE -> backward, so that k_affine_relu_bwd and the re-used encoder backward kernels launch at
their real sizes.  usage (under ncu): python tools/probe_big_kernels.py [batch=16]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch

from model.biggan_generator import BigGAN
from model.E.E_BIG import BE
from model.utils.biggan_config import BigGANConfig

CFG = {"attention_layer_position": 8, "channel_width": 128, "class_embed_dim": 128, "eps": 0.0001,
       "layers": [[False, 16, 16], [True, 16, 16], [False, 16, 16], [True, 16, 8], [False, 8, 8], [True, 8, 8],
                  [False, 8, 8], [True, 8, 4], [False, 4, 4], [True, 4, 2], [False, 2, 2], [True, 2, 1]],
       "n_stats": 51, "num_classes": 1000, "output_dim": 256, "z_dim": 128}
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
G = BigGAN(BigGANConfig.from_dict(CFG)).eval()
with torch.no_grad():
    G.generator.bn.weight.fill_(1.0)
    G.generator.bn.bias.zero_()
G = G.cuda()
E = BE(64, 512, 7, 512, 3, biggan=True).cuda()
E.set_noise_mode("device")
z = torch.randn(batch, 128, device="cuda").clamp_(-2, 2) * 0.4
label = torch.zeros(batch, 1000, device="cuda")
label[:, 30] = 1
for it in range(2):          # pass 0 warms caches (packed weights); ncu skips it with --launch-skip
    with torch.no_grad():
        imgs1, const1 = G(z, label, 0.4)
    const2, w2 = E(imgs1, const1)
    imgs2, _ = G(w2, label, 0.4)
    ((imgs1 - imgs2) ** 2).mean().backward()
    torch.cuda.synchronize()
    print("pass", it, "done", flush=True)
