"""Oracle restatement of the PGGAN generator (reference model/pggan/pggan_generator.py) and the PGGAN encoder
(model/E/E_PG.py).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain PyTorch fp32 over state-dict tensors."""
import math

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


def pixel_norm(x, eps=1e-8):
    """PixelNormLayer.forward, pggan_generator.py:214-216."""
    return x / torch.sqrt(torch.mean(x ** 2, dim=1, keepdim=True) + eps)


def conv_block(sd, name, x, *, ksize=3, padding=1, upsample=False, gain=SQRT2, lrelu=True):
    """ConvBlock.forward (fused_scale=False), pggan_generator.py:319-339."""
    w = sd[name + ".weight"]
    wscale = gain / math.sqrt(ksize * ksize * w.shape[1])
    x = pixel_norm(x)
    if upsample:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(x, w * wscale, sd[name + ".bias"], stride=1, padding=padding)
    return F.leaky_relu(x, 0.2) if lrelu else x


def generator(sd, z, resolution, lod=0):
    """PGGANGenerator.forward, pggan_generator.py:154-204 (label_size = 0). Returns the image."""
    final_log2 = int(np.log2(resolution))
    z = pixel_norm(z)
    x = z.view(z.shape[0], -1, 1, 1)
    image = None
    for res_log2 in range(2, final_log2 + 1):
        cur = final_log2 - res_log2
        b = res_log2 - 2
        if lod < cur + 1:
            if res_log2 == 2:
                x = conv_block(sd, "layer0", x, ksize=4, padding=3)
            else:
                x = conv_block(sd, f"layer{2 * b}", x, upsample=True)
            x = conv_block(sd, f"layer{2 * b + 1}", x)
        if cur - 1 < lod <= cur:
            image = conv_block(sd, f"output{b}", x, ksize=1, padding=0, gain=1.0, lrelu=False)
        elif cur < lod < cur + 1:
            alpha = np.ceil(lod) - lod
            image = (conv_block(sd, f"output{b}", x, ksize=1, padding=0, gain=1.0, lrelu=False) * alpha +
                     F.interpolate(image, scale_factor=2, mode="nearest") * (1 - alpha))
        elif lod >= cur + 1:
            image = F.interpolate(image, scale_factor=2, mode="nearest")
    return image


def e_pg_block(sd, prefix, x, noise_fn=torch.randn):
    """E_PG.BEBlock.forward, model/E/E_PG.py:73-108."""
    residual = x
    x = F.instance_norm(x, eps=1e-8)
    x = F.conv2d(x, sd[prefix + "conv_1.weight"], padding=1)
    x = torch.addcmul(x, sd[prefix + "noise_weight_1"], noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x))
    x = F.leaky_relu(x + sd[prefix + "bias_1"], 0.2)
    if (prefix + "conv_2.weight") in sd:
        x = F.instance_norm(x, eps=1e-8)
        x = F.conv2d(x, sd[prefix + "conv_2.weight"], padding=1)
        x = torch.addcmul(x, sd[prefix + "noise_weight_2"], noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x))
        x = x + sd[prefix + "bias_2"]
        if (prefix + "conv_3.weight") in sd:
            residual = F.conv2d(residual, sd[prefix + "conv_3.weight"], sd[prefix + "conv_3.bias"])
            residual = F.instance_norm(residual, weight=sd[prefix + "instance_norm_3.weight"],
                                       bias=sd[prefix + "instance_norm_3.bias"], eps=1e-8)
        x = F.leaky_relu(x + residual, 0.2)
        x = F.avg_pool2d(x, 2, 2)
    return x


def e_pg_features(sd, x, layer_count, noise_fn=torch.randn):
    """E_PG.BE.forward up to the discarded `new_final` result, model/E/E_PG.py:150-163."""
    x = F.leaky_relu(F.conv2d(x, sd["FromRGB.from_rgb.weight"], sd["FromRGB.from_rgb.bias"]), 0.2)
    for i in range(layer_count):
        x = e_pg_block(sd, f"decode_block.{i}.", x, noise_fn)
    if "new_final.weight" in sd:
        x = F.linear(x.view(x.shape[0], -1), sd["new_final.weight"], sd["new_final.bias"])
    return x
