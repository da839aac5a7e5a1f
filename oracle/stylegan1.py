"""Oracle restatement of the StyleGAN1 generator hot path (reference model/stylegan1/net.py, lreq.py).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain PyTorch fp32 over state-dict tensors."""
import torch
import torch.nn.functional as F


def pixel_norm(x, eps=1e-8):
    """net.py:28-29."""
    return x * torch.rsqrt(torch.mean(x.pow(2.0), dim=1, keepdim=True) + eps)


def style_mod(x, style):
    """net.py:32-34."""
    style = style.view(style.shape[0], 2, x.shape[1], 1, 1)
    return torch.addcmul(style[:, 1], x, style[:, 0] + 1)


def blur(x):
    """Blur.forward, net.py:48-58."""
    c = x.shape[1]
    f = torch.tensor([1.0, 2.0, 1.0])
    k = (f[:, None] * f[None, :])
    k = (k / k.sum()).view(1, 1, 3, 3).repeat(c, 1, 1, 1)
    return F.conv2d(x, k, groups=c, padding=1)


def conv_transpose_fused(x, w):
    """ln.ConvTranspose2d(3, stride 2, pad 1, transform_kernel=True).forward, lreq.py:126-140 (implicit lreq)."""
    w = F.pad(w, (1, 1, 1, 1))
    w = w[:, :, 1:, 1:] + w[:, :, :-1, 1:] + w[:, :, 1:, :-1] + w[:, :, :-1, :-1]
    return F.conv_transpose2d(x, w, stride=2, padding=1)


def decode_block(sd, prefix, x, s1, s2, noise_fn=torch.randn, fused=None):
    """DecodeBlock.forward, net.py:141-169.  `fused`: the block's fused_scale flag (default: the Generator's rule)."""
    if (prefix + "conv_1.weight") in sd:
        w1 = sd[prefix + "conv_1.weight"]
        if fused is None:
            fused = 2 * x.shape[2] >= 128      # Generator: fused_scale = resolution*2 >= 128 (net.py:283)
        if fused:
            x = conv_transpose_fused(x, w1)
        else:
            x = F.conv2d(x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3), w1, padding=1)   # upscale2d :37-43
        x = blur(x)
    x = torch.addcmul(x, sd[prefix + "noise_weight_1"], noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x))
    x = F.leaky_relu(x + sd[prefix + "bias_1"], 0.2)
    x = F.instance_norm(x, eps=1e-8)
    x = style_mod(x, F.linear(s1, sd[prefix + "style_1.weight"], sd[prefix + "style_1.bias"]))
    x = F.conv2d(x, sd[prefix + "conv_2.weight"], padding=1)
    x = torch.addcmul(x, sd[prefix + "noise_weight_2"], noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x))
    x = F.leaky_relu(x + sd[prefix + "bias_2"], 0.2)
    x = F.instance_norm(x, eps=1e-8)
    return style_mod(x, F.linear(s2, sd[prefix + "style_2.weight"], sd[prefix + "style_2.bias"]))


def decode(sd, styles, lod, noise_fn=torch.randn):
    """Generator.decode, net.py:331-336."""
    x = sd["const"]
    for i in range(lod + 1):
        x = decode_block(sd, f"decode_block.{i}.", x, styles[:, 2 * i], styles[:, 2 * i + 1], noise_fn)
    return F.conv2d(x, sd[f"to_rgb.{lod}.to_rgb.weight"], sd[f"to_rgb.{lod}.to_rgb.bias"])


def mapping(sd, z, num_layers, mapping_layers=8, buffer1=None, coefs=0):
    """Mapping.forward, net.py:454-466."""
    x = pixel_norm(z)
    for i in range(mapping_layers):
        x = F.leaky_relu(F.linear(x, sd[f"block_{i + 1}.fc.weight"], sd[f"block_{i + 1}.fc.bias"]), 0.2)
    x = x.view(x.shape[0], 1, x.shape[1]).repeat(1, num_layers, 1)
    if buffer1 is not None:
        x = torch.lerp(buffer1, x, coefs)
    return x
