"""Oracle restatement of the case-1 encoder forward (reference: model/E/E.py, model/utils/net.py,
model/utils/lreq.py).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Pure functions over a state_dict-style mapping with the reference's key names.
Noise: the reference draws `torch.randn([N,1,H,W])` on the CPU inside every conv stage (E.py:60,73).
`noise_fn(shape)` supplies those draws; the default consumes the global CPU generator in the same
order as the reference, so `torch.manual_seed(s); be_forward(...)` reproduces `torch.manual_seed(s); E(x)`.
"""
import torch
import torch.nn.functional as F


def _default_noise(shape):
    return torch.randn(shape)


def _stats(x):
    """mean / biased std over (H,W), no eps -- E.py:51-52, 64-65."""
    mean = torch.mean(x, dim=[2, 3], keepdim=True)
    std = torch.sqrt(torch.mean((x - mean) ** 2, dim=[2, 3], keepdim=True))
    return torch.cat((mean, std), dim=1).view(x.shape[0], -1)


def from_rgb(sd, x, prefix="FromRGB.from_rgb."):
    """FromRGB.forward: 1x1 ln.Conv2d (+bias) then leaky_relu(0.2) -- model/utils/net.py:231-240;
    implicit lreq => raw weights (lreq.py:155-156)."""
    return F.leaky_relu(F.conv2d(x, sd[prefix + "weight"], sd[prefix + "bias"]), 0.2)


def be_block(sd, prefix, x, noise_fn=_default_noise):
    """BEBlock.forward, model/E/E.py:50-85 (fused_scale hard-wired False, :106). Returns (x, w1, w2)."""
    has_last_conv = (prefix + "conv_2.weight") in sd
    w1 = F.linear(_stats(x), sd[prefix + "inver_mod1.weight"], sd[prefix + "inver_mod1.bias"])        # :51-54
    residual = x
    x = F.instance_norm(x, eps=1e-8)                                                                  # :58
    x = F.conv2d(x, sd[prefix + "conv_1.weight"], padding=1)                                          # :59
    nz = noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x)
    x = torch.addcmul(x, sd[prefix + "noise_weight_1"], nz)                                           # :60
    x = F.leaky_relu(x + sd[prefix + "bias_1"], 0.2)                                                  # :61-62
    w2 = F.linear(_stats(x), sd[prefix + "inver_mod2.weight"], sd[prefix + "inver_mod2.bias"])        # :64-67
    x = F.instance_norm(x, eps=1e-8)                                                                  # :69
    if has_last_conv:
        x = F.conv2d(x, sd[prefix + "conv_2.weight"], padding=1)                                      # :72
        nz = noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x)
        x = torch.addcmul(x, sd[prefix + "noise_weight_2"], nz)                                       # :73
        x = F.leaky_relu(x + sd[prefix + "bias_2"], 0.2)                                              # :74-75
        x = F.avg_pool2d(x, 2, 2)                                                                     # :76-77
        residual = F.avg_pool2d(residual, 2, 2)                                                       # :78
    if (prefix + "conv_3.weight") in sd:
        residual = F.conv2d(residual, sd[prefix + "conv_3.weight"], sd[prefix + "conv_3.bias"])       # :81-82
    return 0.111 * x + 0.889 * residual, w1, w2                                                       # :84


def be_forward(sd, x, layer_count, block_num=9, noise_fn=_default_noise):
    """BE.forward, model/E/E.py:122-135.  Returns (const [N,C,4,4], w [N, 2*blocks, 512])."""
    x = from_rgb(sd, x)
    w = None
    for i in range(9 - block_num, layer_count):
        x, w1, w2 = be_block(sd, f"decode_block.{i}.", x, noise_fn)
        w_ = torch.cat((w2.view(x.shape[0], 1, -1), w1.view(x.shape[0], 1, -1)), dim=1)
        w = w_ if w is None else torch.cat((w_, w), dim=1)
    return x, w


def blur3x3(x):
    """Blur.forward, model/utils/net.py:45-55."""
    c = x.shape[1]
    f = torch.tensor([1.0, 2.0, 1.0])
    k = f[:, None] * f[None, :]
    k = (k / k.sum()).view(1, 1, 3, 3).repeat(c, 1, 1, 1)
    return F.conv2d(x, k, groups=c, padding=1)


def strided_transform_conv(x, w):
    """ln.Conv2d(3, stride 2, pad 1, transform_kernel=True).forward, model/utils/lreq.py:144-156 (implicit lreq)."""
    w = F.pad(w, (1, 1, 1, 1))
    w = (w[:, :, 1:, 1:] + w[:, :, :-1, 1:] + w[:, :, 1:, :-1] + w[:, :, :-1, :-1]) * 0.25
    return F.conv2d(x, w, stride=2, padding=1)


def be_blur_block(sd, prefix, x, fused_scale, noise_fn=_default_noise):
    """E_Blur.BEBlock.forward, model/E/E_Blur.py:50-85."""
    has_last_conv = (prefix + "conv_2.weight") in sd
    w1 = F.linear(_stats(x), sd[prefix + "inver_mod1.weight"], sd[prefix + "inver_mod1.bias"])
    residual = x
    x = F.instance_norm(x, eps=1e-8)
    x = F.conv2d(x, sd[prefix + "conv_1.weight"], padding=1)
    x = torch.addcmul(x, sd[prefix + "noise_weight_1"], noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x))
    x = F.leaky_relu(x + sd[prefix + "bias_1"], 0.2)
    w2 = F.linear(_stats(x), sd[prefix + "inver_mod2.weight"], sd[prefix + "inver_mod2.bias"])
    x = F.instance_norm(x, eps=1e-8)
    if has_last_conv:
        x = blur3x3(x)                                                                                # :71
        if fused_scale:
            x = strided_transform_conv(x, sd[prefix + "conv_2.weight"])                               # :72
        else:
            x = F.conv2d(x, sd[prefix + "conv_2.weight"], padding=1)
        x = torch.addcmul(x, sd[prefix + "noise_weight_2"], noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x))
        x = F.leaky_relu(x + sd[prefix + "bias_2"], 0.2)
        if not fused_scale:
            x = F.avg_pool2d(x, 2, 2)
        residual = F.avg_pool2d(residual, 2, 2)
    if (prefix + "conv_3.weight") in sd:
        residual = F.conv2d(residual, sd[prefix + "conv_3.weight"], sd[prefix + "conv_3.bias"])
    return 0.111 * x + 0.889 * residual, w1, w2


def be_blur_forward(sd, x, layer_count, noise_fn=_default_noise):
    """E_Blur.BE.forward, model/E/E_Blur.py:120-134; fused_scale = (1024 / 2^i >= 128), i.e. blocks 0..3 (:99,105)."""
    x = from_rgb(sd, x)
    w = None
    for i in range(layer_count):
        x, w1, w2 = be_blur_block(sd, f"decode_block.{i}.", x, 1024 / 2 ** i >= 128, noise_fn)
        w_ = torch.cat((w2.view(x.shape[0], 1, -1), w1.view(x.shape[0], 1, -1)), dim=1)
        w = w_ if w is None else torch.cat((w_, w), dim=1)
    return x, w
