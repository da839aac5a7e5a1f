"""Oracle restatement of the BigGAN-deep generator (reference model/biggan_generator.py) and the BigGAN encoder
(model/E/E_BIG.py), eval mode.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain PyTorch fp32."""
import math

import torch
import torch.nn.functional as F


def sn_weight(sd, prefix):
    """torch.nn.utils.spectral_norm in eval mode: W_orig / (u . (W_mat v)), no power iteration."""
    w = sd[prefix + "weight_orig"]
    sigma = torch.dot(sd[prefix + "weight_u"], torch.mv(w.reshape(w.shape[0], -1), sd[prefix + "weight_v"]))
    return w / sigma


def bn_stats(sd, prefix, truncation, n_stats=51):
    """BigGANBatchNorm statistics lookup, biggan_generator.py:129-136."""
    step = 1.0 / (n_stats - 1)
    coef, idx = math.modf(truncation / step)
    idx = int(idx)
    means, vars_ = sd[prefix + "running_means"], sd[prefix + "running_vars"]
    if coef != 0.0:
        return means[idx] * coef + means[idx + 1] * (1 - coef), vars_[idx] * coef + vars_[idx + 1] * (1 - coef)
    return means[idx], vars_[idx]


def cbn(sd, prefix, x, truncation, cond, eps):
    """BigGANBatchNorm.forward (conditional), :138-148."""
    mean, var = bn_stats(sd, prefix, truncation, sd[prefix + "running_means"].shape[0])
    weight = 1 + F.linear(cond, sn_weight(sd, prefix + "scale."))[:, :, None, None]
    bias = F.linear(cond, sn_weight(sd, prefix + "offset."))[:, :, None, None]
    return (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps) * weight + bias


def gen_block(sd, prefix, x, cond, truncation, up, eps):
    """GenBlock.forward, :175-203."""
    x0 = x
    x = F.conv2d(F.relu(cbn(sd, prefix + "bn_0.", x, truncation, cond, eps)), sn_weight(sd, prefix + "conv_0."),
                 sd[prefix + "conv_0.bias"])
    x = F.relu(cbn(sd, prefix + "bn_1.", x, truncation, cond, eps))
    if up:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(x, sn_weight(sd, prefix + "conv_1."), sd[prefix + "conv_1.bias"], padding=1)
    x = F.conv2d(F.relu(cbn(sd, prefix + "bn_2.", x, truncation, cond, eps)), sn_weight(sd, prefix + "conv_2."),
                 sd[prefix + "conv_2.bias"], padding=1)
    x = F.conv2d(F.relu(cbn(sd, prefix + "bn_3.", x, truncation, cond, eps)), sn_weight(sd, prefix + "conv_3."),
                 sd[prefix + "conv_3.bias"])
    out_c = x.shape[1]
    if x0.shape[1] != out_c:
        x0 = x0[:, :x0.shape[1] // 2]
    if up:
        x0 = F.interpolate(x0, scale_factor=2, mode="nearest")
    return x + x0


def self_attn(sd, prefix, x):
    """SelfAttn.forward, :75-97."""
    _, ch, h, w = x.shape
    theta = F.conv2d(x, sn_weight(sd, prefix + "snconv1x1_theta.")).view(-1, ch // 8, h * w)
    phi = F.max_pool2d(F.conv2d(x, sn_weight(sd, prefix + "snconv1x1_phi.")), 2, 2).view(-1, ch // 8, h * w // 4)
    attn = torch.softmax(torch.bmm(theta.permute(0, 2, 1), phi), dim=-1)
    g = F.max_pool2d(F.conv2d(x, sn_weight(sd, prefix + "snconv1x1_g.")), 2, 2).view(-1, ch // 2, h * w // 4)
    attn_g = torch.bmm(g, attn.permute(0, 2, 1)).view(-1, ch // 2, h, w)
    return x + sd[prefix + "gamma"] * F.conv2d(attn_g, sn_weight(sd, prefix + "snconv1x1_o_conv."))


def biggan(sd, cfg, z, class_label, truncation):
    """BigGAN.forward + Generator.forward, :232-256, 296-304.  cfg: dict with channel_width, layers,
    attention_layer_position, eps."""
    ch, eps = cfg["channel_width"], cfg["eps"]
    cond = torch.cat((z, F.linear(class_label, sd["embeddings.weight"])), dim=1)
    p = "generator."
    x = F.linear(cond, sn_weight(sd, p + "gen_z."), sd[p + "gen_z.bias"])
    x = x.view(-1, 4, 4, 16 * ch).permute(0, 3, 1, 2).contiguous()
    j = 0
    for i, (up, _, _) in enumerate(cfg["layers"]):
        if i == cfg["attention_layer_position"]:
            x = self_attn(sd, f"{p}layers.{j}.", x)
            j += 1
        x = gen_block(sd, f"{p}layers.{j}.", x, cond, truncation, up, eps)
        j += 1
    mean, var = bn_stats(sd, p + "bn.", truncation, sd[p + "bn.running_means"].shape[0])
    x = F.batch_norm(x, mean, var, sd[p + "bn.weight"], sd[p + "bn.bias"], training=False, momentum=0.0, eps=eps)
    x = F.conv2d(F.relu(x), sn_weight(sd, p + "conv_to_rgb."), sd[p + "conv_to_rgb.bias"], padding=1)
    return torch.tanh(x[:, :3]), cond


def e_big_block(sd, prefix, x, cond, truncation=0.4, noise_fn=torch.randn):
    """E_BIG.BEBlock.forward, model/E/E_BIG.py:129-169 (BN eps 1e-12)."""
    eps = 1e-12
    residual = x
    x = cbn(sd, prefix + "batch_norm_1.", x, truncation, cond, eps)
    x = F.conv2d(x, sd[prefix + "conv_1.weight"], padding=1)
    x = torch.addcmul(x, sd[prefix + "noise_weight_1"], noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x))
    x = F.leaky_relu(x + sd[prefix + "bias_1"], 0.2)
    if (prefix + "conv_2.weight") in sd:
        x = cbn(sd, prefix + "batch_norm_2.", x, truncation, cond, eps)
        x = F.conv2d(x, sd[prefix + "conv_2.weight"], padding=1)
        x = torch.addcmul(x, sd[prefix + "noise_weight_2"], noise_fn([x.shape[0], 1, x.shape[2], x.shape[3]]).to(x))
        x = F.leaky_relu(x + sd[prefix + "bias_2"], 0.2)
        if (prefix + "conv_3.weight") in sd:
            residual = cbn(sd, prefix + "batch_norm_3.", residual, truncation, cond, eps)
            residual = F.conv2d(residual, sd[prefix + "conv_3.weight"], sd[prefix + "conv_3.bias"])
            x = F.leaky_relu(x, 0.2)
        x = F.avg_pool2d(x + residual, 2, 2)
    return x


def e_big_features(sd, x, cond, layer_count, noise_fn=torch.randn):
    """E_BIG.BE.forward up to the heads, :212-222."""
    x = F.leaky_relu(F.conv2d(x, sd["FromRGB.from_rgb.weight"], sd["FromRGB.from_rgb.bias"]), 0.2)
    for i in range(layer_count):
        x = e_big_block(sd, f"decode_block.{i}.", x, cond, 0.4, noise_fn)
    return x


def e_big_forward(sd, x, cond, layer_count, noise_fn=torch.randn):
    """E_BIG.BE.forward, :212-227 (biggan=True)."""
    x = e_big_features(sd, x, cond, layer_count, noise_fn)
    c_v = F.linear(x.view(x.shape[0], -1), sd["new_final_1.weight"], sd["new_final_1.bias"])
    return c_v, F.linear(c_v, sd["new_final_2.weight"], sd["new_final_2.bias"])
