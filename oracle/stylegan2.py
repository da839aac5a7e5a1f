"""Oracle restatement of the StyleGAN2 generator forward (reference: model/stylegan2_generator.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure functions over a `state_dict`-style mapping
of fp32 tensors with the reference's key names.  The modulated convolution is
restated in the reference's own un-fused form (`fused_modulate=False`, :876-877, :908-909), i.e.
y = d[n,o] * conv(x * s[n,i], W * wscale), which SURVEY Appendix E-1 shows equals the grouped-conv path.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

_INIT_RES = 4
SQRT2 = math.sqrt(2.0)


def get_nf(res, fmaps_base=32 << 10, fmaps_max=512):
    """SynthesisModule.get_nf, model/stylegan2_generator.py:488-490."""
    return min(fmaps_base // res, fmaps_max)


def num_layers_for(resolution):
    """model/stylegan2_generator.py:126."""
    return int(np.log2(resolution // _INIT_RES * 2)) * 2


def dense_block(x, weight, bias, *, lr_mul=1.0, additional_bias=0.0, lrelu=True, use_wscale=True, gain=1.0):
    """DenseBlock.forward, model/stylegan2_generator.py:990-996 (+ ctor :956-975)."""
    in_c = weight.shape[1]
    wscale = (gain / math.sqrt(in_c)) * lr_mul if use_wscale else lr_mul
    if x.ndim != 2:
        x = x.view(x.shape[0], -1)
    b = bias * lr_mul if bias is not None else None
    y = F.linear(x, weight * wscale, b) + additional_bias
    if lrelu:
        y = F.leaky_relu(y, 0.2) * SQRT2
    return y


def pixel_norm(x, dim=1, eps=1e-8):
    """PixelNormLayer.forward, model/stylegan2_generator.py:550-553."""
    return x / torch.sqrt(torch.mean(x ** 2, dim=dim, keepdim=True) + eps)


def mapping(sd, z, *, num_layers=8, lr_mul=0.01, prefix="mapping."):
    """MappingModule.forward (label_size == 0), model/stylegan2_generator.py:246-278."""
    w = pixel_norm(z)
    for i in range(num_layers):
        w = dense_block(w, sd[f"{prefix}dense{i}.weight"], sd[f"{prefix}dense{i}.bias"], lr_mul=lr_mul)
    return w


def truncation(w, w_avg, num_layers, trunc_psi=None, trunc_layers=None):
    """TruncationModule.forward (repeat_w=True), model/stylegan2_generator.py:311-333."""
    if w.ndim == 2:
        wp = w.view(-1, 1, w.shape[1]).repeat(1, num_layers, 1)
    else:
        wp = w
    psi = 1.0 if trunc_psi is None else trunc_psi
    layers = 0 if trunc_layers is None else trunc_layers
    if psi < 1.0 and layers > 0:
        coefs = torch.ones(1, num_layers, 1, dtype=wp.dtype)
        coefs[:, :layers] *= psi
        wa = w_avg.view(1, -1, w_avg.shape[-1])
        wp = wa + (wp - wa) * coefs.to(wp.device)
    return wp


def fir_kernel(gain_sq=4.0):
    """UpsamplingLayer kernel: outer([1,3,3,1]) / sum * gain^2, model/stylegan2_generator.py:574-590."""
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = torch.outer(k, k)
    return (k / k.sum() * gain_sq)[None, None]


def filter_after_up(x):
    """ModulateConvBlock.filter = UpsamplingLayer(scale_factor=1, extra_padding=-1, kernel_gain=2):
    pad (1,1,1,1) then 4x4 FIR -- model/stylegan2_generator.py:799-807, 592-615."""
    n, c, h, w = x.shape
    x = F.pad(x.reshape(n * c, 1, h, w), (1, 1, 1, 1))
    x = F.conv2d(x, fir_kernel(4.0).to(x))
    return x.view(n, c, x.shape[2], x.shape[3])


def upsample_skip(img):
    """SynthesisModule.upsample = UpsamplingLayer(scale_factor=2): zero-insert, pad (2,1,2,1), 4x4 FIR
    -- model/stylegan2_generator.py:556-615."""
    n, c, h, w = img.shape
    x = img.view(n, c, h, 1, w, 1)
    x = F.pad(x, (0, 1, 0, 0, 0, 1, 0, 0))
    x = x.view(n * c, 1, 2 * h, 2 * w)
    x = F.pad(x, (2, 1, 2, 1))
    x = F.conv2d(x, fir_kernel(4.0).to(x))
    return x.view(n, c, 2 * h, 2 * w)


def modulate_conv_block(sd, prefix, x, w_latent, *, up=False, ksize=3, demodulate=True, add_noise=True,
                        lrelu=True, noise=None, eps=1e-8):
    """ModulateConvBlock.forward, model/stylegan2_generator.py:855-922.  Returns (x, style).

    `noise`: optional explicit [N|1,1,res,res] tensor (randomize_noise=True path); default = the
    registered buffer `<prefix>noise`.
    """
    weight = sd[prefix + "weight"]                      # [out, in, k, k]
    out_c, in_c = weight.shape[0], weight.shape[1]
    wscale = 1.0 / math.sqrt(ksize * ksize * in_c)      # :823-826, use_wscale, lr_mul = 1
    w = weight * wscale
    style = dense_block(w_latent, sd[prefix + "style.weight"], sd[prefix + "style.bias"], additional_bias=1.0,
                        lrelu=False)                    # :862, ctor :832-836
    n = x.shape[0]
    xs = x * style.view(n, in_c, 1, 1)                  # :877
    if up:
        # :879-896 -- flip, [in, out, kh, kw], stride-2 transposed conv with padding 0, then the FIR filter
        wt = w.flip(2, 3).permute(1, 0, 2, 3)
        y = F.conv_transpose2d(xs, wt, stride=2, padding=0)
        y = filter_after_up(y)
    else:
        y = F.conv2d(xs, w, padding=ksize // 2)         # :897-904
    if demodulate:
        d = torch.rsqrt((w.pow(2).sum(dim=(2, 3))[None] * style.pow(2)[:, None, :]).sum(dim=2) + eps)  # :867-870
        y = y * d.view(n, out_c, 1, 1)                  # :908-909
    if add_noise:
        nz = sd[prefix + "noise"] if noise is None else noise
        y = y + nz * sd[prefix + "noise_strength"].view(1, 1, 1, 1)   # :911-916
    y = y + sd[prefix + "bias"].view(1, -1, 1, 1)       # :918-920 (bscale = lr_mul = 1)
    if lrelu:
        y = F.leaky_relu(y, 0.2) * SQRT2                # :921
    return y, style


def synthesis(sd, wp, resolution, *, prefix="synthesis.", noises=None):
    """SynthesisModule.forward, architecture 'skip', model/stylegan2_generator.py:492-539.

    Returns the reference's result dict ('wp', 'styleNN', 'output_styleK', 'image').
    `noises`: optional dict layer_idx -> noise tensor (randomize_noise=True restated with explicit noise).
    """
    nl = num_layers_for(resolution)
    n = wp.shape[0]
    results = {"wp": wp}
    x = sd[prefix + "early_layer.const"].repeat(n, 1, 1, 1)        # InputBlock :630-632
    image = None
    for idx in range(nl - 1):
        nz = None if noises is None else noises.get(idx)
        x, style = modulate_conv_block(sd, f"{prefix}layer{idx}.", x, wp[:, idx], up=(idx % 2 == 1), noise=nz)
        results[f"style{idx:02d}"] = style
        if idx % 2 == 0:
            temp, style = modulate_conv_block(sd, f"{prefix}output{idx // 2}.", x, wp[:, idx + 1], ksize=1,
                                              demodulate=False, add_noise=False, lrelu=False)
            results[f"output_style{idx // 2}"] = style
            image = temp if idx == 0 else temp + upsample_skip(image)   # :519-522
    results["image"] = image                                           # final_activate = Identity (:475)
    return results


def generator(sd, z, resolution, *, trunc_psi=None, trunc_layers=None):
    """StyleGAN2Generator.forward in eval mode (no w_avg update / style mixing), :165-196."""
    n_map = sum(1 for k in sd if k.startswith("mapping.dense") and k.endswith(".weight"))
    w = mapping(sd, z, num_layers=n_map)
    wp = truncation(w, sd["truncation.w_avg"], num_layers_for(resolution), trunc_psi, trunc_layers)
    out = {"z": pixel_norm(z), "w": w}
    out.update(synthesis(sd, wp, resolution))
    return out
