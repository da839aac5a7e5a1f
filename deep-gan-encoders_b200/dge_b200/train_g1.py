"""Fused training path of the frozen StyleGAN1 generator (model/stylegan1/net.py:141-169, 331-336 under `loss.backward()`:
E_align_s2.py:158 with --mtype 1 and embedding_img.py:87 feed the encoder's styles through `Gs.forward(styles, lod)` and
back-propagate the image loss into the encoder -- d image / d styles is all the generator has to provide).

ONE autograd node for `Generator.decode`.  Forward = the inference chain of `DecodeBlock.run` (tcgen05 convs, `sg1_post`,
`instance_norm_style`).  Backward per block, from the last one, on kernels that already exist for the other paths:

    style_mod + instance norm   in_bwd_stats (sum g, sum g*xn = the gradients of the style shift / scale) and in_bwd_apply
                                (mode 1, `gscale` = s0 + 1): IN Jacobian * lrelu'  -> ACT operand (stage 2) / F32B (stage 1)
    conv_2                      dge_conv_forward on the data-gradient operand
    blur                        its own transpose (symmetric, zero padded): dge_sg1_post mode 0 without noise / bias
    x2 `transform_kernel` layer up_fir_bwd_s2d(box) (transpose of the 2x2 box sum, written space-to-depth) +
                                dge_conv_forward(DOWN4X4S2) = the stride-2 conv that inverts the transposed conv
    nearest x2 + conv (small resolutions)   data-gradient conv, then the 2x2 sum of the gradient (dge_blend with pooling)

and two [N, 2C] x [2C, 512] products per block for the style Linear layers.  `K` is the kernel namespace (see train_e.py).
"""
import torch
from torch.autograd.function import once_differentiable

from . import ops

K = ops
SLOPE = 0.2


def _lin(l):
    """(weight, bias) of an ln.Linear as the forward uses them."""
    w = l.weight.detach() if l.implicit_lreq else l.weight.detach() * l.std
    b = None if l.bias is None else (l.bias.detach() if l.implicit_lreq else l.bias.detach() * l.lrmul)
    return w, b


def _dgrad_ops(block, planes):
    """Cached packed operands of a (frozen) DecodeBlock -- forward ('f1', 'f2') and data-gradient ('c1', 'c2'): conv_2, and conv_1 as a plain 3x3 (nearest-up blocks) or as
    the 16-tap DOWN4X4S2 operand W4[i][o][ky+1][kx+1] = Wt[i][o][ky][kx] (transposed-conv blocks; Wt = conv_1.weight,
    [in, out, 3, 3], un-flipped as lreq.py:127-140 uses it)."""
    srcs = [block.conv_2.weight] + ([block.conv_1.weight] if block.has_first_conv else [])
    key = (K.weight_key(*srcs), planes, K is ops)
    hit = block.__dict__.get('_dge_dgrad_ops')
    if hit is not None and hit[0] == key:
        return hit[1]
    c2 = block.conv_2
    sc2 = 1.0 if c2.implicit_lreq else c2.std
    d = {'c2': K.pack_conv_weight_dgrad(c2.weight.detach(), scale=sc2, planes=planes),
         'f2': K.pack_conv_weight(c2.weight.detach(), scale=sc2, planes=planes)}
    if block.has_first_conv:
        c1 = block.conv_1
        sc = 1.0 if c1.implicit_lreq else c1.std
        if block.fused_scale:
            w4 = c1.weight.new_zeros((block.inputs, block.outputs, 4, 4))
            w4[:, :, 1:, 1:] = c1.weight.detach() * sc
            d['c1'] = K.pack_conv_weight(w4.contiguous(), planes=planes)
            # forward: [in, out, k, k] used un-flipped by conv_transpose2d (lreq.py:127-140)
            d['f1'] = K.pack_conv_weight(c1.weight.detach().permute(1, 0, 2, 3).contiguous(), scale=sc, planes=planes)
        else:
            d['c1'] = K.pack_conv_weight_dgrad(c1.weight.detach(), scale=sc, planes=planes)
            d['f1'] = K.pack_conv_weight(c1.weight.detach(), scale=sc, planes=planes)
    block.__dict__['_dge_dgrad_ops'] = (key, d)
    return d


class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, styles, G, lod):
        n = styles.shape[0]
        x = K.nchw_to_f32b(G.const.detach().float())
        saved = []
        for i in range(lod + 1):
            blk = G.decode_block[i]
            last = i == lod
            nxt = G.decode_block[i + 1] if not last else None
            next_up = 2 if (nxt is not None and not nxt.fused_scale) else 1
            s1, s2 = styles[:, 2 * i + 0], styles[:, 2 * i + 1]
            planes = blk.planes
            nw1, b1 = blk.noise_weight_1.detach().view(-1), blk.bias_1.detach().view(-1)
            nw2, b2 = blk.noise_weight_2.detach().view(-1), blk.bias_2.detach().view(-1)
            if not blk.has_first_conv:
                dev, h, w, nb = x.t.device, x.h, x.w, x.n
                y = K.sg1_post(x, 2, nb, blk.outputs, h, w, noise=blk._noise(nb, h, w, dev), noise_w=nw1, bias=b1)
            elif blk.fused_scale:
                dev, h, w, nb = x.t.device, 2 * x.h, 2 * x.w, x.n
                raw = K.conv(x, _dgrad_ops(blk, planes)['f1'], blk.outputs, K.CONV_UP3X3)['raw_up']
                y = K.sg1_post(raw, 1, nb, blk.outputs, h, w, noise=blk._noise(nb, h, w, dev), noise_w=nw1, bias=b1)
            else:
                dev, h, w, nb = x.t.device, x.h, x.w, x.n
                c1 = K.conv(x, _dgrad_ops(blk, planes)['f1'], blk.outputs, K.CONV_3X3, out_f32b=True)['f32b']
                y = K.sg1_post(c1, 0, nb, blk.outputs, h, w, noise=blk._noise(nb, h, w, dev), noise_w=nw1, bias=b1)
            _, mr1 = K.instance_stats(y, blk.instance_norm_1.eps)
            w_s1, b_s1 = _lin(blk.style_1)
            st1 = torch.nn.functional.linear(s1.float(), w_s1, b_s1)
            xa, _ = K.instance_norm_style(y, mr1, st1, n, planes=planes)
            y2 = K.conv(xa, _dgrad_ops(blk, planes)['f2'], blk.outputs, K.CONV_3X3, noise=blk._noise(n, h, w, dev),
                        noise_batched=True, noise_w=nw2, bias=b2, slope=SLOPE, out_f32b=True)['f32b']
            _, mr2 = K.instance_stats(y2, blk.instance_norm_2.eps)
            w_s2, b_s2 = _lin(blk.style_2)
            st2 = torch.nn.functional.linear(s2.float(), w_s2, b_s2)
            act, f = K.instance_norm_style(y2, mr2, st2, n, up=next_up, planes=planes, out_act=not last, out_f32b=last)
            saved.append((y, mr1, st1, y2, mr2, st2, next_up))
            x = f if last else act
        ctx.G, ctx.lod, ctx.saved, ctx.n, ctx.styles_shape = G, lod, saved, n, tuple(styles.shape)
        rgb = G.to_rgb[lod].to_rgb
        return K.to_rgb_f32b(x, rgb.weight.detach() if rgb.implicit_lreq else rgb.weight.detach() * rgb.std,
                             rgb.scaled_bias())

    @staticmethod
    @once_differentiable
    def backward(ctx, d_img):
        G, lod, n = ctx.G, ctx.lod, ctx.n
        d_styles = torch.zeros(ctx.styles_shape, dtype=torch.float32, device=d_img.device)
        rgb = G.to_rgb[lod].to_rgb
        w_rgb = (rgb.weight.detach() if rgb.implicit_lreq else rgb.weight.detach() * rgb.std).view(rgb.weight.shape[0], -1)
        g = K.nchw_to_f32b(torch.einsum('nihw,ic->nchw', d_img.float(), w_rgb).contiguous())     # ToRGB (:244-253)
        for i in range(lod, -1, -1):
            blk = G.decode_block[i]
            planes = blk.planes
            y, mr1, st1, y2, mr2, st2, next_up = ctx.saved[i]
            c = blk.outputs
            dops = _dgrad_ops(blk, planes)
            if i != lod and next_up == 2:               # the block's output was nearest-upsampled: sum the 2x2 replicas
                g = K.blend(g, g, 2.0, 2.0, pool=3)
            # stage 2: x*(s0+1)+s1 after IN(y2); y2 = lrelu(conv_2 + noise + bias)
            sums = K.in_bwd_stats(g, y2, mr2)
            d_st2 = torch.cat((sums[:, :, 1], sums[:, :, 0]), dim=1).float()             # d scale = sum g*xn, d shift = sum g
            d_styles[:, 2 * i + 1] = d_st2 @ _lin(blk.style_2)[0]
            dpre2, _ = K.in_bwd_apply(g, y2, mr2, None, None, sums, 1, slope=SLOPE, planes=planes,
                                      gscale=(st2[:, :c] + 1).contiguous())
            g = K.conv(dpre2, dops['c2'], c, K.CONV_3X3, out_f32b=True)['f32b']
            # stage 1: x*(s0+1)+s1 after IN(y); y = lrelu(blur(conv_1) + noise + bias)  (block 0: y has batch 1)
            if y.n == 1 and n > 1:
                xn = (y.to_nchw() - mr1[:, :, 0, None, None]) * mr1[:, :, 1, None, None]        # [1, C, 4, 4]
                gn = g.to_nchw()
                d_st1 = torch.cat(((gn * xn).sum(dim=(2, 3)), gn.sum(dim=(2, 3))), dim=1)
                d_styles[:, 2 * i] = d_st1 @ _lin(blk.style_1)[0]
                break                                      # the constant input is not trained
            sums = K.in_bwd_stats(g, y, mr1)
            d_st1 = torch.cat((sums[:, :, 1], sums[:, :, 0]), dim=1).float()
            d_styles[:, 2 * i] = d_st1 @ _lin(blk.style_1)[0]
            if not blk.has_first_conv:
                break
            dpost, _ = K.in_bwd_apply(g, y, mr1, None, None, sums, 1, slope=SLOPE, planes=planes,
                                      gscale=(st1[:, :c] + 1).contiguous(), out_kind="f32b")
            gb = K.sg1_post(dpost, 0, dpost.n, c, dpost.h, dpost.w, slope=1.0)                    # blur^T = blur
            if blk.fused_scale:
                s2d = K.up_fir_bwd_s2d(gb, planes, box=True)
                g = K.conv(s2d, dops['c1'], blk.inputs, K.CONV_DOWN4X4S2, out_f32b=True,
                           out_hw=(gb.h // 2, gb.w // 2))['f32b']
            else:
                g = K.conv(K.f32b_to_act(gb, planes), dops['c1'], blk.inputs, K.CONV_3X3, out_f32b=True)['f32b']
        return d_styles, None, None


def decode(G, styles, lod):
    """`Generator.decode` (:331-336) recorded for backward w.r.t. `styles` as one fused node."""
    return _DecodeFn.apply(styles.float(), G, lod)
