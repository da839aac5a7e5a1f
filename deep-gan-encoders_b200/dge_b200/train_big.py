"""Fused training path of the BigGAN pair (E_align_s2.py mtype 4: `E_BIG.BE` trained against the frozen BigGAN-deep
generator; model/E/E_BIG.py:94-227, model/biggan_generator.py:153-256 under `loss.backward()`).

The conditional batch norms of both networks are per-(sample, channel) affines `a*x + b` whose coefficients come from the
condition vector through two small spectral-norm Linear layers.  Those Linear layers (and the spectral-norm power
iteration of the trainable encoder) stay torch ops on [N, C] tensors; everything that touches a feature map is one
autograd node per block whose forward is the inference kernel chain and whose backward is dge_b200 kernels only:

  E_BIG block   be_head_bwd -> conv_wgrad / data-gradient conv -> in_bwd_stats (= d a, d b) + in_bwd_apply with
                (mean, rstd) := (0, a) and zero statistics, i.e. dx = a * g with the leaky-ReLU mask, the bias /
                noise-weight reductions, the pooled residual gradient fused exactly as in the case-1 encoder (train_e.py)
  GenBlock      4 x [data-gradient conv -> affine_relu_bwd (ReLU mask from a*x + b, d a / d b sums, nearest-x2 transpose,
                the channel-drop identity branch)]; no weight gradients: the generator is frozen
  RGB tail      tanh' on the 3 image channels -> data-gradient conv -> affine_relu_bwd

`K` is the kernel namespace (see train_e.py): tests swap it for the torch emulation to check the chain rule on the CPU.
"""
import torch
from torch.autograd.function import once_differentiable

from . import ops
from . import train_e
from .train_e import _f32b, _packed

K = ops

SLOPE = 0.2                    # E_BIG.py:138,154


def _mr(a, zero_rstd=False):
    """(mean, rstd) table [N, C, 2] that makes the instance-norm kernels compute the affine's linear part: mean 0, rstd a
    (or rstd 1: plain sums of g and g*x)."""
    r = torch.ones_like(a) if zero_rstd else a
    return torch.stack((torch.zeros_like(a), r), dim=2).contiguous()


def _ab_grads(st):
    """in_bwd_stats output [N, C, 2] = (sum g, sum g*x) -> (d a, d b) as fp32 [N, C]."""
    return st[:, :, 1].float(), st[:, :, 0].float()


class _EBigBlockFn(torch.autograd.Function):
    """One E_BIG.BEBlock (E_BIG.py:129-169) on F32B tensors; a_i, b_i [N, C] are the three block norms' affines."""

    @staticmethod
    def forward(ctx, x_t, a1, b1, a2, b2, a3, b3, w1, w2, w3, b3c, nw1, bias1, nw2, bias2, cfg):
        has_second, planes, noise1, noise2 = cfg
        x = _f32b(x_t)
        n, c, h, w = x.n, x.c, x.h, x.w
        a1, b1 = a1.detach().contiguous(), b1.detach().contiguous()
        xn, _ = K.affine_act(x, a1, b1, relu=False, planes=planes)                                       # :134
        y1 = K.conv(xn, _packed(w1, planes), c, K.CONV_3X3, noise=noise1, noise_batched=True,
                    noise_w=nw1.detach().reshape(-1), bias=bias1.detach().reshape(-1), slope=SLOPE,
                    out_f32b=True)['f32b']                                                               # :135-138
        ctx.cfg = (has_second, planes, (n, c, h, w), w3 is not None)
        if not has_second:
            ctx.save_for_backward(x_t, a1, xn.t, y1.t, noise1, w1)
            return y1.t
        a2, b2 = a2.detach().contiguous(), b2.detach().contiguous()
        y1n, _ = K.affine_act(y1, a2, b2, relu=False, planes=planes)                                     # :150
        cout = w2.shape[0]
        if w3 is not None:
            a3, b3 = a3.detach().contiguous(), b3.detach().contiguous()
            xp = K.blend(x, x, 1.0, 0.0, pool=3)            # the 1x1 residual conv commutes with the 2x2 mean (:160-166)
            rp, _ = K.affine_act(xp, a3, b3, relu=False, planes=planes)
            slope2 = SLOPE * SLOPE                          # leaky-ReLU applied twice (:154,163)
        else:
            slope2 = SLOPE
        y2 = K.conv(y1n, _packed(w2, planes), cout, K.CONV_3X3, noise=noise2, noise_batched=True,
                    noise_w=nw2.detach().reshape(-1), bias=bias2.detach().reshape(-1), slope=slope2,
                    out_f32b=True)['f32b']                                                               # :151-154
        if w3 is not None:
            out = K.conv(rp, _packed(w3, planes), cout, K.CONV_1X1, bias=b3c.detach(), blend_src=y2, blend_pool=True,
                         blend_a=1.0, blend_b=1.0, out_f32b=True)['f32b']                                # :160-166
            ctx.save_for_backward(x_t, a1, xn.t, y1.t, noise1, w1, a2, y1n.t, y2.t, noise2, w2, a3, xp.t, rp.t, w3)
        else:
            out = K.blend(y2, x, 1.0, 1.0, pool=3)
            ctx.save_for_backward(x_t, a1, xn.t, y1.t, noise1, w1, a2, y1n.t, y2.t, noise2, w2)
        ctx.slope2 = slope2
        return out.t

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out_t):
        has_second, planes, (n, c, h, w), has_w3 = ctx.cfg
        sv = list(ctx.saved_tensors)
        x_t, a1, xn_t, y1_t, noise1, w1 = sv[:6]
        x, y1 = _f32b(x_t), _f32b(y1_t)
        xn = K.Act.wrap(xn_t, n, c, h, w, planes)
        d_out = _f32b(d_out_t.contiguous())
        dev = d_out_t.device
        zeros = torch.zeros((n, c, 2), dtype=torch.float64, device=dev)
        one = _mr(a1, zero_rstd=True)
        da2 = db2 = da3 = db3 = dw2 = dw3 = db3c = dnw2 = dbias2 = None
        if has_second:
            a2, y1n_t, y2_t, noise2, w2 = sv[6:11]
            y2 = _f32b(y2_t)
            dy2, dres, s = K.be_head_bwd(d_out, y2, noise2, 1.0, 1.0, ctx.slope2, want_dres=has_w3, planes=planes)
            dbias2, dnw2 = s[0].view(1, -1, 1, 1), s[1].view(1, -1, 1, 1)
            if has_w3:
                db3c = s[2]
            dw2 = K.conv_wgrad(dy2, K.Act.wrap(y1n_t, n, c, h, w, planes), 3)
            g1 = K.conv(dy2, _packed(w2, planes, True), c, K.CONV_3X3, out_f32b=True)['f32b']          # d (a2*y1 + b2)
            del dy2
            da2, db2 = _ab_grads(K.in_bwd_stats(g1, y1, one))
            dy1, s = K.in_bwd_apply(g1, y1, _mr(a2), None, None, zeros, 1, noise=noise1, slope=SLOPE, planes=planes)
            del g1
        else:
            dy1, s = K.in_bwd_apply(d_out, y1, one, None, None, zeros, 1, noise=noise1, slope=SLOPE, planes=planes)
        dbias1, dnw1 = s[0].view(1, -1, 1, 1), s[1].view(1, -1, 1, 1)
        dw1 = K.conv_wgrad(dy1, xn, 3)
        g0 = K.conv(dy1, _packed(w1, planes, True), c, K.CONV_3X3, out_f32b=True)['f32b']              # d (a1*x + b1)
        del dy1
        da1, db1 = _ab_grads(K.in_bwd_stats(g0, x, one))
        if not has_second:
            d_x = K.in_bwd_apply(g0, x, _mr(a1), None, None, zeros, 0)
        elif has_w3:
            a3, xp_t, rp_t, w3 = sv[11:15]
            xp = _f32b(xp_t)
            dw3 = K.conv_wgrad(dres, K.Act.wrap(rp_t, n, c, h // 2, w // 2, planes), 1)
            d_rp = K.conv(dres, _packed(w3, planes, True), c, K.CONV_1X1, out_f32b=True)['f32b']       # d (a3*pool(x) + b3)
            da3, db3 = _ab_grads(K.in_bwd_stats(d_rp, xp, one))
            d_xp = K.in_bwd_apply(d_rp, xp, _mr(a3), None, None, zeros, 0)                              # a3 * d_rp
            d_x = K.in_bwd_apply(g0, x, _mr(a1), None, None, zeros, 0, res=d_xp, rscale=0.25, res_pool=True)
        else:
            d_x = K.in_bwd_apply(g0, x, _mr(a1), None, None, zeros, 0, res=d_out, rscale=0.25, res_pool=True)
        return (d_x.t, da1, db1, da2, db2, da3, db3, dw1, dw2, dw3, db3c, dnw1, dbias1, dnw2, dbias2, None)


def ebig_block_forward(block, x_t, cond_vector, truncation=0.4):
    """E_BIG.BEBlock under autograd on an F32B tensor (E_BIG.py:129-169)."""
    n, _, h, w, _ = x_t.shape
    dev = x_t.device
    noise1 = block._noise(n, h, w, dev).reshape(n, h, w).contiguous()                                   # :136 (RNG order kept)
    noise2 = block._noise(n, h, w, dev).reshape(n, h, w).contiguous() if block.has_second_conv else None   # :152
    has_w3 = block.has_second_conv and block.inputs != block.outputs
    a1, b1 = block.batch_norm_1.coeffs_autograd(truncation, cond_vector, n, frozen=False)
    a2 = b2 = a3 = b3 = None
    if block.has_second_conv:
        a2, b2 = block.batch_norm_2.coeffs_autograd(truncation, cond_vector, n, frozen=False)
    if has_w3:
        a3, b3 = block.batch_norm_3.coeffs_autograd(truncation, cond_vector, n, frozen=False)
    cfg = (block.has_second_conv, block.planes, noise1, noise2)
    return _EBigBlockFn.apply(
        x_t, a1, b1, a2, b2, a3, b3, block.conv_1.weight, block.conv_2.weight if block.has_second_conv else None,
        block.conv_3.weight if has_w3 else None, block.conv_3.bias if has_w3 else None, block.noise_weight_1,
        block.bias_1, block.noise_weight_2 if block.has_second_conv else None,
        block.bias_2 if block.has_second_conv else None, cfg)


def ebig_features(E, x, cond_vector, block_num=9):
    """`E_BIG.BE` up to the [N, C, 4, 4] feature map (E_BIG.py:212-221), one fused node per block."""
    conv = E.FromRGB.from_rgb
    f_t, _, _ = train_e._FromRGBFn.apply(x.float().contiguous(), conv.weight, conv.bias, 1e-8)          # :84-92
    cv = cond_vector.float()
    for i in range(9 - block_num, E.layer_count):
        f_t = ebig_block_forward(E.decode_block[i], f_t, cv, truncation=0.4)
    return train_e.f32b_to_nchw(f_t)


# ------------------------------------------------------------------------------------------------
# frozen BigGAN-deep generator
# ------------------------------------------------------------------------------------------------
def _gen_dgrad(prep, planes):
    """Data-gradient operands of a frozen GenBlock's four convs, built from the effective spectral-norm weights the
    forward used and kept next to its operands (`GenBlock._prepared`: cached per parameter version in eval mode)."""
    if 'wd' not in prep:
        prep['wd'] = [K.pack_conv_weight_dgrad(w, planes=planes) for w in prep['eff']]
    return prep['wd']


class _GenBlockFn(torch.autograd.Function):
    """One GenBlock (biggan_generator.py:175-203) with frozen weights; coef = (a0, b0, ..., a3, b3), each [N, C]."""

    @staticmethod
    def forward(ctx, x_t, a0, b0, a1, b1, a2, b2, a3, b3, block):
        x = _f32b(x_t)
        planes = block.planes
        up = 2 if block.up_sample else 1
        p = block._prepared_k(K)
        mid = block.middle_size
        co = [t.detach().contiguous() for t in (a0, b0, a1, b1, a2, b2, a3, b3)]
        t, _ = K.affine_act(x, co[0], co[1], relu=True, planes=planes)                                            # :178-179
        t1 = K.conv(t, p['w'][0], mid, K.CONV_1X1, bias=p['b'][0], out_f32b=True)['f32b']                         # :180
        t, _ = K.affine_act(t1, co[2], co[3], relu=True, up=up, planes=planes)                                    # :182-185
        t2 = K.conv(t, p['w'][1], mid, K.CONV_3X3, bias=p['b'][1], out_f32b=True)['f32b']                         # :186
        t, _ = K.affine_act(t2, co[4], co[5], relu=True, planes=planes)
        t3 = K.conv(t, p['w'][2], mid, K.CONV_3X3, bias=p['b'][2], out_f32b=True)['f32b']                         # :188-190
        t, _ = K.affine_act(t3, co[6], co[7], relu=True, planes=planes)
        out = K.conv(t, p['w'][3], block.out_size, K.CONV_1X1, bias=p['b'][3], preact_add=x, preact_up=up,
                     out_f32b=True)['f32b']                                                                       # :192-203
        ctx.block, ctx.prep = block, p      # the backward must see the SAME effective weights (train mode: the spectral-norm
        ctx.save_for_backward(x_t, t1.t, t2.t, t3.t, *co)      # hook runs a power iteration on every call)
        return out.t

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out_t):
        block = ctx.block
        planes = block.planes
        up = 2 if block.up_sample else 1
        x_t, t1_t, t2_t, t3_t, a0, b0, a1, b1, a2, b2, a3, b3 = ctx.saved_tensors
        x, t1, t2, t3 = _f32b(x_t), _f32b(t1_t), _f32b(t2_t), _f32b(t3_t)
        d_out = _f32b(d_out_t.contiguous())
        pd = _gen_dgrad(ctx.prep, planes)
        mid = block.middle_size
        g = K.conv(K.f32b_to_act(d_out, planes), pd[3], mid, K.CONV_1X1, out_f32b=True)['f32b']
        d, _, s3 = K.affine_relu_bwd(g, t3, a3, b3, planes=planes)
        g = K.conv(d, pd[2], mid, K.CONV_3X3, out_f32b=True)['f32b']
        d, _, s2 = K.affine_relu_bwd(g, t2, a2, b2, planes=planes)
        g = K.conv(d, pd[1], mid, K.CONV_3X3, out_f32b=True)['f32b']
        d, _, s1 = K.affine_relu_bwd(g, t1, a1, b1, up=up, planes=planes)
        g = K.conv(d, pd[0], block.in_size, K.CONV_1X1, out_f32b=True)['f32b']
        del d
        _, dx, s0 = K.affine_relu_bwd(g, x, a0, b0, skip=d_out, skip_up=up, out_act=False, out_f32b=True, planes=planes)
        return (dx.t, s0[:, :, 0], s0[:, :, 1], s1[:, :, 0], s1[:, :, 1], s2[:, :, 0], s2[:, :, 1], s3[:, :, 0],
                s3[:, :, 1], None)


class _RgbTailFn(torch.autograd.Function):
    """bn -> ReLU -> conv_to_rgb -> [:, :3] -> tanh (biggan_generator.py:247-255), frozen."""

    @staticmethod
    def forward(ctx, x_t, a, b, gen):
        x = _f32b(x_t)
        planes = gen.planes
        a, b = a.detach().contiguous(), b.detach().contiguous()
        p = gen._rgb_prepared_k(K)
        ctx.prep = p
        t, _ = K.affine_act(x, a, b, relu=True, planes=planes)
        rgb16 = K.conv(t, p['w'], 16, K.CONV_3X3, bias=p['b'], out_nchw=True)['nchw']
        img = K.tanh_slice_nchw(rgb16, 3)
        ctx.gen = gen
        ctx.save_for_backward(x_t, a, b, img)
        return img

    @staticmethod
    @once_differentiable
    def backward(ctx, d_img):
        gen = ctx.gen
        planes = gen.planes
        x_t, a, b, img = ctx.saved_tensors
        x = _f32b(x_t)
        d16 = d_img.new_zeros((x.n, 16, x.h, x.w))
        d16[:, :3] = d_img * (1.0 - img * img)
        g = K.conv(K.nchw_to_act(d16, planes=planes), ctx.prep['wd'], x.c, K.CONV_3X3, out_f32b=True)['f32b']
        _, dx, s = K.affine_relu_bwd(g, x, a, b, out_act=False, out_f32b=True, planes=planes)
        return dx.t, s[:, :, 0], s[:, :, 1], None


def generator_forward(gen, cond_vector, truncation):
    """`Generator.forward` (biggan_generator.py:232-256) recorded for backward w.r.t. the condition vector."""
    from model.biggan_generator import GenBlock
    ch = gen.config.channel_width
    n = cond_vector.shape[0]
    z = torch.nn.functional.linear(cond_vector, gen._frozen(gen.gen_z), gen.gen_z.bias.detach())        # :233
    z = z.view(-1, 4, 4, 16 * ch).permute(0, 3, 1, 2).contiguous()                                   # :237-239
    x_t = train_e.nchw_to_f32b(z)
    for layer in gen.layers:
        if isinstance(layer, GenBlock):
            co = []
            for bn in (layer.bn_0, layer.bn_1, layer.bn_2, layer.bn_3):
                co += bn.coeffs_autograd(truncation, cond_vector, n, frozen=True)
            x_t = _GenBlockFn.apply(x_t, *co, layer)
        else:                                              # SelfAttn: torch graph on NCHW (bmm / softmax / max-pool)
            x_t = train_e.nchw_to_f32b(layer._forward_autograd(train_e.f32b_to_nchw(x_t)))
    a, b = gen.bn.coeffs_autograd(truncation, None, n, frozen=True)
    return _RgbTailFn.apply(x_t, a, b, gen)
