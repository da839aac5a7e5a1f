"""Fused training path of the case-1 encoder (model/E/E.py:50-135 under `loss.backward()`, E_align_s2.py:152-221).

One autograd node per encoder block instead of ~25 ATen nodes.  Tensors travel between the nodes in the kernels' own
layout (F32B, fp32 [N][C/8][H][W][8]); a node's forward is the inference chain of `BEBlock.run` (stats -> IN -> tcgen05
conv with fused noise/bias/lrelu -> ... -> 1x1 residual conv with the pool + blend epilogue) and its backward is

    be_head_bwd      d_out -> ga/4 * lrelu' -> ACT operand (+ bias_2 / noise_weight_2 / conv_3.bias gradients)
    conv_wgrad       conv_2.weight.grad                       dge_conv_forward (dgrad operand) -> d(IN_2 output)
    in_bwd_stats / in_bwd_apply (mode 1)   IN_2 Jacobian + style_2 (mean, std) gradient + lrelu' -> ACT operand
                                           (+ bias_1 / noise_weight_1 gradients)
    conv_wgrad       conv_1.weight.grad                       dge_conv_forward (dgrad operand) -> d(IN_1 output)
    conv_wgrad / dge_conv_forward          conv_3 (1x1 residual): weight gradient, data gradient
    in_bwd_stats / in_bwd_apply (mode 0)   IN_1 Jacobian + style_1 gradient + pooled residual gradient -> d_x (F32B)

i.e. 11 launches per block, every one a dge_b200 kernel.  The two `inver_mod` Linear layers stay torch ops on the
[N, 2C] style vectors the node returns (two tiny GEMMs per block).  `K` is the kernel namespace (dge_b200.ops); the CPU
tests swap it for a torch emulation of the same C-ABI functions to check the chain rule without a GPU.
"""
import torch
from torch.autograd.function import once_differentiable

from . import graphs
from . import ops

K = ops          # kernel namespace (tests/emu_ops.py replaces it on the CPU)

GA, GB = 0.111, 0.889          # E.py:84
SLOPE = 0.2                    # E.py:62,75; net.py:239


def _capturing():
    return K is ops and torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def _packed(w, planes, dgrad=False):
    """Packed forward / data-gradient operand of a conv weight, reused until the weight changes (ops.weight_key: an
    optimiser step retires it).  One iteration of the inversion loop runs the encoder twice on the same weights, and
    the two backward passes of an iteration straddle a step.  The cache lives ON the parameter object, so it dies with
    it (a recycled device address can never alias another model's weights)."""
    if _capturing():                 # a replayed chain packs from the LIVE parameter on every replay (graphs.py)
        return K.pack_conv_weight_dgrad(w, planes=planes) if dgrad else K.pack_conv_weight(w, planes=planes)
    cache = w.__dict__.setdefault('_dge_packed', {})
    slot = (planes, dgrad, K is ops)
    key = K.weight_key(w)
    hit = cache.get(slot)
    if hit is not None and hit[0] == key:
        return hit[1]
    op = K.pack_conv_weight_dgrad(w, planes=planes) if dgrad else K.pack_conv_weight(w, planes=planes)
    cache[slot] = (key, op)
    return op


def _f32b(t):
    n, c8, h, w, _ = t.shape
    return K.F32B.wrap(t, n, c8 * 8, h, w)


# ---- case-2 encoder (model/E/E_Blur.py): the stride-2 `transform_kernel` conv in space-to-depth form -----------------
def _k4(w):
    """lreq.py:144-147: the 4x4 stride-2 kernel = 0.25 * (sum of the four 1-pixel shifts of the zero-padded 3x3 kernel)."""
    k = torch.nn.functional.pad(w, (1, 1, 1, 1))
    return (k[:, :, 1:, 1:] + k[:, :, :-1, 1:] + k[:, :, 1:, :-1] + k[:, :, :-1, :-1]) * 0.25


_TAP_TO_S2D = ((0, 1), (1, 0), (1, 1), (2, 0))     # 4x4 tap a reads row 2Y + a - 1 = 2(Y + d) + p  ->  (d + 1, p)


_s2d_index_cache = {}


def _s2d_index(device):
    """Index arrays [4, 4] (phase_y, phase_x, tap_y, tap_x) of the 16 slabs the 4x4 kernel occupies in [2][2][C][3][3]."""
    hit = _s2d_index_cache.get(device)
    if hit is None:
        d = torch.tensor([t[0] for t in _TAP_TO_S2D], device=device)
        p = torch.tensor([t[1] for t in _TAP_TO_S2D], device=device)
        hit = tuple(t.contiguous() for t in (p[:, None].expand(4, 4), p[None, :].expand(4, 4), d[:, None].expand(4, 4),
                                             d[None, :].expand(4, 4)))
        _s2d_index_cache[device] = hit
    return hit


def _w_s2d(w):
    """The same conv over the space-to-depth map (channel (2*py + px)*C + i holds x[2Y+py][2X+px]): a stride-1 3x3 conv
    [Cout][4C][3][3] with 16 of its 36 (tap, phase) slabs non-zero."""
    k4 = _k4(w)
    co, c = k4.shape[:2]
    ws = k4.new_zeros((co, 2, 2, c, 3, 3))
    py, px, dy, dx = _s2d_index(w.device)
    ws[:, py, px, :, dy, dx] = k4.permute(2, 3, 0, 1)          # (advanced indices split by a slice: result dims lead)
    return ws.reshape(co, 4 * c, 3, 3)


def _dw3_from_s2d(dws, c):
    """Transpose of `_w_s2d`: gradient of the 3x3 `transform_kernel` parameter from the gradient of the 3x3 conv over the
    space-to-depth operand (gather the 16 slabs -> d K4; each 3x3 weight feeds the four 4x4 taps it was shifted onto)."""
    co = dws.shape[0]
    py, px, dy, dx = _s2d_index(dws.device)
    dk4 = dws.view(co, 2, 2, c, 3, 3)[:, py, px, :, dy, dx].permute(2, 3, 0, 1)        # [co][c][4][4]
    return (dk4[:, :, :3, :3] + dk4[:, :, 1:, :3] + dk4[:, :, :3, 1:] + dk4[:, :, 1:, 1:]) * 0.25


def _packed_strided(w, planes, dgrad):
    """Forward (16-tap DOWN4X4S2) / data-gradient (3x3 over space-to-depth) operands of a `transform_kernel` conv, cached
    on the parameter like `_packed`."""
    wd = w.detach()
    if _capturing():
        return K.pack_conv_weight_dgrad(_w_s2d(wd).contiguous(), planes=planes) if dgrad else \
            K.pack_conv_weight(_k4(wd).contiguous(), planes=planes)
    cache = w.__dict__.setdefault('_dge_packed', {})
    slot = (planes, 's2d-dgrad' if dgrad else 'down4', K is ops)
    key = K.weight_key(w)
    hit = cache.get(slot)
    if hit is not None and hit[0] == key:
        return hit[1]
    op = K.pack_conv_weight_dgrad(_w_s2d(wd).contiguous(), planes=planes) if dgrad else \
        K.pack_conv_weight(_k4(wd).contiguous(), planes=planes)
    cache[slot] = (key, op)
    return op


def _depth_to_space(g):
    """F32B [N][4C/8][H/2][W/2][8] with channel (2*py + px)*C + i  ->  F32B [N][C/8][H][W][8]."""
    n, c4, h2, w2 = g.n, g.c, g.h, g.w
    c = c4 // 4
    t = g.t.view(n, 2, 2, c // 8, h2, w2, 8).permute(0, 3, 4, 1, 5, 2, 6).reshape(n, c // 8, 2 * h2, 2 * w2, 8)
    return K.F32B.wrap(t.contiguous(), n, c, 2 * h2, 2 * w2)


class _FromRGBFn(torch.autograd.Function):
    """img NCHW -> (f F32B tensor, style0, mean_rstd0): net.py:231-240 + the first block's statistics (E.py:51-53)."""

    @staticmethod
    def forward(ctx, img, weight, bias, eps):
        f, style, mr = K.from_rgb_stats_any(img, weight, bias, SLOPE, eps)
        ctx.save_for_backward(img, f.t, weight)
        ctx.mark_non_differentiable(style, mr)
        return f.t, style, mr

    @staticmethod
    @once_differentiable
    def backward(ctx, d_f, _ds, _dmr):
        img, f_t, weight = ctx.saved_tensors
        d_img = None
        if ctx.needs_input_grad[0]:      # embedding_img.py:88 -- E(imgs2) back-propagates into the generator
            sums, d_img = K.from_rgb_bwd(_f32b(d_f.contiguous()), _f32b(f_t), img, SLOPE, weight=weight)
        else:
            sums = K.from_rgb_bwd(_f32b(d_f.contiguous()), _f32b(f_t), img, SLOPE)
        cimg = img.shape[1]
        dw = sums[:, :cimg].reshape(weight.shape)
        db = sums[:, 3].contiguous()
        return d_img, dw, db, None


def _be_fwd(x_t, pre_style, pre_mr, w1, w2, w3, b3, nw1, b1, nw2, b2, cfg):
    """Forward kernel chain of one BEBlock -> ((out F32B tensor, style1, style2), saved tensors, meta)."""
    has_last, planes, eps1, eps2, noise1, noise2 = cfg[:6]
    blur = cfg[6] if len(cfg) > 6 else None        # E_Blur: 'strided' (transform_kernel conv_2) / 'plain' / None (E.py)
    x = _f32b(x_t)
    n, c, h, w = x.n, x.c, x.h, x.w
    cout = w2.shape[0] if has_last else (w3.shape[0] if w3 is not None else c)
    if pre_style is not None:
        style1, mr1 = pre_style, pre_mr
    else:
        style1, mr1 = K.instance_stats(x, eps1)                                            # E.py:51-53 + IN stats
    rp = None
    if has_last and w3 is not None:
        xn, rp = K.instance_norm_pool(x, mr1, planes=planes)                               # :58 and :78
    else:
        xn, _ = K.instance_norm(x, mr1, planes=planes)                                     # :58
    y1 = K.conv(xn, _packed(w1, planes), c, K.CONV_3X3, noise=noise1, noise_batched=True,
                noise_w=nw1.detach().reshape(-1), bias=b1.detach().reshape(-1), slope=SLOPE, out_f32b=True)['f32b']
    style2, mr2 = K.instance_stats(y1, eps2)                                               # :64-66
    y2 = None
    if has_last:
        if blur:
            y1n = K.instance_norm_blur(y1, mr2, s2d=blur == 'strided', planes=planes)      # E_Blur.py:69-71
        else:
            y1n, _ = K.instance_norm(y1, mr2, planes=planes)                               # :69
        # training keeps the activated conv_2 output before the pool / blend: its sign is the leaky-ReLU mask
        if blur == 'strided':                                                              # E_Blur.py:72 (half size)
            y2 = K.conv(y1n, _packed_strided(w2, planes, False), cout, K.CONV_DOWN4X4S2, noise=noise2,
                        noise_batched=True, noise_w=nw2.detach().reshape(-1), bias=b2.detach().reshape(-1),
                        slope=SLOPE, out_f32b=True)['f32b']
        else:
            y2 = K.conv(y1n, _packed(w2, planes), cout, K.CONV_3X3, noise=noise2, noise_batched=True,
                        noise_w=nw2.detach().reshape(-1), bias=b2.detach().reshape(-1), slope=SLOPE,
                        out_f32b=True)['f32b']                                              # :72-75
        y2_pool = blur != 'strided'
        if w3 is not None:
            out = K.conv(rp, _packed(w3, planes), cout, K.CONV_1X1, bias=b3.detach(),
                         blend_src=y2, blend_pool=y2_pool, blend_a=GA, blend_b=GB, out_f32b=True)['f32b']   # :76-84
        else:
            out = K.blend(y2, x, GA, GB, pool=3 if y2_pool else 2)
    else:
        y1n = None
        _, y1n_f = K.instance_norm(y1, mr2, out_act=False, out_f32b=True)                  # :69
        if w3 is not None:
            rp = K.f32b_to_act(x, planes)
            out = K.conv(rp, _packed(w3, planes), cout, K.CONV_1X1, bias=b3.detach(),
                         blend_src=y1n_f, blend_pool=False, blend_a=GA, blend_b=GB, out_f32b=True)['f32b']
        else:
            out = K.blend(y1n_f, x, GA, GB, pool=False)
    meta = {'cfg': (has_last, planes, (n, c, h, w), cout), 'blur': blur,
            'has': (w3 is not None, y1n is not None, rp is not None, y2 is not None)}
    saved = [x_t, mr1, style1, xn.t, y1.t, mr2, style2, noise1, w1]
    if has_last:
        saved += [y1n.t, y2.t, noise2, w2]
    if w3 is not None:
        saved += [rp.t, w3]
    s1 = style1.clone() if pre_style is not None else style1
    return (out.t, s1, style2), saved, meta


def _be_bwd(meta, sv, d_out_t, d_style1, d_style2):
    """Backward kernel chain of one BEBlock -> (d x, d conv_1.w, d conv_2.w, d conv_3.w, d conv_3.b, d noise_weight_1,
    d bias_1, d noise_weight_2, d bias_2); entries of absent parameters are None."""
    has_last, planes, (n, c, h, w), cout = meta['cfg']
    has_w3 = meta['has'][0]
    sv = list(sv)
    x_t, mr1, style1, xn_t, y1_t, mr2, style2, noise1, w1 = sv[:9]
    p = 9
    if has_last:
        y1n_t, y2_t, noise2, w2 = sv[p:p + 4]
        p += 4
    if has_w3:
        rp_t, w3 = sv[p:p + 2]
    x, y1 = _f32b(x_t), _f32b(y1_t)
    xn = K.Act.wrap(xn_t, n, c, h, w, planes)
    d_out = _f32b(d_out_t.contiguous())
    dw2 = dnw2 = db2 = dw3 = db3 = None
    blur = meta['blur']
    if has_last and blur == 'strided':
        # conv_2 ran at stride 2: y2 and d_out share a size, so the head is a scaled leaky-ReLU mask (no pool) ...
        y2 = _f32b(y2_t)
        ga_tab = torch.zeros((n, cout, 2), dtype=torch.float32, device=mr2.device)             # (mean, rstd) = (0, GA)
        ga_tab[:, :, 1] = GA
        zeros = torch.zeros((n, cout, 2), dtype=torch.float64, device=mr2.device)
        dy2, s = K.in_bwd_apply(d_out, y2, ga_tab, None, None, zeros, 1, noise=noise2, slope=SLOPE, planes=planes)
        db2, dnw2 = s[0].view(1, -1, 1, 1), s[1].view(1, -1, 1, 1)
        dres = None
        if has_w3:
            dres = K.scale_f32b(d_out, GB, to_act=True, planes=planes)
            db3 = GB * K.f32b_channel_sums(d_out)
        # ... and its gradients are those of a 3x3 conv over the space-to-depth operand the forward kept
        xs = K.Act.wrap(y1n_t, n, 4 * c, h // 2, w // 2, planes)
        dws = K.conv_wgrad(dy2, xs, 3)
        dw2 = _dw3_from_s2d(dws, c)
        g1 = _depth_to_space(K.conv(dy2, _packed_strided(w2, planes, True), 4 * c, K.CONV_3X3, out_f32b=True)['f32b'])
        del dy2
        g1 = K.sg1_post(g1, 0, n, c, h, w, slope=1.0)                                     # blur^T = blur (:71)
    elif has_last:
        y2 = _f32b(y2_t)
        dy2, dres, s = K.be_head_bwd(d_out, y2, noise2, GA, GB, SLOPE, want_dres=has_w3, planes=planes)
        db2, dnw2 = s[0].view(1, -1, 1, 1), s[1].view(1, -1, 1, 1)
        if has_w3:
            db3 = s[2]
        dw2 = K.conv_wgrad(dy2, K.Act.wrap(y1n_t, n, c, h, w, planes), 3)
        g1 = K.conv(dy2, _packed(w2, planes, True), c, K.CONV_3X3, out_f32b=True)['f32b']
        del dy2
        if blur:
            g1 = K.sg1_post(g1, 0, n, c, h, w, slope=1.0)                                 # blur^T = blur (E_Blur.py:71)
    else:
        g1 = K.scale_f32b(d_out, GA)                       # out = GA * IN_2(y1) + GB * residual  (E.py:69,84)
        dres = None
        if has_w3:
            dres = K.scale_f32b(d_out, GB, to_act=True, planes=planes)
            db3 = GB * K.f32b_channel_sums(d_out)
    # IN_2 + the style (mean, std) of y1 + lrelu' of conv_1's activation  ->  operand of conv_1's gradients
    st2 = K.in_bwd_stats(g1, y1, mr2)
    dy1, s = K.in_bwd_apply(g1, y1, mr2, style2, d_style2, st2, 1, noise=noise1, slope=SLOPE, planes=planes)
    db1, dnw1 = s[0].view(1, -1, 1, 1), s[1].view(1, -1, 1, 1)
    del g1
    dw1 = K.conv_wgrad(dy1, xn, 3)
    g0 = K.conv(dy1, _packed(w1, planes, True), c, K.CONV_3X3, out_f32b=True)['f32b']
    del dy1
    # residual branch (E.py:78-84)
    if has_w3:
        rh, rw = (h // 2, w // 2) if has_last else (h, w)
        dw3 = K.conv_wgrad(dres, K.Act.wrap(rp_t, n, c, rh, rw, planes), 1)
        d_rp = K.conv(dres, _packed(w3, planes, True), c, K.CONV_1X1, out_f32b=True)['f32b']
        rscale = 0.25 if has_last else 1.0
    else:
        d_rp, rscale = d_out, GB * (0.25 if has_last else 1.0)
    st1 = K.in_bwd_stats(g0, x, mr1)
    d_x = K.in_bwd_apply(g0, x, mr1, style1, d_style1, st1, 0, res=d_rp, rscale=rscale, res_pool=has_last)
    return d_x.t, dw1, dw2, dw3, db3, dnw1, db1, dnw2, db2


class _BEBlockFn(torch.autograd.Function):
    """One BEBlock (E.py:50-85).  Inputs: x (F32B tensor), optional precomputed (style1, mean_rstd1), the block's
    parameters.  Outputs: (out F32B tensor, style1 [N, 2C], style2 [N, 2C]).

    CUDA-graph replay (dge_b200/graphs.py, opt-in): unlike the frozen nodes this one's parameters change between replays
    (LREQAdam steps twice per iteration, and the two backward passes of an iteration straddle a step), so the captured
    chains pack their conv operands from the LIVE parameters on every replay (`_packed` under capture) and the slot key is
    the parameters' storage, not their values.  An iteration of the inversion loop keeps TWO encoder passes alive
    (`E(imgs1)`, `E(imgs2)`: embedding_img.py:86,88), so every block alternates between two slots; a third pass alive at
    once is refused at its backward like any overwritten pass."""

    @staticmethod
    def forward(ctx, x_t, pre_style, pre_mr, w1, w2, w3, b3, nw1, b1, nw2, b2, cfg, owner=None):
        params = (w1, w2, w3, b3, nw1, b1, nw2, b2)
        if graphs.GRAPHS and owner is not None and K is ops and x_t.is_cuda:
            noise1, noise2 = cfg[4], cfg[5]
            has_n2, has_pre = noise2 is not None, pre_style is not None
            dyn = [x_t, noise1] + ([noise2] if has_n2 else []) + ([pre_style, pre_mr] if has_pre else [])

            def fwd(*d):
                i = 2 + has_n2
                cfg2 = tuple(cfg[:4]) + (d[1], d[2] if has_n2 else None) + tuple(cfg[6:])
                outs, saved, meta = _be_fwd(d[0], d[i] if has_pre else None, d[i + 1] if has_pre else None, *params, cfg2)
                return outs, (saved, meta)

            rr = owner.__dict__.get('_dge_rr', 0)
            owner.__dict__['_dge_rr'] = rr + 1
            slot = ('be', tuple(x_t.shape), has_pre, has_n2, tuple(cfg[:4]), tuple(cfg[6:]), rr % 2)
            key = tuple(None if p is None else (p.data_ptr(), tuple(p.shape)) for p in params)
            outs, ctx.handle = graphs.forward(owner, slot, key, dyn, fwd, 'train_e (encoder block)')
            return outs
        ctx.handle = None
        outs, saved, ctx.meta = _be_fwd(x_t, pre_style, pre_mr, *params, cfg)
        ctx.save_for_backward(*saved)
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out_t, d_style1, d_style2):
        if ctx.handle is not None:
            g = graphs.backward(ctx.handle, (d_out_t, d_style1, d_style2),
                                lambda sm, a, b, c: _be_bwd(sm[1], sm[0], a, b, c), 'train_e (encoder block)')
        else:
            g = _be_bwd(ctx.meta, ctx.saved_tensors, d_out_t, d_style1, d_style2)
        return (g[0], None, None) + tuple(g[1:]) + (None, None)


class _F32BToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return _f32b(t).to_nchw()

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return K.nchw_to_f32b(g.contiguous().float()).t


class _NCHWToF32B(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return K.nchw_to_f32b(x.contiguous().float()).t

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return _f32b(g.contiguous()).to_nchw()


def nchw_to_f32b(x):
    return _NCHWToF32B.apply(x)


def f32b_to_nchw(t):
    return _F32BToNCHW.apply(t)


def block_forward(block, x_t, pre=None):
    """BEBlock under autograd on an F32B tensor -> (out F32B tensor, w1, w2)  (E.py:50-85)."""
    n, _, h, w, _ = x_t.shape
    dev = x_t.device
    noise1 = block._noise(n, h, w, dev).reshape(n, h, w).contiguous()                          # E.py:60 (RNG order kept)
    noise2 = None
    has_w3 = block.inputs != block.outputs
    if has_w3 and not block.has_last_conv:
        # E.py:84 adds the `inputs`-channel features to the `outputs`-channel residual: the reference fails here too
        raise RuntimeError(f'BEBlock without a last conv needs inputs == outputs (got {block.inputs}, {block.outputs}): '
                           'choose start_features so that the last block runs at maxf (16 -> 1024, 32 -> 512, 64 -> 256)')
    blur = None
    if hasattr(block, 'blur') and block.has_last_conv:                                         # model/E/E_Blur.py
        blur = 'strided' if block.fused_scale else 'plain'
    if blur == 'strided':                                                                      # E_Blur.py:73: half-size noise
        noise2 = block._noise(n, h // 2, w // 2, dev).reshape(n, h // 2, w // 2).contiguous()
    elif block.has_last_conv:
        noise2 = block._noise(n, h, w, dev).reshape(n, h, w).contiguous()                      # :73
    cfg = (block.has_last_conv, block.planes, block.instance_norm_1.eps, block.instance_norm_2.eps, noise1, noise2, blur)
    pre_style, pre_mr = pre if pre is not None else (None, None)
    out_t, style1, style2 = _BEBlockFn.apply(
        x_t, pre_style, pre_mr, block.conv_1.weight, block.conv_2.weight if block.has_last_conv else None,
        block.conv_3.weight if has_w3 else None, block.conv_3.bias if has_w3 else None, block.noise_weight_1,
        block.bias_1, block.noise_weight_2, block.bias_2, cfg, block)
    lin = torch.nn.functional.linear
    w1 = lin(style1, block.inver_mod1.weight, block.inver_mod1.bias)                           # :54
    w2 = lin(style2, block.inver_mod2.weight, block.inver_mod2.bias)                           # :67
    return out_t, w1, w2


def encoder_forward(E, x, block_num=9):
    """`BE.forward` (E.py:122-135) recorded for backward with one fused node per block."""
    conv = E.FromRGB.from_rgb
    first = 9 - block_num
    eps0 = E.decode_block[first].instance_norm_1.eps if first < E.layer_count else 1e-8
    f_t, style0, mr0 = _FromRGBFn.apply(x.float().contiguous(), conv.weight, conv.bias, eps0)
    w = torch.tensor(0)
    n = x.shape[0]
    for i in range(first, E.layer_count):
        f_t, w1, w2 = block_forward(E.decode_block[i], f_t, pre=(style0, mr0) if i == first else None)
        w_ = torch.cat((w2.view(n, 1, 512), w1.view(n, 1, 512)), dim=1)                        # E.py:131
        w = w_ if i == first else torch.cat((w_, w), dim=1)
    return _F32BToNCHW.apply(f_t), w
