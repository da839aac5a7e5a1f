"""Plain-torch (CPU) statement of the C-ABI functions the fused training path calls (include/dge_b200.h), on the kernels'
own layouts (ACT bf16 [N][C/8][planes][H][W][8], F32B fp32 [N][C/8][H][W][8]).  TEST INFRASTRUCTURE: it is
  * the executable specification every new CUDA kernel is compared with on the GPU (tests/test_train_kernels_gpu.py),
  * a stand-in for `dge_b200.ops` so that the chain rule coded in dge_b200/train_e.py / train_g.py can be checked against
    gradients of the unmodified reference without a GPU (tests/test_train_fused_cpu.py).
The product never imports this module.
"""
import torch
import torch.nn.functional as F

from dge_b200.ops import weight_key  # noqa: F401  (host-side cache key, no device work)

CONV_3X3, CONV_1X1, CONV_UP3X3, CONV_DOWN4X4S2 = 0, 1, 2, 3


def _to_blocked(x):            # NCHW -> [N][C/8][H][W][8]
    n, c, h, w = x.shape
    return x.reshape(n, c // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous()


def _from_blocked(t):          # [N][C/8][H][W][8] -> NCHW
    n, c8, h, w, _ = t.shape
    return t.permute(0, 1, 4, 2, 3).reshape(n, c8 * 8, h, w).contiguous()


class F32B:
    def __init__(self, n, c, h, w, device="cpu"):
        self.n, self.c, self.h, self.w = n, c, h, w
        self.t = torch.empty((n, c // 8, h, w, 8), dtype=torch.float32, device=device)

    @classmethod
    def wrap(cls, t, n, c, h, w):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n * c * h * w
        o = cls.__new__(cls)
        o.n, o.c, o.h, o.w, o.t = n, c, h, w, t.view(n, c // 8, h, w, 8)
        return o

    @classmethod
    def of(cls, x):
        n, c, h, w = x.shape
        return cls.wrap(_to_blocked(x.float()), n, c, h, w)

    def to_nchw(self):
        return _from_blocked(self.t)


class Act:
    def __init__(self, n, c, h, w, planes=2, device="cpu"):
        self.n, self.c, self.h, self.w, self.planes = n, c, h, w, planes
        self.t = torch.empty((n, c // 8, planes, h, w, 8), dtype=torch.bfloat16, device=device)

    @classmethod
    def wrap(cls, t, n, c, h, w, planes=2):
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() == n * c * planes * h * w
        o = cls.__new__(cls)
        o.n, o.c, o.h, o.w, o.planes, o.t = n, c, h, w, planes, t.view(n, c // 8, planes, h, w, 8)
        return o

    @classmethod
    def of(cls, x, planes=2):
        """fp32 NCHW -> hi (+ lo) bf16 planes, the split the kernels use (dge_common.cuh split_bf16)."""
        n, c, h, w = x.shape
        b = _to_blocked(x.float())
        hi = b.to(torch.bfloat16)
        parts = [hi]
        if planes == 2:
            parts.append((b - hi.float()).to(torch.bfloat16))
        return cls.wrap(torch.stack(parts, dim=2).contiguous(), n, c, h, w, planes)

    def planes_nchw(self):
        """-> list of fp32 NCHW planes (hi, lo)."""
        return [_from_blocked(self.t[:, :, p].float()) for p in range(self.planes)]

    def to_nchw(self):
        ps = self.planes_nchw()
        return ps[0] if len(ps) == 1 else ps[0] + ps[1]


def nchw_to_f32b(x):
    return F32B.of(x)


def nchw_to_act(x, scale=None, planes=2, batch=None):
    x = x.float()
    if batch is not None and x.shape[0] == 1:
        x = x.expand(batch, -1, -1, -1)
    if scale is not None:
        x = x * scale.view(x.shape[0], -1, 1, 1)
    return Act.of(x, planes)


def f32b_to_act(x, planes=2):
    return Act.of(x.to_nchw(), planes)


def scale_f32b(x, a, to_act=False, planes=2):
    y = F32B.wrap(x.t * float(a), x.n, x.c, x.h, x.w)
    return f32b_to_act(y, planes) if to_act else y


def f32b_channel_sums(x):
    return x.t.sum(dim=(0, 2, 3)).reshape(-1)


# ------------------------------------------------------------------------------------------------
# weights / conv
# ------------------------------------------------------------------------------------------------
class _Packed:
    """Stands for a WPK tensor: the fp32 OIHW weight the packing kernel was given (split into hi + lo at use)."""

    def __init__(self, w, planes, dgrad=False, flip=False):
        self.w, self.planes, self.dgrad, self.flip = w.detach().float(), planes, dgrad, flip

    def data_ptr(self):
        return self.w.data_ptr()


def pack_conv_weight(w, scale=1.0, flip=False, planes=2):
    return _Packed(w * scale, planes, flip=flip)


def pack_conv_weight_dgrad(w, scale=1.0, planes=2):
    return _Packed(w * scale, planes, dgrad=True)


def _split(t):
    hi = t.to(torch.bfloat16).float()
    return hi, (t - hi).to(torch.bfloat16).float()


def _conv3(xa, w, planes, fn):
    """hi*hi + hi*lo + lo*hi (planes == 2) or hi*hi (planes == 1) with fp32 accumulation."""
    xp = xa.planes_nchw()
    wh, wl = _split(w)
    y = fn(xp[0], wh)
    if planes == 2:
        y = y + fn(xp[0], wl) + fn(xp[1], wh)
    return y


def conv(x, wpk, cout, kind=CONV_3X3, *, demod=None, noise=None, noise_batched=False, noise_w=None, noise_scalar=0.0,
         bias=None, slope=1.0, gain=1.0, blend_src=None, blend_pool=False, blend_a=0.0, blend_b=1.0, preact_add=None,
         preact_up=1, out_act=False, out_planes=None, out_scale=None, out_f32b=False, out_f32b_into=None,
         out_f32b_pool=False, out_nchw=False, rgb_w=None, rgb_out=None, checker=False, out_hw=None):
    assert isinstance(x, Act) and isinstance(wpk, _Packed) and out_f32b_into is None
    n, h, w_ = x.n, x.h, x.w
    wt = wpk.w
    if kind == CONV_UP3X3:
        # raw transposed conv: t[2Y+ky][2X+kx] += x[Y][X] * W[o][i][ky][kx] with the weight the PACKER already flipped
        wf = wt.flip(2, 3) if wpk.flip else wt
        raw = _conv3(x, wf, x.planes, lambda a, b: F.conv_transpose2d(a, b.transpose(0, 1), stride=2))
        return {"raw_up": _to_blocked(raw)}
    if kind == CONV_DOWN4X4S2:
        # x is space-to-depth: channel block 2*py+px holds xin[2y+py][2x+px]; 4x4 stride-2 pad-1 conv of xin
        c = x.c // 4
        planes = x.planes_nchw()
        outs = None
        hh, ww = out_hw if out_hw is not None else (h, w_)

        def d2s(p):
            v = p.view(n, 2, 2, c, p.shape[2], p.shape[3])                  # [n][py][px][c][h][w]
            return v.permute(0, 3, 4, 1, 5, 2).reshape(n, c, 2 * p.shape[2], 2 * p.shape[3])
        xin = [d2s(p) for p in planes]
        wh, wl = _split(wt)
        f = lambda a, b: F.conv2d(a, b, stride=2, padding=1)
        y = f(xin[0], wh)
        if x.planes == 2:
            y = y + f(xin[0], wl) + f(xin[1], wh)
        y = y[:, :, :hh, :ww]
        h, w_ = hh, ww
    else:
        k = 3 if kind == CONV_3X3 else 1
        if wpk.dgrad:
            wt = wt.flip(2, 3).transpose(0, 1).contiguous()
        y = _conv3(x, wt, x.planes, lambda a, b: F.conv2d(a, b, padding=k // 2))
    if demod is not None:
        y = y * demod.view(n, cout, 1, 1)
    if noise is not None:
        nz = noise.reshape(n if noise_batched else 1, 1, h, w_)
        y = y + nz * (noise_w.view(1, -1, 1, 1) if noise_w is not None else noise_scalar)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    if preact_add is not None:             # identity branch added before the activation (channel drop + nearest up)
        sk = preact_add.to_nchw()[:, :cout]
        if preact_up > 1:
            sk = sk.repeat_interleave(preact_up, dim=2).repeat_interleave(preact_up, dim=3)
        y = y + sk
    y = torch.where(y < 0, y * slope, y) * gain
    if blend_src is not None:
        s = blend_src.to_nchw()
        if blend_pool:
            s = F.avg_pool2d(s, 2, 2)
        y = blend_a * s + blend_b * y
    res = {}
    if rgb_w is not None:
        rgb_out += torch.einsum("nkc,nchw->nkhw", rgb_w, y)
    if out_f32b:
        res["f32b"] = F32B.of(y)
    if out_f32b_pool:
        res["f32b_pool"] = F32B.of(F.avg_pool2d(y, 2, 2))
    if out_nchw:
        res["nchw"] = y
    if out_act:
        ys = y if out_scale is None else y * out_scale.view(n, cout, 1, 1)
        res["act"] = Act.of(ys, out_planes or x.planes)
    return res


def conv_wgrad(dy, x, ksize, out=None, accumulate=False):
    dp, xp = dy.planes_nchw(), x.planes_nchw()
    shape = (dy.c, x.c, ksize, ksize)
    gw = lambda d, xx: torch.nn.grad.conv2d_weight(xx, shape, d, padding=ksize // 2)
    g = gw(dp[0], xp[0])
    if dy.planes == 2:
        g = g + gw(dp[0], xp[1]) + gw(dp[1], xp[0])
    if accumulate:
        out += g
        return out
    return g


# ------------------------------------------------------------------------------------------------
# encoder forward pieces
# ------------------------------------------------------------------------------------------------
def instance_stats(x, eps=1e-8):
    v = x.to_nchw().double()
    m = v.mean(dim=(2, 3))
    var = (v * v).mean(dim=(2, 3)) - m * m
    var = var.clamp_min(0)
    style = torch.cat((m, var.sqrt()), dim=1).float()
    mr = torch.stack((m, 1.0 / torch.sqrt(var + eps)), dim=2).float().contiguous()
    return style, mr


def _in_apply(x, mr):
    return (x.to_nchw() - mr[:, :, 0, None, None]) * mr[:, :, 1, None, None]


def instance_norm(x, mean_rstd, planes=2, out_act=True, out_f32b=False, gamma=None, beta=None):
    y = _in_apply(x, mean_rstd)
    return (Act.of(y, planes) if out_act else None), (F32B.of(y) if out_f32b else None)


def instance_norm_pool(x, mean_rstd, planes=2):
    return Act.of(_in_apply(x, mean_rstd), planes), Act.of(F.avg_pool2d(x.to_nchw(), 2, 2), planes)


def instance_norm_blur(x, mean_rstd, s2d=False, planes=2):
    y = _blur3(_in_apply(x, mean_rstd))
    if s2d:                                # channel (2*py + px)*C + i holds y[2Y+py][2X+px]
        n, c, h, w = y.shape
        y = y.view(n, c, h // 2, 2, w // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, h // 2, w // 2)
    return Act.of(y, planes)


def from_rgb_stats_any(img, w, b, slope=0.2, eps=1e-8):
    f = F.leaky_relu(F.conv2d(img.float(), w.detach().float(), None if b is None else b.detach().float()), slope)
    fb = F32B.of(f)
    return (fb,) + instance_stats(fb, eps)


def blend(a_src, b_src, a, b, pool):
    pool = 3 if pool is True else int(pool)
    va, vb = a_src.to_nchw(), b_src.to_nchw()
    if pool & 1:
        va = F.avg_pool2d(va, 2, 2)
    if pool & 2:
        vb = F.avg_pool2d(vb, 2, 2)
    return F32B.of(a * va + b * vb)


# ------------------------------------------------------------------------------------------------
# training step: backward kernels (csrc/train_bwd.cu)
# ------------------------------------------------------------------------------------------------
def be_head_bwd(d_out, y2, noise, ga, gb, slope, want_dres=True, planes=2):
    d = d_out.to_nchw()
    y = y2.to_nchw()
    up = d.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3) * (ga * 0.25)
    dy2 = torch.where(y > 0, up, up * slope)
    dres = gb * d
    nz = noise.view(y2.n, 1, y2.h, y2.w) if noise is not None else torch.zeros(())
    sums = torch.stack((dy2.sum(dim=(0, 2, 3)), (dy2 * nz).sum(dim=(0, 2, 3)), dres.sum(dim=(0, 2, 3))))
    return Act.of(dy2, planes), (Act.of(dres, planes) if want_dres else None), sums


def in_bwd_stats(g, x, mean_rstd):
    gv, xn = g.to_nchw().double(), _in_apply(x, mean_rstd).double()
    return torch.stack((gv.sum(dim=(2, 3)), (gv * xn).sum(dim=(2, 3))), dim=2).contiguous()


def in_bwd_apply(g, x, mean_rstd, style, dstyle, sums, mode, res=None, rscale=0.0, res_pool=False, noise=None,
                 slope=0.2, planes=2, gscale=None, out_kind="act"):
    gv, xv = g.to_nchw(), x.to_nchw()
    n, c, h, w = xv.shape
    hw = h * w
    m, r = mean_rstd[:, :, 0, None, None], mean_rstd[:, :, 1, None, None]
    a = (sums[:, :, 0] / hw).float()[:, :, None, None]
    b = (sums[:, :, 1] / hw).float()[:, :, None, None]
    xc = xv - m
    v = r * (gv - a - xc * r * b)
    if gscale is not None:
        v = v * gscale[:, :, None, None]
    if dstyle is not None:
        sd = style[:, c:, None, None]
        v = v + dstyle[:, :c, None, None] / hw + torch.where(sd > 0, dstyle[:, c:, None, None] / hw / sd, 0.0) * xc
    if mode == 0:
        if res is not None:
            rv = res.to_nchw()
            if res_pool:
                rv = rv.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
            v = v + rscale * rv
        return F32B.of(v)
    v = torch.where(xv > 0, v, v * slope)
    nz = noise.view(n, 1, h, w) if noise is not None else torch.zeros(())
    return (Act.of(v, planes) if out_kind == "act" else F32B.of(v)), \
        torch.stack((v.sum(dim=(0, 2, 3)), (v * nz).sum(dim=(0, 2, 3))))


def affine_act(x, a, b, relu=False, up=1, planes=2, out_act=True, out_f32b=False):
    y = x.to_nchw() * a[:, :, None, None] + b[:, :, None, None]
    if relu:
        y = F.relu(y)
    if up > 1:
        y = y.repeat_interleave(up, dim=2).repeat_interleave(up, dim=3)
    return (Act.of(y, planes) if out_act else None), (F32B.of(y) if out_f32b else None)


def tanh_slice_nchw(x, nch):
    return torch.tanh(x[:, :nch]).contiguous()


def _sum_pool(t, k):
    return t if k == 1 else F.avg_pool2d(t, k, k) * float(k * k)


def affine_relu_bwd(g, x, a, b, slope=0.0, up=1, skip=None, skip_up=1, out_act=True, out_f32b=False, planes=2):
    xv = x.to_nchw()
    gs = _sum_pool(g.to_nchw(), up)
    pre = xv * a[:, :, None, None] + b[:, :, None, None]
    d = torch.where(pre > 0, gs, gs * slope)
    sums = torch.stack(((d * xv).sum(dim=(2, 3)), d.sum(dim=(2, 3))), dim=2).contiguous()
    v = d * a[:, :, None, None]
    if skip is not None:
        v = v.clone()
        v[:, :skip.c] += _sum_pool(skip.to_nchw(), skip_up)
    return (Act.of(v, planes) if out_act else None), (F32B.of(v) if out_f32b else None), sums


def from_rgb_bwd(d_f, f, img, slope=0.2, weight=None):
    d = d_f.to_nchw()
    d = torch.where(f.to_nchw() > 0, d, d * slope)
    cimg = img.shape[1]
    out = torch.zeros((f.c, 4))
    out[:, :cimg] = torch.einsum("nchw,nihw->ci", d, img.float())
    out[:, 3] = d.sum(dim=(0, 2, 3))
    if weight is None:
        return out
    return out, torch.einsum("nchw,ci->nihw", d, weight.detach().view(f.c, cimg))


# ------------------------------------------------------------------------------------------------
# StyleGAN2 synthesis: forward pieces and the backward kernels
# ------------------------------------------------------------------------------------------------
_F4 = torch.tensor([0.25, 0.75, 0.75, 0.25])


def _fir_pad1(t):
    """out[y][x] = sum_{a,b<4} f[a] f[b] t[y+a-1][x+b-1] (zero outside): (2h+1)^2 -> (2h)^2  (stencil_tma.cu)."""
    c = t.shape[1]
    k = (_F4[:, None] * _F4[None, :]).expand(c, 1, 4, 4).contiguous()
    return F.conv2d(F.pad(t, (1, 1, 1, 1)), k, groups=c)


def sg2_prep_all(S, wp32, layers, outputs):
    styles, demods, rgb_styles, rgb_ws = [], [], [], []
    for j, m in enumerate(layers):
        st = m.style
        b = None if st.bias is None else st.bias.detach() * st.bscale
        s = F.linear(wp32[:, j], st.weight.detach() * st.wscale, b) + st.additional_bias
        styles.append(s)
        if m.demodulate:
            w2 = ((m.weight.detach() * m.wscale) ** 2).sum(dim=(2, 3))
            demods.append(torch.rsqrt((s * s) @ w2.t() + m.eps))
        else:
            demods.append(None)
    for k, m in enumerate(outputs):
        st = m.style
        b = None if st.bias is None else st.bias.detach() * st.bscale
        s = F.linear(wp32[:, 2 * k + 1], st.weight.detach() * st.wscale, b) + st.additional_bias
        rgb_styles.append(s)
        rgb_ws.append((m.weight.detach().view(m.out_c, m.in_c) * m.wscale)[None] * s[:, None, :])
    return styles, demods, rgb_styles, rgb_ws, {'styles': styles, 'demods': demods}


def sg2_prep_bwd(S, handle, layers, outputs, sums, layer_offs, const_off, n):
    """dge_sg2_prep_bwd restated with torch ops on the same arena (include/dge_b200.h)."""
    styles, demods = handle['styles'], handle['demods']
    d_wp = torch.zeros((n, S.num_layers, S.w_space_dim))

    def lsum(i):
        c = layers[i].out_c
        return sums[layer_offs[i]:layer_offs[i] + n * c * 5].view(n, c, 5)

    for i, m in enumerate(layers):
        ds = sums[const_off:const_off + n * m.in_c].view(n, m.in_c) if i == 0 else lsum(i - 1)[:, :, 0]
        if m.demodulate:
            w2 = ((m.weight.detach() * m.wscale) ** 2).sum(dim=(2, 3))
            ds = ds - styles[i] * ((lsum(i)[:, :, 4] * demods[i] * demods[i]) @ w2)
        d_wp[:, i] += ds @ (m.style.weight.detach() * m.style.wscale)
    for k, m in enumerate(outputs):
        w = m.weight.detach().view(m.out_c, m.in_c) * m.wscale
        ds = (lsum(2 * k)[:, :, 1:4] * w.t().unsqueeze(0)).sum(dim=2)
        d_wp[:, 2 * k + 1] += ds @ (m.style.weight.detach() * m.style.wscale)
    return d_wp


def up_fir_epilogue(raw_up, n, c, h_out, w_out, *, demod=None, noise=None, noise_batched=False, noise_scalar=0.0,
                    bias=None, slope=1.0, gain=1.0, out_scale=None, planes=2, out_act=True, out_nchw=False):
    y = _fir_pad1(_from_blocked(raw_up))
    if demod is not None:
        y = y * demod.view(n, c, 1, 1)
    if noise is not None:
        y = y + noise.reshape(n if noise_batched else 1, 1, h_out, w_out) * noise_scalar
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    y = torch.where(y < 0, y * slope, y) * gain
    res = {}
    if out_nchw:
        res["nchw"] = y
    if out_act:
        res["act"] = Act.of(y if out_scale is None else y * out_scale.view(n, c, 1, 1), planes)
    return res


def _up2_axis(x, dim):
    """out[2m] = (x[m-1] + 3 x[m]) / 4, out[2m+1] = (3 x[m] + x[m+1]) / 4 along `dim` (zero outside)."""
    n = x.shape[dim]
    pad = [0, 0] * (x.ndim - 1 - dim) + [1, 1]
    xp = F.pad(x, pad)
    a, b, c = xp.narrow(dim, 0, n), xp.narrow(dim, 1, n), xp.narrow(dim, 2, n)
    even, odd = (a + 3 * b) / 4, (3 * b + c) / 4
    return torch.stack((even, odd), dim=dim + 1).flatten(dim, dim + 1)


def rgb_init(img_in, bias, n, nch, h_out, w_out, device):
    out = torch.zeros((n, nch, h_out, w_out)) if bias is None else bias.view(1, nch, 1, 1).expand(n, nch, h_out, w_out).clone()
    if img_in is not None:
        out = out + _up2_axis(_up2_axis(img_in, 2), 3)
    return out


def rgb_up_bwd(d_out):
    n, ch, ho, wo = d_out.shape
    x = torch.zeros((n, ch, ho // 2, wo // 2), requires_grad=True)
    with torch.enable_grad():
        y = _up2_axis(_up2_axis(x, 2), 3)
    return torch.autograd.grad(y, x, d_out)[0]


def sg2_layer_bwd(ya, ya_scale, dxs, dimg, rgbw, noise, noise_batched, noise_scalar, bias, demod, gain, slope,
                  out_kind="act", planes=2, sums=None):
    n, c, h, w = ya.n, ya.c, ya.h, ya.w
    sums_out = sums
    y = ya.to_nchw()
    sn = torch.ones((n, c)) if ya_scale is None else ya_scale
    isn = torch.where(sn != 0, 1.0 / sn, torch.zeros_like(sn))
    y = y * isn.view(n, c, 1, 1)
    dx = torch.zeros_like(y) if dxs is None else dxs.to_nchw()
    dy = dx * sn.view(n, c, 1, 1)
    sums = torch.zeros((n, c, 5))
    sums[:, :, 0] = (dx * y).sum(dim=(2, 3))
    if dimg is not None:
        dy = dy + torch.einsum("nkc,nkhw->nchw", rgbw, dimg)
        sums[:, :, 1:4] = torch.einsum("nkhw,nchw->nck", dimg, y)
    pos = y > 0
    dpre = dy * torch.where(pos, gain, gain * slope)
    pre = y * torch.where(pos, 1.0 / gain, 1.0 / (gain * slope) if slope != 0 else 0.0)
    nz = 0.0 if noise is None else noise.reshape(n if noise_batched else 1, 1, h, w) * noise_scalar
    b = 0.0 if bias is None else bias.view(1, c, 1, 1)
    sums[:, :, 4] = (dpre * (pre - nz - b)).sum(dim=(2, 3))
    v = dpre if demod is None else dpre * demod.view(n, c, 1, 1)
    if sums_out is not None:
        sums_out.copy_(sums)
        sums = sums_out
    return (Act.of(v, planes) if out_kind == "act" else F32B.of(v)), sums


def _box2(t):
    """out[y][x] = sum_{a,b<2} t[y+a][x+b]: (2h+1)^2 -> (2h)^2  (dge_sg1_post mode 1)."""
    return t[:, :, :-1, :-1] + t[:, :, 1:, :-1] + t[:, :, :-1, 1:] + t[:, :, 1:, 1:]


def up_fir_bwd_s2d(dconv, planes=2, box=False):
    d = dconv.to_nchw()
    n, c, ho, wo = d.shape
    t = torch.zeros((n, c, ho + 1, wo + 1), requires_grad=True)
    with torch.enable_grad():
        y = _box2(t) if box else _fir_pad1(t)
    dt = torch.autograd.grad(y, t, d)[0]
    dt = F.pad(dt, (0, 1, 0, 1))                                            # (2h+2) x (2w+2), zeros beyond the raw map
    hs, ws = ho // 2 + 1, wo // 2 + 1
    v = dt.view(n, c, hs, 2, ws, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, hs, ws)   # channel = (2py+px)*c + ch
    return Act.of(v, planes)


# ------------------------------------------------------------------------------------------------
# StyleGAN1 pieces (csrc/elementwise.cu: k_sg1_post, k_instance_norm_style, k_to_rgb_f32b)
# ------------------------------------------------------------------------------------------------
def _blur3(x):
    c = x.shape[1]
    f = torch.tensor([0.25, 0.5, 0.25])
    k = (f[:, None] * f[None, :]).expand(c, 1, 3, 3).contiguous()
    return F.conv2d(x, k, padding=1, groups=c)


def sg1_post(src, mode, n, c, h_out, w_out, noise=None, noise_w=None, bias=None, slope=0.2):
    if mode == 1:
        v = _blur3(_box2(_from_blocked(src)))
    elif mode == 0:
        v = _blur3(src.to_nchw())
    else:
        v = src.to_nchw()
    if noise is not None:
        v = v + noise.reshape(n, 1, h_out, w_out) * noise_w.view(1, -1, 1, 1)
    if bias is not None:
        v = v + bias.view(1, -1, 1, 1)
    return F32B.of(torch.where(v < 0, v * slope, v))


def instance_norm_style(x, mean_rstd, style, n, up=1, planes=2, out_act=True, out_f32b=False):
    c = x.c
    y = _in_apply(x, mean_rstd)
    if y.shape[0] == 1 and n > 1:
        y = y.expand(n, -1, -1, -1)
    if style is not None:
        y = y * (style[:, :c, None, None] + 1) + style[:, c:, None, None]
    if up == 2:
        y = y.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    return (Act.of(y, planes) if out_act else None), (F32B.of(y) if out_f32b else None)


def to_rgb_f32b(x, w, bias):
    return F.conv2d(x.to_nchw(), w.detach().float().view(w.shape[0], -1, 1, 1), None if bias is None else bias.detach().float())
