"""Plain-torch (CPU) statement of the C-ABI functions the fused training path calls (include/dge_b200.h), on the kernels'
own layouts (ACT bf16 [N][C/8][planes][H][W][8], F32B fp32 [N][C/8][H][W][8]).  TEST INFRASTRUCTURE: it is
  * the executable specification every new CUDA kernel is compared with on the GPU (tests/test_train_kernels_gpu.py),
  * a stand-in for `dge_b200.ops` so that the chain rule coded in dge_b200/train_e.py / train_g.py can be checked against
    gradients of the unmodified reference without a GPU (tests/test_train_fused_cpu.py).
The product never imports this module.
"""
import torch
import torch.nn.functional as F

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
         out_f32b_pool=False, out_nchw=False, rgb_w=None, rgb_out=None, checker=False, in_hw=None):
    assert isinstance(x, Act) and isinstance(wpk, _Packed) and preact_add is None and out_f32b_into is None
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
        hh, ww = in_hw if in_hw is not None else (h, w_)

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
                 slope=0.2, planes=2):
    gv, xv = g.to_nchw(), x.to_nchw()
    n, c, h, w = xv.shape
    hw = h * w
    m, r = mean_rstd[:, :, 0, None, None], mean_rstd[:, :, 1, None, None]
    a = (sums[:, :, 0] / hw).float()[:, :, None, None]
    b = (sums[:, :, 1] / hw).float()[:, :, None, None]
    xc = xv - m
    v = r * (gv - a - xc * r * b)
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
    return Act.of(v, planes), torch.stack((v.sum(dim=(0, 2, 3)), (v * nz).sum(dim=(0, 2, 3))))


def from_rgb_bwd(d_f, f, img, slope=0.2):
    d = d_f.to_nchw()
    d = torch.where(f.to_nchw() > 0, d, d * slope)
    cimg = img.shape[1]
    out = torch.zeros((f.c, 4))
    out[:, :cimg] = torch.einsum("nchw,nihw->ci", d, img.float())
    out[:, 3] = d.sum(dim=(0, 2, 3))
    return out
