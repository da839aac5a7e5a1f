"""Tensor-level wrappers over the C ABI (include/dge_b200.h).

PyTorch is used only for device memory and streams; every function launches hand-written sm_100a
kernels on torch's current CUDA stream and raises if the library / device is unavailable.
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvArgs, DgeError, check

CONV_3X3, CONV_1X1, CONV_UP3X3, CONV_DOWN4X4S2 = 0, 1, 2, 3
FLAG_CHECKER = 1

_device_checked = False

# ------------------------------------------------------------------------------------------------
# weight-cache epoch
# ------------------------------------------------------------------------------------------------
# The modules cache tensors DERIVED from their parameters (packed WPK weights, demodulation Gram diagonals, ...).
# torch's per-tensor `_version` only sees in-place ops on the Parameter object itself: an optimiser that writes through
# `p.data` (the reference's LREQAdam: `p.data.addcdiv_`, custom_adam.py:74) or through a raw device pointer
# (dge_lreq_adam_step) changes the values without bumping it.  Every cache key therefore also carries this epoch, which
# is advanced by EVERY torch optimiser step (global post-step hook below) and by LREQAdam.step itself; code that
# mutates `p.data` by hand outside an optimiser calls `invalidate_weight_caches()`.
_weights_epoch = 0          # global part (invalidate_weight_caches() without arguments)
_param_epoch = {}           # per-storage part: data_ptr -> epoch, advanced for the parameters an optimiser just updated


def weights_epoch():
    return _weights_epoch


def invalidate_weight_caches(params=None):
    """Retire cached tensors derived from `params` (an iterable of tensors), or from every parameter when None.  An
    optimiser step only retires what it updated: the frozen generator keeps its packed weights while the encoder trains."""
    global _weights_epoch
    if params is None:
        _weights_epoch += 1
        return
    for p in params:
        k = p.data_ptr()
        _param_epoch[k] = _param_epoch.get(k, 0) + 1


def weight_key(*tensors):
    """Cache key of tensors derived from these parameters: storage, in-place version and the optimiser epochs."""
    return (_weights_epoch,) + tuple((t.data_ptr(), t._version, _param_epoch.get(t.data_ptr(), 0), t.device.index)
                                     for t in tensors)


def _optimizer_stepped(optimizer, *_a, **_k):
    invalidate_weight_caches(p for g in optimizer.param_groups for p in g['params'])


try:  # any torch.optim.Optimizer subclass, including the unmodified reference LREQAdam
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook
    _reg_post_hook(_optimizer_stepped)
except ImportError:  # pragma: no cover - torch < 2.1
    pass


def lib():
    global _device_checked
    L = _lib.load()
    if not _device_checked:
        if not torch.cuda.is_available():
            raise DgeError("dge_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
        check(L.dge_device_ok())
        _device_checked = True
    return L


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """torch's current CUDA stream as a raw cudaStream_t (the C entry points: ~0.3 us instead of ~15 us through
    torch.cuda.current_stream(), which matters at ~600 launches per training iteration)."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())              # a plain int: ctypes converts it for the `void*` parameter
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------
# optional per-launch timing (bench.py roofline leg): CUDA events on the launching stream
# ------------------------------------------------------------------------------------------------
_prof = None


class profile:
    """with ops.profile() as rec: ...  -> rec.events: list of (kernel, key, start_event, end_event)."""

    def __enter__(self):
        global _prof
        self.events = []
        _prof = self.events
        return self

    def __exit__(self, *exc):
        global _prof
        _prof = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, key, e0, e1 in self.events:
            d = out.setdefault((name, key), [0, 0.0])
            d[0] += 1
            d[1] += e0.elapsed_time(e1)
        return out


class _Rec:
    def __init__(self, name, key):
        self.name, self.key = name, key

    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e0.record()

    def __exit__(self, *exc):
        if _prof is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.name, self.key, self.e0, e1))


class _NoRec:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOREC = _NoRec()


def _rec(name, key):
    """Per-launch timing scope: a shared no-op object unless `ops.profile()` is active (~700 launches per training
    iteration go through here; at batch 1 the loop is host-bound)."""
    return _NOREC if _prof is None else _Rec(name, key)


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "dge_b200 ops need contiguous CUDA tensors"
    return t.data_ptr()                                # a plain int: the bindings declare `void*` argtypes


def _f32(t):
    assert t is None or t.dtype == torch.float32
    return _p(t)


# ------------------------------------------------------------------------------------------------
# layout containers
# ------------------------------------------------------------------------------------------------
class Act:
    """ACT layout: bf16 [N][C/8][planes][H][W][8] (planes=2: x = hi + lo)."""

    def __init__(self, n, c, h, w, planes=2, device="cuda"):
        assert c % 16 == 0, f"ACT tensors need C % 16 == 0 (got {c})"
        self.n, self.c, self.h, self.w, self.planes = n, c, h, w, planes
        self.t = torch.empty((n, c // 8, planes, h, w, 8), dtype=torch.bfloat16, device=device)

    @classmethod
    def wrap(cls, t, n, c, h, w, planes=2):
        """View an existing bf16 buffer of the right size as an ACT tensor (saved-for-backward storage)."""
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() == n * c * planes * h * w
        o = cls.__new__(cls)
        o.n, o.c, o.h, o.w, o.planes, o.t = n, c, h, w, planes, t
        return o

    def to_nchw(self):
        out = torch.empty((self.n, self.c, self.h, self.w), dtype=torch.float32, device=self.t.device)
        check(lib().dge_act_to_nchw(_p(self.t), _p(out), self.n, self.c, self.h, self.w, self.planes, _stream()))
        return out


class F32B:
    """F32B layout: fp32 [N][C/8][H][W][8]."""

    def __init__(self, n, c, h, w, device="cuda"):
        assert c % 8 == 0
        self.n, self.c, self.h, self.w = n, c, h, w
        self.t = torch.empty((n, c // 8, h, w, 8), dtype=torch.float32, device=device)

    @classmethod
    def wrap(cls, t, n, c, h, w):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n * c * h * w
        o = cls.__new__(cls)
        o.n, o.c, o.h, o.w, o.t = n, c, h, w, t
        return o

    def to_nchw(self):
        out = torch.empty((self.n, self.c, self.h, self.w), dtype=torch.float32, device=self.t.device)
        check(lib().dge_f32b_to_nchw(_p(self.t), _p(out), self.n, self.c, self.h, self.w, _stream()))
        return out


def nchw_to_act(x, scale=None, planes=2, batch=None):
    """NCHW fp32 -> ACT. `batch` > x.shape[0]==1 broadcasts the single sample (const input)."""
    x = x.contiguous()
    n0, c, h, w = x.shape
    n = batch if batch is not None else n0
    bstride = 0 if (batch is not None and n0 == 1) else c * h * w
    out = Act(n, c, h, w, planes, x.device)
    check(lib().dge_nchw_to_act(_f32(x), bstride, _f32(scale), _p(out.t), n, c, h, w, planes, _stream()))
    return out


def nchw_to_f32b(x):
    x = x.contiguous()
    n, c, h, w = x.shape
    out = F32B(n, c, h, w, x.device)
    check(lib().dge_nchw_to_f32b(_f32(x), _p(out.t), n, c, h, w, _stream()))
    return out


def f32b_to_act(x, planes=2):
    assert isinstance(x, F32B)
    out = Act(x.n, x.c, x.h, x.w, planes, x.t.device)
    check(lib().dge_f32b_to_act(_p(x.t), _p(out.t), x.n, x.c, x.h, x.w, planes, _stream()))
    return out


def to_rgb_nchw(x, rgb_w, bias):
    x = x.contiguous()
    n, c, h, w = x.shape
    nch = rgb_w.shape[1]
    out = torch.empty((n, nch, h, w), dtype=torch.float32, device=x.device)
    bb = None if bias is None else bias.detach().contiguous()
    check(lib().dge_to_rgb_nchw(_f32(x), _f32(rgb_w), _f32(bb), _p(out), n, c, nch, h, w, _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------------
def pack_conv_weight(w, scale=1.0, flip=False, planes=2):
    """OIHW fp32 -> WPK bf16 [taps][Cin/8][planes][Cout][8]."""
    w = w.detach().contiguous()
    cout, cin, k, k2 = w.shape
    assert k == k2
    out = torch.empty((k * k, cin // 8, planes, cout, 8), dtype=torch.bfloat16, device=w.device)
    check(lib().dge_pack_conv_weight(_f32(w), _p(out), cout, cin, k, int(flip), float(scale), planes, _stream()))
    return out


def pack_conv_weight_dgrad(w, scale=1.0, planes=2):
    """OIHW fp32 -> the data-gradient operand (transposed + flipped): conv(dy_act, this, cin) = dL/dx of conv2d(x, w)."""
    w = w.detach().contiguous()
    cout, cin, k, k2 = w.shape
    assert k == k2
    out = torch.empty((k * k, cout // 8, planes, cin, 8), dtype=torch.bfloat16, device=w.device)
    check(lib().dge_pack_conv_weight_dgrad(_f32(w), _p(out), cout, cin, k, float(scale), planes, _stream()))
    return out


def conv_wgrad(dy, x, ksize, out=None, accumulate=False):
    """dL/dW of conv2d(x, W, padding=ksize//2): `dy` and `x` are Act tensors (same batch, size and planes); returns
    the fp32 [cout, cin, k, k] gradient (added into `out` when accumulate=True)."""
    assert isinstance(dy, Act) and isinstance(x, Act)
    assert (dy.n, dy.h, dy.w, dy.planes) == (x.n, x.h, x.w, x.planes), "conv_wgrad: dy / x geometry mismatch"
    if out is None:
        assert not accumulate
        out = torch.empty((dy.c, x.c, ksize, ksize), dtype=torch.float32, device=x.t.device)
    assert out.shape == (dy.c, x.c, ksize, ksize) and out.dtype == torch.float32 and out.is_contiguous()
    check(lib().dge_conv_wgrad(_p(dy.t), _p(x.t), _f32(out), x.n, dy.c, x.c, x.h, x.w, ksize, x.planes,
                               1 if accumulate else 0, _stream()))
    return out


def weight_sqsum(w, scale=1.0):
    w = w.detach().contiguous()
    cout, cin, k, _ = w.shape
    out = torch.empty((cout, cin), dtype=torch.float32, device=w.device)
    check(lib().dge_weight_sqsum(_f32(w), _p(out), cout, cin, k, float(scale), _stream()))
    return out


def demod(w2, style, eps=1e-8):
    n, cin = style.shape
    cout = w2.shape[0]
    out = torch.empty((n, cout), dtype=torch.float32, device=style.device)
    check(lib().dge_demod(_f32(w2), _f32(style.contiguous()), _p(out), n, cout, cin, float(eps), _stream()))
    return out


def rgb_weights(w, style, scale):
    """w [nch][cin] (1x1 ToRGB weight), style [n][cin] -> [n][nch][cin]."""
    w = w.detach().contiguous().view(w.shape[0], -1)
    nch, cin = w.shape
    n = style.shape[0]
    out = torch.empty((n, nch, cin), dtype=torch.float32, device=style.device)
    check(lib().dge_rgb_weights(_f32(w), _f32(style.contiguous()), _p(out), n, nch, cin, float(scale), _stream()))
    return out


def sg2_prep(items_dev, n_items, wp, arena, n, num_layers, wdim):
    """All style affines + demod coefficients + ToRGB weights of one synthesis pass in one launch (dge_sg2_prep)."""
    with _rec("sg2_prep", (n, n_items)):
        check(lib().dge_sg2_prep(_p(items_dev), n_items, _f32(wp), _p(arena), n, num_layers, wdim, _stream()))


def dense(x, w, b=None, wscale=1.0, bscale=1.0, add_bias=0.0, slope=1.0, gain=1.0):
    x = x.contiguous()
    w = w.detach().contiguous()
    n, k = x.shape
    m = w.shape[0]
    assert w.shape[1] == k
    y = torch.empty((n, m), dtype=torch.float32, device=x.device)
    bb = None if b is None else b.detach().contiguous()
    check(lib().dge_dense(_f32(x), _f32(w), _f32(bb), _p(y), n, k, m, float(wscale), float(bscale), float(add_bias),
                          float(slope), float(gain), _stream()))
    return y


def pixel_norm(x, eps=1e-8):
    x = x.contiguous()
    y = torch.empty_like(x)
    check(lib().dge_pixel_norm(_f32(x), _p(y), x.shape[0], x.shape[1], float(eps), _stream()))
    return y


# ------------------------------------------------------------------------------------------------
# conv
# ------------------------------------------------------------------------------------------------
_splitk_ws = {}


def conv(x, wpk, cout, kind=CONV_3X3, *, demod=None, noise=None, noise_batched=False, noise_w=None,
         noise_scalar=0.0, bias=None, slope=1.0, gain=1.0, blend_src=None, blend_pool=False, blend_a=0.0,
         blend_b=1.0, preact_add=None, preact_up=1, out_act=False, out_planes=None, out_scale=None, out_f32b=False,
         out_f32b_into=None, out_f32b_pool=False, out_nchw=False, rgb_w=None, rgb_out=None, checker=False,
         out_hw=None):
    """tcgen05 implicit-GEMM conv with fused epilogue (dge_conv_forward). Returns a dict of outputs.
    `out_hw` = (h, w) smaller than x's stored extent: only that top-left output domain is computed (args in_h / in_w)."""
    assert isinstance(x, Act)
    dev = x.t.device
    a = ConvArgs()
    a.kind, a.flags = kind, (FLAG_CHECKER if checker else 0)
    a.n, a.h, a.w, a.cin, a.cout, a.planes = x.n, x.h, x.w, x.c, cout, x.planes
    if out_hw is not None and tuple(out_hw) != (x.h, x.w):
        a.h, a.w, a.in_h, a.in_w = int(out_hw[0]), int(out_hw[1]), x.h, x.w
    oh, ow = a.h, a.w
    a.x, a.wpk = x.t.data_ptr(), wpk.data_ptr()
    keep = [x.t, wpk]
    res = {}

    def ptr(t):
        if t is None:
            return None
        assert t.is_cuda and t.is_contiguous()
        keep.append(t)
        return t.data_ptr()

    if kind == CONV_UP3X3:
        raw = torch.empty((x.n, cout // 8, 2 * x.h + 1, 2 * x.w + 1, 8), dtype=torch.float32, device=dev)
        a.out_raw_up = ptr(raw)
        res["raw_up"] = raw
    else:
        a.demod = ptr(demod)
        a.noise = ptr(noise)
        a.noise_bstride = oh * ow if (noise is not None and noise_batched) else 0
        a.noise_w = ptr(noise_w)
        a.noise_scalar = float(noise_scalar)
        a.bias = ptr(bias)
        a.slope, a.gain = float(slope), float(gain)
        if preact_add is not None:
            a.preact_add = ptr(preact_add.t if isinstance(preact_add, F32B) else preact_add)
            a.preact_c = preact_add.c if isinstance(preact_add, F32B) else 0
            a.preact_up = int(preact_up)
        if blend_src is not None:
            a.blend_src = ptr(blend_src.t if isinstance(blend_src, F32B) else blend_src)
            a.blend_pool, a.blend_a, a.blend_b = int(blend_pool), float(blend_a), float(blend_b)
        if out_act:
            o = Act(x.n, cout, oh, ow, out_planes or x.planes, dev)
            a.out_act, a.out_planes = ptr(o.t), o.planes
            a.out_scale = ptr(out_scale)
            res["act"] = o
        if out_f32b:
            o = F32B(x.n, cout, oh, ow, dev)
            a.out_f32b = ptr(o.t)
            res["f32b"] = o
        if out_f32b_pool:                    # 2x2 mean of the epilogue value at half resolution (E.py:78-84)
            o = F32B(x.n, cout, oh // 2, ow // 2, dev)
            a.out_f32b, a.out_pool = ptr(o.t), 1
            res["f32b_pool"] = o
        if out_f32b_into is not None:        # write into a caller-provided F32B slice (e.g. one sample of a batch)
            assert out_f32b_into.numel() == x.n * cout * oh * ow and out_f32b_into.dtype == torch.float32
            a.out_f32b = ptr(out_f32b_into)
        if out_nchw:
            o = torch.empty((x.n, cout, oh, ow), dtype=torch.float32, device=dev)
            a.out_nchw = ptr(o)
            res["nchw"] = o
        if rgb_w is not None:
            a.rgb_w, a.rgb_out = ptr(rgb_w), ptr(rgb_out)
    if not checker:
        wkey = (kind, a.n, a.h, a.w, a.cin, a.cout)          # what the library's split-K decision reads (conv_mma.cu)
        wsb = _splitk_ws.get(wkey)
        if wsb is None:
            wsb = _splitk_ws[wkey] = int(lib().dge_conv_splitk_ws_bytes(ctypes.byref(a)))
        if wsb:                              # small maps: split-K scratch (zeroed by the call)
            ws = torch.empty((wsb // 4,), dtype=torch.float32, device=dev)
            a.splitk_ws = ptr(ws)
    name = ("conv3x3", "conv1x1", "conv_up3x3", "conv_down4x4s2")[kind]
    with _rec(name, (x.n, oh, ow, x.c, cout, x.planes)):
        check(lib().dge_conv_forward(ctypes.byref(a), _stream()))
    return res


def up_fir_epilogue(raw_up, n, c, h_out, w_out, *, demod=None, noise=None, noise_batched=False, noise_scalar=0.0,
                    bias=None, slope=1.0, gain=1.0, out_scale=None, planes=2, out_act=True, out_nchw=False):
    dev = raw_up.device
    res = {}
    act = Act(n, c, h_out, w_out, planes, dev) if out_act else None
    nchw = torch.empty((n, c, h_out, w_out), dtype=torch.float32, device=dev) if out_nchw else None
    with _rec("up_fir_epilogue", (n, h_out, w_out, c, planes)):
        check(lib().dge_up_fir_epilogue(
            _f32(raw_up), _f32(demod), _f32(noise), (h_out * w_out if noise_batched else 0), float(noise_scalar),
            _f32(bias), float(slope), float(gain), _f32(out_scale), _p(act.t) if act else None, _p(nchw), n, c, h_out,
            w_out, planes, _stream()))
    if act is not None:
        res["act"] = act
    if nchw is not None:
        res["nchw"] = nchw
    return res


def rgb_init(img_in, bias, n, nch, h_out, w_out, device):
    out = torch.empty((n, nch, h_out, w_out), dtype=torch.float32, device=device)
    with _rec("rgb_init", (n, h_out, w_out)):
        check(lib().dge_rgb_init(_f32(img_in), _f32(bias), _p(out), n, nch, h_out, w_out, _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# encoder pieces
# ------------------------------------------------------------------------------------------------
def from_rgb(img, w, b, slope=0.2):
    img = img.contiguous()
    n, cimg, h, wd = img.shape
    w2 = w.detach().contiguous().view(w.shape[0], -1)
    c = w2.shape[0]
    out = F32B(n, c, h, wd, img.device)
    bb = None if b is None else b.detach().contiguous()
    with _rec("from_rgb", (n, h, wd, c)):
        check(lib().dge_from_rgb(_f32(img), _f32(w2), _f32(bb), _p(out.t), n, cimg, c, h, wd, float(slope), _stream()))
    return out


def from_rgb_stats(img, w, b, slope=0.2, eps=1e-8):
    """FromRGB + the instance statistics of its output in one pass -> (F32B, style, mean_rstd); c in (16, 32)."""
    img = img.contiguous()
    n, cimg, h, wd = img.shape
    w2 = w.detach().contiguous().view(w.shape[0], -1)
    c = w2.shape[0]
    dev = img.device
    out = F32B(n, c, h, wd, dev)
    scratch = torch.empty((2 * n * c,), dtype=torch.float64, device=dev)
    style = torch.empty((n, 2 * c), dtype=torch.float32, device=dev)
    mr = torch.empty((n, c, 2), dtype=torch.float32, device=dev)
    bb = None if b is None else b.detach().contiguous()
    with _rec("from_rgb_stats", (n, h, wd, c)):
        check(lib().dge_from_rgb_stats(_f32(img), _f32(w2), _f32(bb), _p(out.t), _p(scratch), _p(style), _p(mr), n, cimg,
                                       c, h, wd, float(slope), float(eps), _stream()))
    return out, style, mr


def from_rgb_stats_any(img, w, b, slope=0.2, eps=1e-8):
    """FromRGB + instance statistics for any width: the one-pass kernel for c in (16, 32), else two kernels."""
    if w.shape[0] in (16, 32):
        return from_rgb_stats(img, w, b, slope, eps)
    f = from_rgb(img, w, b, slope)
    return (f,) + instance_stats(f, eps)


def scale_f32b(x, a, to_act=False, planes=2):
    """a * x on an F32B map (tiny maps of the last encoder block), optionally as an ACT operand."""
    assert isinstance(x, F32B)
    y = F32B.wrap(x.t * float(a), x.n, x.c, x.h, x.w)
    return f32b_to_act(y, planes) if to_act else y


def f32b_channel_sums(x):
    """[c] = sum over (n, h, w) of an F32B map (tiny maps only: a torch reduction)."""
    assert isinstance(x, F32B)
    return x.t.sum(dim=(0, 2, 3)).reshape(-1)


def instance_norm_pool(x, mean_rstd, planes=2):
    """IN(x) -> ACT and avg_pool2d(x, 2, 2) -> ACT in one pass over x."""
    assert isinstance(x, F32B)
    dev = x.t.device
    act = Act(x.n, x.c, x.h, x.w, planes, dev)
    pool = Act(x.n, x.c, x.h // 2, x.w // 2, planes, dev)
    with _rec("instance_norm_pool", (x.n, x.h, x.w, x.c, planes)):
        check(lib().dge_instance_norm_pool(_p(x.t), _f32(mean_rstd), _p(act.t), _p(pool.t), x.n, x.c, x.h, x.w, planes,
                                           _stream()))
    return act, pool


def instance_stats(x, eps=1e-8):
    """F32B -> (style [n][2c] = mean||std, mean_rstd [n][c][2])."""
    assert isinstance(x, F32B)
    dev = x.t.device
    scratch = torch.empty((2 * x.n * x.c,), dtype=torch.float64, device=dev)
    style = torch.empty((x.n, 2 * x.c), dtype=torch.float32, device=dev)
    mr = torch.empty((x.n, x.c, 2), dtype=torch.float32, device=dev)
    with _rec("instance_stats", (x.n, x.h, x.w, x.c)):
        check(lib().dge_instance_stats(_p(x.t), _p(scratch), _p(style), _p(mr), x.n, x.c, x.h, x.w, float(eps),
                                       _stream()))
    return style, mr


def instance_norm(x, mean_rstd, planes=2, out_act=True, out_f32b=False, gamma=None, beta=None):
    assert isinstance(x, F32B)
    dev = x.t.device
    act = Act(x.n, x.c, x.h, x.w, planes, dev) if out_act else None
    f = F32B(x.n, x.c, x.h, x.w, dev) if out_f32b else None
    g = None if gamma is None else gamma.detach().contiguous()
    b = None if beta is None else beta.detach().contiguous()
    with _rec("instance_norm", (x.n, x.h, x.w, x.c, planes)):
        check(lib().dge_instance_norm_affine(_p(x.t), _f32(mean_rstd), _f32(g), _f32(b), _p(act.t) if act else None,
                                             _p(f.t) if f else None, x.n, x.c, x.h, x.w, planes, _stream()))
    return act, f


def sg1_post(src, mode, n, c, h_out, w_out, noise=None, noise_w=None, bias=None, slope=0.2):
    """src: F32B (modes 0, 2) or the raw_up tensor of a CONV_UP3X3 launch (mode 1)."""
    t = src.t if isinstance(src, F32B) else src
    out = F32B(n, c, h_out, w_out, t.device)
    with _rec("sg1_post", (n, h_out, w_out, c, mode)):
        check(lib().dge_sg1_post(_f32(t), mode, _f32(noise), _f32(noise_w), _f32(bias), float(slope), _p(out.t), n, c,
                                 h_out, w_out, _stream()))
    return out


def instance_norm_style(x, mean_rstd, style, n, up=1, planes=2, out_act=True, out_f32b=False):
    assert isinstance(x, F32B)
    dev = x.t.device
    act = Act(n, x.c, x.h * up, x.w * up, planes, dev) if out_act else None
    f = F32B(n, x.c, x.h * up, x.w * up, dev) if out_f32b else None
    st = None if style is None else style.contiguous()
    with _rec("instance_norm_style", (n, x.h, x.w, x.c, up)):
        check(lib().dge_instance_norm_style(_p(x.t), x.n, _f32(mean_rstd), _f32(st), up, _p(act.t) if act else None,
                                            _p(f.t) if f else None, n, x.c, x.h, x.w, planes, _stream()))
    return act, f


def to_rgb_f32b(x, w, bias):
    assert isinstance(x, F32B)
    w2 = w.detach().contiguous().view(w.shape[0], -1)
    out = torch.empty((x.n, w2.shape[0], x.h, x.w), dtype=torch.float32, device=x.t.device)
    bb = None if bias is None else bias.detach().contiguous()
    check(lib().dge_to_rgb_f32b(_p(x.t), _f32(w2), _f32(bb), _p(out), x.n, x.c, w2.shape[0], x.h, x.w, _stream()))
    return out


def cbn_coeffs(mean, var, eps, n, scale=None, offset=None, weight=None, bias=None):
    c = mean.numel()
    dev = mean.device
    a = torch.empty((n, c), dtype=torch.float32, device=dev)
    b = torch.empty((n, c), dtype=torch.float32, device=dev)
    cg = lambda t: None if t is None else t.detach().contiguous()
    check(lib().dge_cbn_coeffs(_f32(cg(scale)), _f32(cg(offset)), _f32(cg(weight)), _f32(cg(bias)), _f32(cg(mean)),
                               _f32(cg(var)), float(eps), _p(a), _p(b), n, c, _stream()))
    return a, b


def affine_act(x, a, b, relu=False, up=1, planes=2, out_act=True, out_f32b=False):
    assert isinstance(x, F32B)
    dev = x.t.device
    act = Act(x.n, x.c, x.h * up, x.w * up, planes, dev) if out_act else None
    f = F32B(x.n, x.c, x.h * up, x.w * up, dev) if out_f32b else None
    with _rec("affine_act", (x.n, x.h, x.w, x.c, up)):
        check(lib().dge_affine_act(_p(x.t), _f32(a), _f32(b), int(relu), up, _p(act.t) if act else None,
                                   _p(f.t) if f else None, x.n, x.c, x.h, x.w, planes, _stream()))
    return act, f


def maxpool2(x):
    assert isinstance(x, F32B)
    out = F32B(x.n, x.c, x.h // 2, x.w // 2, x.t.device)
    check(lib().dge_maxpool2_f32b(_p(x.t), _p(out.t), x.n, x.c, x.h // 2, x.w // 2, _stream()))
    return out


def channel_softmax_to_act(x, planes=2):
    assert isinstance(x, F32B)
    out = Act(x.n, x.c, x.h, x.w, planes, x.t.device)
    check(lib().dge_channel_softmax_to_act(_p(x.t), _p(out.t), x.n, x.c, x.h, x.w, planes, _stream()))
    return out


def tanh_slice_nchw(x, nch):
    x = x.contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n, nch, h, w), dtype=torch.float32, device=x.device)
    check(lib().dge_tanh_slice_nchw(_f32(x), _p(out), n, c, nch, h * w, _stream()))
    return out


def pixelnorm_to_act(x, up=1, eps=1e-8, planes=2):
    assert isinstance(x, F32B)
    out = Act(x.n, x.c, x.h * up, x.w * up, planes, x.t.device)
    with _rec("pixelnorm_to_act", (x.n, x.h, x.w, x.c, up)):
        check(lib().dge_pixelnorm_to_act(_p(x.t), _p(out.t), x.n, x.c, x.h, x.w, up, float(eps), planes, _stream()))
    return out


def pixelnorm_to_rgb(x, w, bias, eps=1e-8):
    assert isinstance(x, F32B)
    w2 = w.contiguous().view(w.shape[0], -1)
    out = torch.empty((x.n, w2.shape[0], x.h, x.w), dtype=torch.float32, device=x.t.device)
    bb = None if bias is None else bias.detach().contiguous()
    check(lib().dge_pixelnorm_to_rgb(_p(x.t), _f32(w2), _f32(bb), _p(out), x.n, x.c, w2.shape[0], x.h, x.w, float(eps),
                                     _stream()))
    return out


def upsample_nearest_nchw(x):
    x = x.contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n, c, 2 * h, 2 * w), dtype=torch.float32, device=x.device)
    check(lib().dge_upsample_nearest_nchw(_f32(x), _p(out), n * c, h, w, _stream()))
    return out


def axpby(x, y, a, b):
    x, y = x.contiguous(), y.contiguous()
    out = torch.empty_like(x)
    check(lib().dge_axpby(_f32(x), _f32(y), _p(out), float(a), float(b), x.numel(), _stream()))
    return out


def avgpool_to_act(x, planes=2):
    assert isinstance(x, F32B)
    out = Act(x.n, x.c, x.h // 2, x.w // 2, planes, x.t.device)
    with _rec("avgpool_to_act", (x.n, x.h, x.w, x.c, planes)):
        check(lib().dge_avgpool_to_act(_p(x.t), _p(out.t), x.n, x.c, x.h, x.w, planes, _stream()))
    return out


def instance_norm_blur(x, mean_rstd, s2d=False, planes=2):
    """IN + 3x3 Blur -> ACT; s2d=True writes the space-to-depth operand of CONV_DOWN4X4S2 (4C channels, half size)."""
    assert isinstance(x, F32B)
    out = Act(x.n, 4 * x.c, x.h // 2, x.w // 2, planes, x.t.device) if s2d else Act(x.n, x.c, x.h, x.w, planes, x.t.device)
    with _rec("instance_norm_blur", (x.n, x.h, x.w, x.c, int(s2d))):
        check(lib().dge_instance_norm_blur(_p(x.t), _f32(mean_rstd), _p(out.t), int(s2d), x.n, x.c, x.h, x.w, planes,
                                           _stream()))
    return out


def blend(a_src, b_src, a, b, pool):
    """pool: False/0 none, True/3 both inputs are double resolution, 1 only a_src, 2 only b_src."""
    assert isinstance(a_src, F32B) and isinstance(b_src, F32B)
    pool = 3 if pool is True else int(pool)
    ho, wo = (a_src.h // 2, a_src.w // 2) if (pool & 1) else (a_src.h, a_src.w)
    out = F32B(a_src.n, a_src.c, ho, wo, a_src.t.device)
    with _rec("blend", (a_src.n, ho, wo, a_src.c, int(pool))):
        check(lib().dge_blend(_p(a_src.t), _p(b_src.t), _p(out.t), float(a), float(b), int(pool), a_src.n, a_src.c, ho,
                              wo, _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# training step: point-wise / reduction backward kernels (csrc/train_bwd.cu)
# ------------------------------------------------------------------------------------------------
def be_head_bwd(d_out, y2, noise, ga, gb, slope, want_dres=True, planes=2):
    """Backward of an encoder block's tail (E.py:72-84) -> (dy2 Act, dres Act | None, sums fp32 [3, co])."""
    assert isinstance(d_out, F32B) and isinstance(y2, F32B)
    assert (y2.n, y2.c, y2.h, y2.w) == (d_out.n, d_out.c, 2 * d_out.h, 2 * d_out.w), "be_head_bwd: geometry mismatch"
    dev = y2.t.device
    dy2 = Act(y2.n, y2.c, y2.h, y2.w, planes, dev)
    dres = Act(d_out.n, d_out.c, d_out.h, d_out.w, planes, dev) if want_dres else None
    sums = torch.empty((3, y2.c), dtype=torch.float32, device=dev)
    with _rec("be_head_bwd", (y2.n, y2.h, y2.w, y2.c, planes)):
        check(lib().dge_be_head_bwd(_p(d_out.t), _p(y2.t), _f32(noise), float(ga), float(gb), float(slope), _p(dy2.t),
                                    _p(dres.t) if dres is not None else None, _p(sums), y2.n, y2.c, y2.h, y2.w, planes,
                                    _stream()))
    return dy2, dres, sums


def in_bwd_stats(g, x, mean_rstd):
    """Reduction pass of the instance-norm backward -> fp64 [n, c, 2] = (sum g, sum g*xn)."""
    assert isinstance(g, F32B) and isinstance(x, F32B) and (g.n, g.c, g.h, g.w) == (x.n, x.c, x.h, x.w)
    sums = torch.empty((x.n, x.c, 2), dtype=torch.float64, device=x.t.device)
    with _rec("in_bwd_stats", (x.n, x.h, x.w, x.c)):
        check(lib().dge_in_bwd_stats(_p(g.t), _p(x.t), _f32(mean_rstd), _p(sums), x.n, x.c, x.h, x.w, _stream()))
    return sums


def in_bwd_apply(g, x, mean_rstd, style, dstyle, sums, mode, res=None, rscale=0.0, res_pool=False, noise=None,
                 slope=0.2, planes=2, gscale=None, out_kind="act"):
    """Apply pass of the instance-norm backward.  mode 0 -> F32B dx (+ rscale*res); mode 1 -> (dx*lrelu'(x) as Act, or as
    F32B with out_kind="f32b"; sums2 fp32 [2, c] = bias / noise-weight gradients).  gscale [n, c]: the incoming gradient is
    gscale * g (style modulation after the norm, StyleGAN1)."""
    assert isinstance(g, F32B) and isinstance(x, F32B) and (g.n, g.c, g.h, g.w) == (x.n, x.c, x.h, x.w)
    dev = x.t.device
    ds = None if dstyle is None else dstyle.contiguous()
    gs = None if gscale is None else gscale.contiguous()
    if mode == 0:
        out = F32B(x.n, x.c, x.h, x.w, dev)
        if res is not None:
            assert isinstance(res, F32B) and res.c == x.c and res.n == x.n
            assert (res.h, res.w) == ((x.h // 2, x.w // 2) if res_pool else (x.h, x.w)), "in_bwd_apply: residual size"
        with _rec("in_bwd_apply0", (x.n, x.h, x.w, x.c)):
            check(lib().dge_in_bwd_apply(_p(g.t), _p(x.t), _f32(mean_rstd), _f32(style), _f32(ds), _p(sums), _f32(gs), 0,
                                         _p(res.t) if res is not None else None, float(rscale), int(bool(res_pool)), None,
                                         float(slope), _p(out.t), None, None, x.n, x.c, x.h, x.w, planes, _stream()))
        return out
    out = Act(x.n, x.c, x.h, x.w, planes, dev) if out_kind == "act" else F32B(x.n, x.c, x.h, x.w, dev)
    sums2 = torch.empty((2, x.c), dtype=torch.float32, device=dev)
    with _rec("in_bwd_apply1", (x.n, x.h, x.w, x.c, planes)):
        check(lib().dge_in_bwd_apply(_p(g.t), _p(x.t), _f32(mean_rstd), _f32(style), _f32(ds), _p(sums), _f32(gs), 1, None,
                                     0.0, 0, _f32(noise), float(slope), _p(out.t) if out_kind != "act" else None,
                                     _p(out.t) if out_kind == "act" else None, _p(sums2), x.n, x.c, x.h, x.w, planes,
                                     _stream()))
    return out, sums2


def affine_relu_bwd(g, x, a, b, slope=0.0, up=1, skip=None, skip_up=1, out_act=True, out_f32b=False, planes=2):
    """Backward of t = act(a*x + b) [-> nearest x`up`] (biggan_generator.py:138-150, 178-190) -> (Act | None, F32B | None,
    sums fp32 [n, c, 2] = (d a, d b)).  `skip` (F32B, its first channels, `skip_up` x finer): added to dx (:192-203)."""
    assert isinstance(g, F32B) and isinstance(x, F32B)
    assert (g.n, g.c, g.h, g.w) == (x.n, x.c, x.h * up, x.w * up), "affine_relu_bwd: geometry mismatch"
    assert a.shape == (x.n, x.c) and b.shape == (x.n, x.c)
    if skip is not None:
        assert isinstance(skip, F32B) and (skip.n, skip.h, skip.w) == (x.n, x.h * skip_up, x.w * skip_up) and skip.c <= x.c
    dev = x.t.device
    act = Act(x.n, x.c, x.h, x.w, planes, dev) if out_act else None
    f = F32B(x.n, x.c, x.h, x.w, dev) if out_f32b else None
    sums = torch.empty((x.n, x.c, 2), dtype=torch.float32, device=dev)
    with _rec("affine_relu_bwd", (x.n, x.h, x.w, x.c, up)):
        check(lib().dge_affine_relu_bwd(_p(g.t), _p(x.t), _f32(a.contiguous()), _f32(b.contiguous()), float(slope), int(up),
                                        _p(skip.t) if skip is not None else None, skip.c if skip is not None else 0,
                                        int(skip_up), _p(f.t) if f is not None else None,
                                        _p(act.t) if act is not None else None, _p(sums), x.n, x.c, x.h, x.w, planes,
                                        _stream()))
    return act, f, sums


def from_rgb_bwd(d_f, f, img, slope=0.2, weight=None):
    """FromRGB backward -> fp32 [c, 4] = (dW[c, 0..2], db[c]); with `weight` ([c, cimg, 1, 1]) also the image gradient:
    -> (sums, d_img NCHW)."""
    assert isinstance(d_f, F32B) and isinstance(f, F32B)
    img = img.contiguous()
    sums = torch.empty((f.c, 4), dtype=torch.float32, device=f.t.device)
    d_img = torch.empty_like(img) if weight is not None else None
    wq = None if weight is None else weight.detach().contiguous()
    with _rec("from_rgb_bwd", (f.n, f.h, f.w, f.c)):
        check(lib().dge_from_rgb_bwd(_p(d_f.t), _p(f.t), _f32(img), _f32(wq), float(slope), _p(sums), _p(d_img), f.n,
                                     img.shape[1], f.c, f.h, f.w, _stream()))
    return sums if weight is None else (sums, d_img)


def sg2_prep_all(S, wp32, layers, outputs):
    """(styles, demods, rgb_styles, rgb_weights, handle) of one synthesis pass: one launch (SynthesisModule._prep);
    `handle` is what `sg2_prep_bwd` needs to run the transpose."""
    return S._prep(wp32, layers, outputs, with_handle=True)


def sg2_prep_bwd(S, handle, layers, outputs, sums, layer_offs, const_off, n):
    """Transpose of `sg2_prep_all` in one launch (dge_sg2_prep_bwd): the reductions of the synthesis backward -> d wp
    [n, num_layers, w_space_dim].  `sums`: one fp32 arena holding, at float offset layer_offs[i], layer i's
    `sg2_layer_bwd` sums [n, out_c, 5] and, at const_off, the style gradient of layer 0 [n, in_c]."""
    c = handle['cache']
    key = (tuple(layer_offs), const_off)
    g = c.get('gsrc')
    if g is None or g[0] != key:
        nl = len(layers)
        rows = []
        for i in range(nl):
            s_off, s_stride = (const_off, 1) if i == 0 else (layer_offs[i - 1], 5)
            rows.append((s_off, s_stride, layer_offs[i] + 4 if layers[i].demodulate else -1, -1))
        for k in range(len(outputs)):
            rows.append((-1, 0, -1, layer_offs[2 * k] + 1))
        g = c['gsrc'] = (key, torch.tensor(rows, dtype=torch.int64).to(sums.device))
    d_wp = torch.empty((n, S.num_layers, S.w_space_dim), dtype=torch.float32, device=sums.device)
    with _rec("sg2_prep_bwd", (n, c['n_items'])):
        check(lib().dge_sg2_prep_bwd(_p(c['items']), c['n_items'], _p(handle['arena']), _p(g[1]), _p(sums), _p(d_wp), n,
                                     S.num_layers, S.w_space_dim, _stream()))
    return d_wp


def sg2_layer_bwd(ya, ya_scale, dxs, dimg, rgbw, noise, noise_batched, noise_scalar, bias, demod, gain, slope,
                  out_kind="act", planes=2, sums=None):
    """Backward of a synthesis layer's epilogue + its consumers -> (d_conv as Act | F32B, sums fp32 [n, c, 5]);
    `sums`: a contiguous [n, c, 5] fp32 destination (a slice of the arena `sg2_prep_bwd` reads), else allocated."""
    assert isinstance(ya, Act) and (dxs is None or (isinstance(dxs, F32B) and (dxs.n, dxs.c, dxs.h, dxs.w) ==
                                                    (ya.n, ya.c, ya.h, ya.w)))
    dev = ya.t.device
    out = Act(ya.n, ya.c, ya.h, ya.w, planes, dev) if out_kind == "act" else F32B(ya.n, ya.c, ya.h, ya.w, dev)
    if sums is None:
        sums = torch.empty((ya.n, ya.c, 5), dtype=torch.float32, device=dev)
    assert sums.shape == (ya.n, ya.c, 5) and sums.is_contiguous() and sums.dtype == torch.float32
    di = None if dimg is None else dimg.contiguous()
    with _rec("sg2_layer_bwd", (ya.n, ya.h, ya.w, ya.c, out_kind)):
        check(lib().dge_sg2_layer_bwd(
            _p(ya.t), _f32(ya_scale), _p(dxs.t) if dxs is not None else None, _f32(di), _f32(rgbw), _f32(noise),
            (ya.h * ya.w if (noise is not None and noise_batched) else 0), float(noise_scalar), _f32(bias), _f32(demod),
            float(gain), float(slope), _p(out.t) if out_kind == "act" else None, planes,
            _p(out.t) if out_kind != "act" else None, _p(sums), ya.n, ya.c, ya.h, ya.w, ya.planes, _stream()))
    return out, sums


def up_fir_bwd_s2d(dconv, planes=2, box=False):
    """FIR transpose of the x2 layer (box=True: transpose of StyleGAN1's 2x2 box sum) as the space-to-depth operand of the
    stride-2 data-gradient conv: F32B [n, c, 2h, 2w] -> Act [n, 4c, h+1, w+1]."""
    assert isinstance(dconv, F32B) and dconv.h % 2 == 0 and dconv.w % 2 == 0
    h, w = dconv.h // 2, dconv.w // 2
    out = Act(dconv.n, 4 * dconv.c, h + 1, w + 1, planes, dconv.t.device)
    with _rec("up_fir_bwd_s2d", (dconv.n, h, w, dconv.c)):
        check(lib().dge_up_fir_bwd_s2d(_p(dconv.t), _p(out.t), dconv.n, dconv.c, h, w, planes, int(bool(box)), _stream()))
    return out


def rgb_up_bwd(d_out):
    """Transpose of the skip image's x2 up-sampling: [n, ch, 2h, 2w] -> [n, ch, h, w]."""
    d_out = d_out.contiguous()
    n, ch, ho, wo = d_out.shape
    out = torch.empty((n, ch, ho // 2, wo // 2), dtype=torch.float32, device=d_out.device)
    check(lib().dge_rgb_up_bwd(_f32(d_out), _p(out), n * ch, ho // 2, wo // 2, _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# LPIPS-VGG16 pieces (csrc/train_bwd.cu)
# ------------------------------------------------------------------------------------------------
def lpips_input(x, shift, scale, planes=2):
    """ScalingLayer + padding to 16 channels: NCHW [n, 3, h, w] -> Act [n, 16, h, w]."""
    x = x.contiguous()
    n, c, h, w = x.shape
    assert c == 3, "lpips_input: RGB images"
    out = Act(n, 16, h, w, planes, x.device)
    check(lib().dge_lpips_input(_f32(x), _p(out.t), *[float(v) for v in shift], *[float(v) for v in scale], n, h, w, planes,
                                _stream()))
    return out


def maxpool_to_act(x, planes=2):
    assert isinstance(x, F32B)
    out = Act(x.n, x.c, x.h // 2, x.w // 2, planes, x.t.device)
    with _rec("maxpool_to_act", (x.n, x.h, x.w, x.c)):
        check(lib().dge_maxpool_to_act(_p(x.t), _p(out.t), x.n, x.c, x.h, x.w, planes, _stream()))
    return out


def relu_pool_bwd(y, g_same=None, g_pool=None, planes=2, clamp_pos=False):
    """(y > 0) * (g_same + arg-max-routed g_pool) -> Act; y is the activated conv output (F32B or Act).
    clamp_pos: additionally max(., 0) (guided back-propagation)."""
    assert isinstance(y, (F32B, Act)) and (g_same is not None or g_pool is not None)
    out = Act(y.n, y.c, y.h, y.w, planes, y.t.device)
    is_act = isinstance(y, Act)
    with _rec("relu_pool_bwd", (y.n, y.h, y.w, y.c)):
        check(lib().dge_relu_pool_bwd(None if is_act else _p(y.t), _p(y.t) if is_act else None, y.planes if is_act else 0,
                                      _p(g_same.t) if g_same is not None else None,
                                      _p(g_pool.t) if g_pool is not None else None, _p(out.t), y.n, y.c, y.h, y.w, planes,
                                      int(bool(clamp_pos)), _stream()))
    return out


def lpips_dist(f, lin_w, out, eps=1e-10):
    """Adds one tap's distance into out[nb]; f = F32B features of the 2*nb images (first half vs second half)."""
    assert isinstance(f, F32B) and f.n % 2 == 0
    with _rec("lpips_dist", (f.n, f.h, f.w, f.c)):
        check(lib().dge_lpips_dist(_p(f.t), _f32(lin_w), _p(out), None, None, None, f.n // 2, f.c, f.h, f.w, float(eps),
                                   _stream()))


def lpips_dist_bwd(f, lin_w, go, want_a, want_b, eps=1e-10):
    """-> (ga, gb): F32B gradients of the tap distance w.r.t. the first / second half of f (None if not wanted)."""
    nb = f.n // 2
    ga = F32B(nb, f.c, f.h, f.w, f.t.device) if want_a else None
    gb = F32B(nb, f.c, f.h, f.w, f.t.device) if want_b else None
    with _rec("lpips_dist_bwd", (f.n, f.h, f.w, f.c)):
        check(lib().dge_lpips_dist(_p(f.t), _f32(lin_w), None, _f32(go.contiguous()), _p(ga.t) if ga is not None else None,
                                   _p(gb.t) if gb is not None else None, nb, f.c, f.h, f.w, float(eps), _stream()))
    return ga, gb


def launch_count():
    return int(_lib.load().dge_launch_count())


def launch_count_reset():
    _lib.load().dge_launch_count_reset()
