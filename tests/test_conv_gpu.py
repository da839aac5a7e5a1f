"""Kernel-level parity of the tcgen05 conv (dge_conv_forward) against a plain PyTorch fp32 reference
of the same op (F.conv2d / F.conv_transpose2d with TF32 disabled), through the C ABI.

Tolerances: split-precision bf16x3 (planes=2) 2e-4 of the output scale (observed ~1e-5);
plain bf16 (planes=1) 3e-2.  The CUDA-core checker kernel shares packing + epilogue with the
tcgen05 kernel and is run first so a failure localises to TMA/UMMA or to packing/epilogue.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    from dge_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return ops


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def _lrelu(v, slope, gain):
    return torch.where(v < 0, v * slope, v) * gain


PLAIN_CASES = [
    # n, cin, cout, h, w, planes
    (2, 32, 32, 16, 16, 2),
    (1, 16, 16, 16, 8, 2),
    (3, 64, 128, 20, 12, 2),
    (2, 128, 256, 9, 17, 2),
    (2, 64, 512, 8, 8, 2),
    (2, 256, 64, 4, 4, 2),
    (2, 32, 48, 33, 31, 2),
    (2, 64, 64, 16, 16, 1),
    (1, 64, 128, 16, 24, 2),     # CTA-pair mode with an ODD number of M tiles (the last pair has a dead second tile)
    (3, 128, 256, 16, 8, 2),     # pair mode, 3 M tiles x 1 N tile
    (1, 32, 512, 16, 24, 2),     # pair mode, odd M tiles x 2 N tiles
    (2, 16, 32, 64, 48, 2),      # 4-block tiles (32x16) with a ragged right edge
    (1, 16, 16, 40, 24, 2),      # 2x2-block tiles, ragged bottom edge
    (8, 512, 512, 4, 4, 2),      # split-K (small map, long K chain): partial sums + finish kernel
    (8, 512, 512, 8, 8, 2),
    (4, 256, 512, 16, 16, 2),
    (1, 512, 256, 5, 7, 2),
]


@pytest.mark.parametrize("checker", [True, False], ids=["checker", "tcgen05"])
@pytest.mark.parametrize("case", PLAIN_CASES, ids=lambda c: "n%d_ci%d_co%d_%dx%d_p%d" % c)
def test_conv3x3_full_epilogue(case, checker):
    ops = _setup()
    n, cin, cout, h, w, planes = case
    g = torch.Generator(device="cuda").manual_seed(1234 + cin + cout)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3.0 * cin ** 0.5)
    s_in = torch.rand(n, cin, device="cuda", generator=g) + 0.5
    s_out = torch.rand(n, cout, device="cuda", generator=g) + 0.5
    dm = torch.rand(n, cout, device="cuda", generator=g) + 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    noise = torch.randn(h, w, device="cuda", generator=g)
    strength = 0.37

    xa = ops.nchw_to_act(x, scale=s_in, planes=planes)
    wpk = ops.pack_conv_weight(wt, scale=1.0, planes=planes)
    res = ops.conv(xa, wpk, cout, ops.CONV_3X3, demod=dm, noise=noise, noise_scalar=strength, bias=bias, slope=0.2,
                   gain=2 ** 0.5, out_act=True, out_scale=s_out, out_f32b=True, out_nchw=True, checker=checker)
    torch.cuda.synchronize()

    ref = F.conv2d(x * s_in[:, :, None, None], wt, padding=1) * dm[:, :, None, None]
    ref = ref + noise[None, None] * strength + bias[None, :, None, None]
    ref = _lrelu(ref, 0.2, 2 ** 0.5)
    tol = 2e-4 if planes == 2 else 3e-2
    assert _rel(res["nchw"], ref) < tol
    assert _rel(res["f32b"].to_nchw(), ref) < tol
    assert _rel(res["act"].to_nchw(), ref * s_out[:, :, None, None]) < (tol if planes == 2 else 4e-2)


@pytest.mark.parametrize("checker", [True, False], ids=["checker", "tcgen05"])
def test_conv3x3_encoder_epilogue_and_rgb(checker):
    ops = _setup()
    n, cin, cout, h, w = 2, 32, 64, 24, 16
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3.0 * cin ** 0.5)
    nw = torch.randn(cout, device="cuda", generator=g)
    bias = torch.randn(cout, device="cuda", generator=g)
    noise = torch.randn(n, h, w, device="cuda", generator=g)
    rgbw = torch.randn(n, 3, cout, device="cuda", generator=g)
    rgb0 = torch.randn(n, 3, h, w, device="cuda", generator=g)
    rgb = rgb0.clone()
    xa = ops.nchw_to_act(x)
    wpk = ops.pack_conv_weight(wt)
    res = ops.conv(xa, wpk, cout, ops.CONV_3X3, noise=noise, noise_batched=True, noise_w=nw, bias=bias, slope=0.2,
                   out_f32b=True, rgb_w=rgbw, rgb_out=rgb, checker=checker)
    torch.cuda.synchronize()
    ref = F.conv2d(x, wt, padding=1) + noise[:, None] * nw[None, :, None, None] + bias[None, :, None, None]
    ref = _lrelu(ref, 0.2, 1.0)
    assert _rel(res["f32b"].to_nchw(), ref) < 2e-4
    ref_rgb = rgb0 + torch.einsum("nchw,nkc->nkhw", ref, rgbw)
    assert _rel(rgb, ref_rgb) < 2e-4


@pytest.mark.parametrize("checker", [True, False], ids=["checker", "tcgen05"])
@pytest.mark.parametrize("pool", [False, True])
def test_conv1x1_blend(pool, checker):
    ops = _setup()
    n, cin, cout, h, w = 2, 32, 64, 12, 20
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 1, 1, device="cuda", generator=g) / cin ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    src = torch.randn(n, cout, 2 * h if pool else h, 2 * w if pool else w, device="cuda", generator=g)
    res = ops.conv(ops.nchw_to_act(x), ops.pack_conv_weight(wt), cout, ops.CONV_1X1, bias=bias,
                   blend_src=ops.nchw_to_f32b(src), blend_pool=pool, blend_a=0.111, blend_b=0.889, out_f32b=True,
                   checker=checker)
    torch.cuda.synchronize()
    s = F.avg_pool2d(src, 2, 2) if pool else src
    ref = 0.111 * s + 0.889 * (F.conv2d(x, wt) + bias[None, :, None, None])
    assert _rel(res["f32b"].to_nchw(), ref) < 2e-4


@pytest.mark.parametrize("case", [(2, 16, 32, 32, 16), (2, 32, 64, 64, 32), (1, 64, 128, 16, 24), (8, 512, 512, 8, 8),
                                  (8, 512, 512, 16, 16)],
                         ids=lambda c: "n%d_ci%d_co%d_%dx%d" % c)
def test_conv3x3_pooled_output(case):
    """out_pool: the epilogue value averaged over 2x2 (avg_pool2d(2,2) fused into the producer, E.py:76-77)."""
    ops = _setup()
    n, cin, cout, h, w = case
    g = torch.Generator(device="cuda").manual_seed(5 + cout)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3.0 * cin ** 0.5)
    nw = torch.randn(cout, device="cuda", generator=g)
    bias = torch.randn(cout, device="cuda", generator=g)
    noise = torch.randn(n, h, w, device="cuda", generator=g)
    res = ops.conv(ops.nchw_to_act(x), ops.pack_conv_weight(wt), cout, ops.CONV_3X3, noise=noise, noise_batched=True,
                   noise_w=nw, bias=bias, slope=0.2, out_f32b_pool=True)
    torch.cuda.synchronize()
    ref = F.conv2d(x, wt, padding=1) + noise[:, None] * nw[None, :, None, None] + bias[None, :, None, None]
    ref = F.avg_pool2d(_lrelu(ref, 0.2, 1.0), 2, 2)
    assert _rel(res["f32b_pool"].to_nchw(), ref) < 2e-4


UP_CASES = [(2, 32, 32, 8, 8), (2, 64, 32, 16, 16), (1, 64, 128, 9, 5), (2, 128, 256, 8, 8), (2, 32, 512, 4, 4),
            (1, 128, 64, 16, 24), (3, 256, 128, 8, 8),   # (pair mode with odd tile counts)
            (8, 512, 512, 4, 4), (8, 512, 512, 8, 8)]    # (split-K: parts added in place into the raw map)


@pytest.mark.parametrize("checker", [True, False], ids=["checker", "tcgen05"])
@pytest.mark.parametrize("case", UP_CASES, ids=lambda c: "n%d_ci%d_co%d_%dx%d" % c)
def test_up_conv_raw_and_fir(case, checker):
    ops = _setup()
    n, cin, cout, h, w = case
    g = torch.Generator(device="cuda").manual_seed(99 + cout)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3.0 * cin ** 0.5)
    dm = torch.rand(n, cout, device="cuda", generator=g) + 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    noise = torch.randn(2 * h, 2 * w, device="cuda", generator=g)
    s_out = torch.rand(n, cout, device="cuda", generator=g) + 0.5
    res = ops.conv(ops.nchw_to_act(x), ops.pack_conv_weight(wt, flip=True), cout, ops.CONV_UP3X3, checker=checker)
    torch.cuda.synchronize()
    wf = wt.flip(2, 3).permute(1, 0, 2, 3).contiguous()  # [in, out, k, k]
    raw_ref = F.conv_transpose2d(x, wf, stride=2)
    raw = ops.F32B.__new__(ops.F32B)
    raw.n, raw.c, raw.h, raw.w, raw.t = n, cout, 2 * h + 1, 2 * w + 1, res["raw_up"]
    assert _rel(raw.to_nchw(), raw_ref) < 2e-4

    fir = torch.tensor([1., 3., 3., 1.], device="cuda")
    k2 = torch.outer(fir, fir)
    k2 = (k2 / k2.sum() * 4.0)[None, None]
    y = F.conv2d(F.pad(raw_ref, (1, 1, 1, 1)).reshape(n * cout, 1, 2 * h + 3, 2 * w + 3), k2).reshape(n, cout, 2 * h, 2 * w)
    y = y * dm[:, :, None, None] + noise[None, None] * 0.5 + bias[None, :, None, None]
    y = _lrelu(y, 0.2, 2 ** 0.5)
    out = ops.up_fir_epilogue(res["raw_up"], n, cout, 2 * h, 2 * w, demod=dm, noise=noise, noise_scalar=0.5, bias=bias,
                              slope=0.2, gain=2 ** 0.5, out_scale=s_out, out_act=True, out_nchw=True)
    torch.cuda.synchronize()
    assert _rel(out["nchw"], y) < 2e-4
    assert _rel(out["act"].to_nchw(), y * s_out[:, :, None, None]) < 2e-4


@pytest.mark.parametrize("case", [(2, 32, 64, 24, 16, 3), (2, 128, 64, 16, 16, 3), (1, 64, 128, 9, 17, 1)],
                         ids=lambda c: "n%d_ci%d_co%d_%dx%d_k%d" % c)
def test_conv_data_gradient(case):
    """dL/dx of y = conv2d(x, W) through the same tcgen05 kernel with the transposed + flipped operand
    (dge_pack_conv_weight_dgrad), against torch.autograd."""
    ops = _setup()
    n, cin, cout, h, w, k = case
    g = torch.Generator(device="cuda").manual_seed(31 + cout)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g, requires_grad=True)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (k * cin ** 0.5)
    dy = torch.randn(n, cout, h, w, device="cuda", generator=g)
    y = F.conv2d(x, wt, padding=k // 2)
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    wd = ops.pack_conv_weight_dgrad(wt)
    kind = ops.CONV_3X3 if k == 3 else ops.CONV_1X1
    dx = ops.conv(ops.nchw_to_act(dy), wd, cin, kind, out_f32b=True)["f32b"].to_nchw()
    torch.cuda.synchronize()
    assert _rel(dx, dx_ref) < 2e-4


@pytest.mark.parametrize("case", [(2, 16, 32, 24, 16, 3), (2, 128, 128, 16, 16, 3), (1, 64, 128, 9, 17, 1),
                                  (2, 512, 512, 4, 4, 3), (1, 144, 192, 20, 40, 3), (2, 32, 32, 64, 80, 3),
                                  (3, 256, 64, 33, 31, 1), (2, 64, 64, 7, 8, 3), (3, 32, 48, 11, 5, 3),
                                  (8, 512, 512, 8, 8, 3),
                                  # thin layers: the cross-stacked mode (rows = (kx, o), columns = (ky, i, plane)) with
                                  # R = 16 / 32 x 2 blocks / 24 / 40 output channels per CTA, ragged and <= 8-wide maps
                                  (2, 16, 16, 40, 40, 3), (1, 32, 64, 33, 47, 3), (2, 16, 48, 21, 70, 3),
                                  (1, 32, 80, 64, 64, 3), (4, 16, 32, 128, 128, 3), (2, 32, 32, 8, 8, 3),
                                  (2, 16, 16, 4, 6, 3), (1, 32, 32, 256, 256, 3)],
                         ids=lambda c: "n%d_ci%d_co%d_%dx%d_k%d" % c)
def test_conv_weight_gradient(case):
    """dL/dW of y = conv2d(x, W) (dge_conv_wgrad: pixels are the contraction index, both operands MN-major tiles of the
    ACT layout), against torch.autograd in fp32 (TF32 off).  Ragged maps, partial channel blocks, tiny maps, 1x1."""
    ops = _setup()
    n, cin, cout, h, w, k = case
    g = torch.Generator(device="cuda").manual_seed(57 + cout + h)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (k * cin ** 0.5)).requires_grad_(True)
    dy = torch.randn(n, cout, h, w, device="cuda", generator=g)
    y = F.conv2d(x, wt, padding=k // 2)
    (dw_ref,) = torch.autograd.grad(y, wt, dy)
    dw = ops.conv_wgrad(ops.nchw_to_act(dy), ops.nchw_to_act(x), k)
    torch.cuda.synchronize()
    assert dw.shape == dw_ref.shape
    assert _rel(dw, dw_ref) < 2e-4
    # accumulate=True adds a second gradient into the same buffer
    ops.conv_wgrad(ops.nchw_to_act(dy), ops.nchw_to_act(x), k, out=dw, accumulate=True)
    torch.cuda.synchronize()
    assert _rel(dw, 2 * dw_ref) < 2e-4


def test_conv_weight_gradient_plain_bf16_is_close():
    """planes=1 (plain bf16 operands) is the fast mode: close, but not the parity mode."""
    ops = _setup()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2, 64, 32, 32, device="cuda", generator=g)
    dy = torch.randn(2, 64, 32, 32, device="cuda", generator=g)
    wt = torch.zeros(64, 64, 3, 3, device="cuda", requires_grad=True)
    (dw_ref,) = torch.autograd.grad(F.conv2d(x, wt, padding=1), wt, dy)
    dw = ops.conv_wgrad(ops.nchw_to_act(dy, planes=1), ops.nchw_to_act(x, planes=1), 3)
    torch.cuda.synchronize()
    assert _rel(dw, dw_ref) < 2e-2
    # thin layer (cross-stacked mode, one MMA per K step)
    x = torch.randn(2, 32, 40, 56, device="cuda", generator=g)
    dy = torch.randn(2, 64, 40, 56, device="cuda", generator=g)
    wt = torch.zeros(64, 32, 3, 3, device="cuda", requires_grad=True)
    (dw_ref,) = torch.autograd.grad(F.conv2d(x, wt, padding=1), wt, dy)
    dw = ops.conv_wgrad(ops.nchw_to_act(dy, planes=1), ops.nchw_to_act(x, planes=1), 3)
    torch.cuda.synchronize()
    assert _rel(dw, dw_ref) < 2e-2
