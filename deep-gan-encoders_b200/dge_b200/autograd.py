"""Autograd bridge for the training step (E_align_s2.py:205-233 calls `loss.backward()` on the encoder).

`conv2d(x, w)` is `F.conv2d(x, w, padding=k//2)` for stride-1 1x1 / 3x3 convs with all three contractions on the
tcgen05 kernels:
    forward   dge_conv_forward                      y  = conv(x, W)
    dL/dx     dge_conv_forward on the transposed + flipped operand (dge_pack_conv_weight_dgrad)
    dL/dW     dge_conv_wgrad (pixels as the contraction index)
in split precision (bf16 hi+lo operands, fp32 accumulate).  It is the generic building block (LPIPS-VGG16, the
StyleGAN1 / BigGAN / E_Blur / E_BIG training graphs); the benchmarked StyleGAN2 pair has whole-block fused functions
(dge_b200/train_e.py, dge_b200/train_g.py).

`lib_conv2d` / `lib_conv_transpose2d` are the few convolutions still sent through cuDNN in a training graph (3-channel
from_rgb / ToRGB of the secondary families, depth-wise blur / SSIM windows).  PyTorch lets cuDNN round conv operands to
TF32 by default (~5e-4 per conv, 1e-2 on a chained image gradient -- outside the 1e-3 bar), so these run with TF32
switched off for exactly the duration of the call, forward and backward, and the process-wide flags are restored.
CUDA-only: there is no CPU fallback.
"""
import contextlib

import torch
import torch.nn.functional as F
from torch.autograd.function import once_differentiable

from . import ops


class _Conv2dTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, planes):
        cout, cin, k, k2 = w.shape
        assert k == k2 and k in (1, 3), "conv2d: 1x1 / 3x3 only"
        assert cin % 16 == 0 and cout % 16 == 0, "conv2d: channel counts must be multiples of 16"
        xa = ops.nchw_to_act(x.detach().float(), planes=planes)
        kind = ops.CONV_3X3 if k == 3 else ops.CONV_1X1
        y = ops.conv(xa, ops.pack_conv_weight(w, planes=planes), cout, kind, out_nchw=True)["nchw"]
        ctx.kind, ctx.planes, ctx.geom = kind, planes, (xa.n, xa.c, xa.h, xa.w)
        # the bf16 hi/lo image of x is read only by the weight gradient: frozen weights (the generator, the LPIPS VGG)
        # do not keep it alive
        if ctx.needs_input_grad[1]:
            ctx.save_for_backward(w, xa.t)
        else:
            ctx.save_for_backward(w)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        w = ctx.saved_tensors[0]
        cout, cin, k, _ = w.shape
        dya = ops.nchw_to_act(dy.contiguous().float(), planes=ctx.planes)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = ops.conv(dya, ops.pack_conv_weight_dgrad(w, planes=ctx.planes), cin, ctx.kind, out_nchw=True)["nchw"]
        if ctx.needs_input_grad[1]:
            n, c, h, wd = ctx.geom
            xa = ops.Act.wrap(ctx.saved_tensors[1], n, c, h, wd, ctx.planes)
            dw = ops.conv_wgrad(dya, xa, k)
        return dx, dw, None


def conv2d(x, w, planes=2):
    """Differentiable `F.conv2d(x, w, padding=w.shape[-1] // 2)` (stride 1, no bias) on the tcgen05 kernels."""
    if not (x.is_cuda and w.is_cuda):
        raise ops.DgeError("dge_b200.autograd.conv2d runs on a B200 only; no CPU fallback")
    return _Conv2dTC.apply(x, w, planes)


@contextlib.contextmanager
def fp32_library_convs():
    """TF32 off for cuDNN / cuBLAS inside the block; the previous process-wide flags come back on exit."""
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


class _LibConv(torch.autograd.Function):
    """aten::convolution / convolution_backward in true fp32 (see the module docstring)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding, groups, transposed):
        ctx.cfg = (stride, padding, groups, transposed, b is not None)
        ctx.save_for_backward(x, w)
        with fp32_library_convs():
            return torch.ops.aten.convolution(x, w, b, [stride, stride], [padding, padding], [1, 1], transposed,
                                              [0, 0], groups)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, padding, groups, transposed, has_b = ctx.cfg
        mask = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_b and ctx.needs_input_grad[2]]
        with fp32_library_convs():
            dx, dw, db = torch.ops.aten.convolution_backward(
                dy.contiguous(), x, w, [w.shape[1] * groups if transposed else w.shape[0]] if has_b else None,
                [stride, stride], [padding, padding], [1, 1], transposed, [0, 0], groups, mask)
        return dx, dw, db, None, None, None, None


def lib_conv2d(x, w, b=None, stride=1, padding=0, groups=1):
    return _LibConv.apply(x, w, b, stride, padding, groups, False)


def lib_conv_transpose2d(x, w, b=None, stride=1, padding=0, groups=1):
    return _LibConv.apply(x, w, b, stride, padding, groups, True)
