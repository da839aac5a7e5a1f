"""Autograd bridge for the training step (E_align_s2.py:205-233 calls `loss.backward()` on the encoder).

`conv2d(x, w)` is `F.conv2d(x, w, padding=k//2)` for the stride-1 1x1 / 3x3 convs of the encoder with all three
contractions on the tcgen05 kernels:
    forward   dge_conv_forward                      y  = conv(x, W)
    dL/dx     dge_conv_forward on the transposed + flipped operand (dge_pack_conv_weight_dgrad)
    dL/dW     dge_conv_wgrad (pixels as the contraction index)
in split precision (bf16 hi+lo operands, fp32 accumulate).  The point-wise / reduction steps around the convs
(instance norm, noise, bias, leaky-relu, pools, the style GEMVs) stay torch CUDA ops in the training path of this
build -- see model/E/E.py `BE._forward_autograd`; the fused forward-only kernels remain the inference path.
CUDA-only: there is no CPU fallback.
"""
import torch

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
        ctx.xa, ctx.kind, ctx.planes = xa, kind, planes      # the bf16 hi/lo image of x is what the backward reads
        ctx.save_for_backward(w)
        return y

    @staticmethod
    def backward(ctx, dy):
        (w,) = ctx.saved_tensors
        cout, cin, k, _ = w.shape
        dya = ops.nchw_to_act(dy.contiguous().float(), planes=ctx.planes)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = ops.conv(dya, ops.pack_conv_weight_dgrad(w, planes=ctx.planes), cin, ctx.kind, out_nchw=True)["nchw"]
        if ctx.needs_input_grad[1]:
            dw = ops.conv_wgrad(dya, ctx.xa, k)
        # ctx.xa stays: E_align_s2.py:205 calls backward(retain_graph=True) and walks the encoder's graph a second time
        return dx, dw, None


def require_fp32_library_convs():
    """The training path still sends a few convolutions through cuDNN (from_rgb / ToRGB with 3 channels, the x2
    transposed convs, the FIR and SSIM windows).  PyTorch lets cuDNN run those in TF32 by default: operands rounded
    to 10 mantissa bits, ~5e-4 relative error per conv, outside the 1e-3 parity bar once chained (measured: 1e-2 on
    the image gradient).  The reference's numbers are fp32, so the path pins the library convs to fp32."""
    if torch.backends.cudnn.allow_tf32:
        torch.backends.cudnn.allow_tf32 = False
    if torch.backends.cuda.matmul.allow_tf32:
        torch.backends.cuda.matmul.allow_tf32 = False


def conv2d(x, w, planes=2):
    """Differentiable `F.conv2d(x, w, padding=w.shape[-1] // 2)` (stride 1, no bias) on the tcgen05 kernels."""
    if not (x.is_cuda and w.is_cuda):
        raise ops.DgeError("dge_b200.autograd.conv2d runs on a B200 only; no CPU fallback")
    return _Conv2dTC.apply(x, w, planes)
