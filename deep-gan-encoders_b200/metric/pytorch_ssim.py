"""SSIM -- drop-in for the reference's `metric/pytorch_ssim.py` (`_ssim` :18-38, `SSIM` :40-61, `ssim` :63-74).

The five 11x11 depthwise gaussian convolutions, the SSIM map and its mean are one fused kernel
(`dge_ssim_sum`); only `size_average=True` (the only mode the inversion scripts use) is implemented.
Inputs that require grad (training) take the differentiable torch form `_ssim_mean_autograd` instead.
"""
import ctypes

import torch
import torch.nn.functional as F

from dge_b200 import ops


def _ssim_mean_autograd(a, b):
    """Differentiable mean SSIM (training path, `_ssim` :18-38): normalised 11-tap gaussian (sigma 1.5) applied
    separably per channel with zero padding 5, local moments, C1 = 0.01^2, C2 = 0.03^2, mean of the map."""
    from dge_b200 import autograd as tc
    c = a.shape[1]
    t = torch.arange(11, dtype=torch.float32, device=a.device) - 5.0
    g = torch.exp(-(t * t) / (2 * 1.5 ** 2))
    g = g / g.sum()
    win = (g[:, None] * g[None, :]).expand(c, 1, 11, 11).contiguous()

    def blur(v):
        return tc.lib_conv2d(v, win, padding=5, groups=c)

    mu_a, mu_b = blur(a), blur(b)
    var_a, var_b, cov = blur(a * a) - mu_a * mu_a, blur(b * b) - mu_b * mu_b, blur(a * b) - mu_a * mu_b
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    num = (2 * mu_a * mu_b + c1) * (2 * cov + c2)
    den = (mu_a * mu_a + mu_b * mu_b + c1) * (var_a + var_b + c2)
    return (num / den).mean()


class _SsimMeanFn(torch.autograd.Function):
    """mean(SSIM map) with the fused forward kernel and the two-pass fused backward (dge_ssim_grad); SSIM is symmetric in
    its arguments, so the gradient w.r.t. the first image is the same kernel with the roles swapped."""

    @staticmethod
    def forward(ctx, a, b):
        n, c, h, w = a.shape
        out = torch.empty(1, dtype=torch.float64, device=a.device)
        ops.check(ops.lib().dge_ssim_sum(ops._p(a), ops._p(b), n * c, h, w, ops._p(out), ops._stream()))
        ctx.save_for_backward(a, b)
        return (out / float(a.numel())).float().view(())

    @staticmethod
    def backward(ctx, go):
        a, b = ctx.saved_tensors
        n, c, h, w = a.shape
        go = go.contiguous().float().view(1)
        scratch = torch.empty((3,) + tuple(a.shape), dtype=torch.float32, device=a.device)
        grads = [None, None]
        for i, (x, y) in enumerate(((b, a), (a, b))):         # gradient w.r.t. y with x as the other image
            if ctx.needs_input_grad[i]:
                g = torch.empty_like(y)
                ops.check(ops.lib().dge_ssim_grad(ops._p(x), ops._p(y), ops._p(go), ops._p(scratch), ops._p(g), n * c, h, w,
                                                  ops._stream()))
                grads[i] = g
        return grads[0], grads[1]


FUSED_TRAIN = True     # False: `_ssim_mean_autograd` (separate torch nodes), the cross-check of the fused node


def _ssim_mean(img1, img2):
    if not (img1.is_cuda and img2.is_cuda):
        raise ops.DgeError('ssim: dge_b200 runs on a B200 only; there is no CPU fallback')
    if torch.is_grad_enabled() and (img1.requires_grad or img2.requires_grad):
        if FUSED_TRAIN:
            return _SsimMeanFn.apply(img1.float().contiguous(), img2.float().contiguous())
        return _ssim_mean_autograd(img1.float(), img2.float())
    assert img1.shape == img2.shape and img1.ndim == 4
    a, b = img1.float().contiguous(), img2.float().contiguous()
    n, c, h, w = a.shape
    out = torch.empty(1, dtype=torch.float64, device=a.device)
    ops.check(ops.lib().dge_ssim_sum(ops._p(a), ops._p(b), n * c, h, w, ops._p(out), ops._stream()))
    return (out / float(a.numel())).float().view(())


class SSIM(torch.nn.Module):
    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        if window_size != 11 or not size_average:
            raise NotImplementedError('dge_b200 SSIM: window_size=11, size_average=True (the reference defaults)')
        self.window_size = window_size
        self.size_average = size_average
        self.channel = 1

    def forward(self, img1, img2):
        self.channel = img1.size(1)
        return _ssim_mean(img1, img2)


def ssim(img1, img2, window_size=11, size_average=True):
    if window_size != 11 or not size_average:
        raise NotImplementedError('dge_b200 ssim: window_size=11, size_average=True (the reference defaults)')
    return _ssim_mean(img1, img2)
