"""Equalised-learning-rate primitives -- drop-in for the reference's `model/utils/lreq.py`
(:39-173: Linear, Conv2d, ConvTranspose2d).

Same constructor arguments, parameter names/shapes, init distributions and the
`lr_equalization_coef` attribute `LREQAdam` reads off the parameters (lreq.py:60-62, 118-120).
`implicit_lreq` is globally True in the reference (:23-24): the forward uses the raw weights.
Forward = dge_b200 kernels (dense GEMV / tcgen05 conv); forward-only, CUDA-only.
"""
import numpy as np
import torch
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

from dge_b200 import ops


class Bool:
    def __init__(self):
        self.value = False

    def __bool__(self):
        return self.value

    __nonzero__ = __bool__

    def set(self, value):
        self.value = value


use_implicit_lreq = Bool()
use_implicit_lreq.set(True)


def make_tuple(x, n):
    if isinstance(x, (tuple, list)):
        return tuple(x)
    return tuple(x for _ in range(n))


def _guard(module_name, *tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ops.DgeError(f'{module_name}: dge_b200 runs on a B200 only (got a {t.device} tensor); no CPU fallback')
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(f'{module_name}: dge_b200 kernels are forward-only in this build; use torch.no_grad()')


class Linear(nn.Module):
    """Reference lreq.py:39-75."""

    def __init__(self, in_features, out_features, bias=True, gain=np.sqrt(2.0), lrmul=1.0,
                 implicit_lreq=use_implicit_lreq):
        super().__init__()
        self.in_features = in_features
        self.weight = Parameter(torch.Tensor(out_features, in_features))
        if bias:
            self.bias = Parameter(torch.Tensor(out_features))
        else:
            self.register_parameter('bias', None)
        self.std = 0
        self.gain = gain
        self.lrmul = lrmul
        self.implicit_lreq = implicit_lreq
        self.reset_parameters()

    def reset_parameters(self):
        self.std = self.gain / np.sqrt(self.in_features) * self.lrmul
        if not self.implicit_lreq:
            init.normal_(self.weight, mean=0, std=1.0 / self.lrmul)
        else:
            init.normal_(self.weight, mean=0, std=self.std / self.lrmul)
            setattr(self.weight, 'lr_equalization_coef', self.std)
            if self.bias is not None:
                setattr(self.bias, 'lr_equalization_coef', self.lrmul)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, input):
        _guard('ln.Linear', input, self.weight)
        if not self.implicit_lreq:
            return ops.dense(input.float(), self.weight, self.bias, wscale=self.std, bscale=self.lrmul)
        return ops.dense(input.float(), self.weight, self.bias)


class Conv2d(nn.Module):
    """Reference lreq.py:78-156."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, dilation=1,
                 groups=1, bias=True, gain=np.sqrt(2.0), transpose=False, transform_kernel=False, lrmul=1.0,
                 implicit_lreq=use_implicit_lreq):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = make_tuple(kernel_size, 2)
        self.stride = make_tuple(stride, 2)
        self.padding = make_tuple(padding, 2)
        self.output_padding = make_tuple(output_padding, 2)
        self.dilation = make_tuple(dilation, 2)
        self.groups = groups
        self.gain = gain
        self.lrmul = lrmul
        self.transpose = transpose
        self.fan_in = np.prod(self.kernel_size) * in_channels // groups
        self.transform_kernel = transform_kernel
        if transpose:
            self.weight = Parameter(torch.Tensor(in_channels, out_channels // groups, *self.kernel_size))
        else:
            self.weight = Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.std = 0
        self.implicit_lreq = implicit_lreq
        self.reset_parameters()

    def reset_parameters(self):
        self.std = self.gain / np.sqrt(self.fan_in)
        if not self.implicit_lreq:
            init.normal_(self.weight, mean=0, std=1.0 / self.lrmul)
        else:
            init.normal_(self.weight, mean=0, std=self.std / self.lrmul)
            setattr(self.weight, 'lr_equalization_coef', self.std)
            if self.bias is not None:
                setattr(self.bias, 'lr_equalization_coef', self.lrmul)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    # -- helpers used by the fused encoder blocks ---------------------------------------------------
    def is_plain(self, k):
        """True for the stride-1 'same' convs the tcgen05 kernel implements directly."""
        return (not self.transpose and not self.transform_kernel and self.kernel_size == (k, k)
                and self.stride == (1, 1) and self.padding == (k // 2, k // 2) and self.dilation == (1, 1)
                and self.groups == 1 and self.in_channels % 16 == 0 and self.out_channels % 16 == 0)

    def packed(self, planes=2):
        """WPK weights, cached per parameter version (the weights change every optimiser step)."""
        key = (ops.weight_key(self.weight), planes)
        if getattr(self, '_wpk_key', None) != key:
            scale = 1.0 if self.implicit_lreq else self.std
            self._wpk = ops.pack_conv_weight(self.weight, scale=scale, planes=planes)
            self._wpk_key = key
        return self._wpk

    def packed_down4(self, planes=2):
        """`transform_kernel` strided conv (lreq.py:144-147): W4 = 0.25 * (sum of the four 1-pixel shifts of the zero-padded
        3x3 kernel), used with stride 2 / padding 1 -> 16-tap WPK for DGE_CONV_DOWN4X4S2."""
        assert self.transform_kernel and not self.transpose and self.stride == (2, 2) and self.kernel_size == (3, 3)
        key = (ops.weight_key(self.weight), planes, 'down4')
        if getattr(self, '_wpk4_key', None) != key:
            w = torch.nn.functional.pad(self.weight.detach(), (1, 1, 1, 1))
            w4 = (w[:, :, 1:, 1:] + w[:, :, :-1, 1:] + w[:, :, 1:, :-1] + w[:, :, :-1, :-1]) * 0.25
            scale = 1.0 if self.implicit_lreq else self.std
            self._wpk4 = ops.pack_conv_weight(w4.contiguous(), scale=scale, planes=planes)
            self._wpk4_key = key
        return self._wpk4

    def scaled_bias(self):
        if self.bias is None:
            return None
        return self.bias.detach() if self.implicit_lreq else (self.bias.detach() * self.lrmul).contiguous()

    def forward(self, x):
        _guard('ln.Conv2d', x, self.weight)
        k = self.kernel_size[0]
        if not (self.is_plain(1) or self.is_plain(3)):
            raise NotImplementedError('stand-alone ln.Conv2d: only 1x1 / 3x3 stride-1 same convs with channel counts '
                                      'that are multiples of 16 are implemented (the case-1 encoder path)')
        xa = ops.nchw_to_act(x.float())
        kind = ops.CONV_3X3 if k == 3 else ops.CONV_1X1
        return ops.conv(xa, self.packed(), self.out_channels, kind, bias=self.scaled_bias(), out_nchw=True)['nchw']


class ConvTranspose2d(Conv2d):
    """Reference lreq.py:159-173 (parameters/init only; used by the StyleGAN1 generator)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, dilation=1,
                 groups=1, bias=True, gain=np.sqrt(2.0), transform_kernel=False, lrmul=1.0,
                 implicit_lreq=use_implicit_lreq):
        super().__init__(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=stride,
                         padding=padding, output_padding=output_padding, dilation=dilation, groups=groups,
                         bias=bias, gain=gain, transpose=True, transform_kernel=transform_kernel, lrmul=lrmul,
                         implicit_lreq=implicit_lreq)
