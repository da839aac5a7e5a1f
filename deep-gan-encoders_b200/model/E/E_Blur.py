"""B200-native case-2 encoder -- drop-in for the reference's `model/E/E_Blur.py` (BEBlock :14-85, BE :88-134), the
encoder `embedding_img.py` uses.  Differences from `model/E/E.py`: a depthwise 3x3 `Blur` before `conv_2`, and for
`resolution >= 128` (counter hard-wired to start at 1024, :99,105 -- so the first FOUR blocks regardless of the image
size) `conv_2` is the stride-2 `transform_kernel` conv (4x4 effective) instead of conv + avg-pool.

Kernels: the instance-norm apply, the blur and (for the strided blocks) a space-to-depth re-layout are ONE pass
(`dge_instance_norm_blur`); the strided conv runs on the tensor cores as a 16-tap conv over the 4 input phases
(`DGE_CONV_DOWN4X4S2`), exact at the borders (the blurred intermediate is zero-padded).

Training: as in `model/E/E.py`, a call that must be recorded for backward runs ONE fused autograd node per block
(`dge_b200/train_e.py`, `FUSED_TRAIN`): the blur is its own transpose, and both gradients of the stride-2 `transform_kernel`
conv are a stride-1 3x3 conv over the space-to-depth operand the forward already wrote.  `_forward_autograd` -- the graph of
separate torch nodes (cuDNN for the blur and the strided conv) -- is kept as the cross-check (`FUSED_TRAIN = False`).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import model.utils.lreq as ln
from model.E.E import _mean_std, _wants_grad
from model.utils.net import FromRGB
from model.stylegan1.net import Blur
from dge_b200 import autograd as tc
from dge_b200 import ops

DEFAULT_PLANES = 2
FUSED_TRAIN = True      # one fused autograd node per block (dge_b200/train_e.py); False: separate torch nodes


class BEBlock(nn.Module):
    def __init__(self, inputs, outputs, latent_size, has_last_conv=True, fused_scale=True):
        super().__init__()
        self.has_last_conv = has_last_conv
        self.noise_weight_1 = nn.Parameter(torch.zeros(1, inputs, 1, 1))
        self.bias_1 = nn.Parameter(torch.zeros(1, inputs, 1, 1))
        self.instance_norm_1 = nn.InstanceNorm2d(inputs, affine=False, eps=1e-8)
        self.inver_mod1 = ln.Linear(2 * inputs, latent_size, gain=1)
        self.conv_1 = ln.Conv2d(inputs, inputs, 3, 1, 1, bias=False)
        self.noise_weight_2 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.bias_2 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.instance_norm_2 = nn.InstanceNorm2d(inputs, affine=False, eps=1e-8)
        self.inver_mod2 = ln.Linear(2 * inputs, latent_size, gain=1)
        self.blur = Blur(inputs)
        if has_last_conv:
            if fused_scale:
                self.conv_2 = ln.Conv2d(inputs, outputs, 3, 2, 1, bias=False, transform_kernel=True)
            else:
                self.conv_2 = ln.Conv2d(inputs, outputs, 3, 1, 1, bias=False)
        self.fused_scale = fused_scale
        self.inputs = inputs
        self.outputs = outputs
        if self.inputs != self.outputs:
            self.conv_3 = ln.Conv2d(inputs, outputs, 1, 1, 0)
        self.planes = DEFAULT_PLANES
        self.noise_mode = 'reference'

    def _noise(self, n, h, w, device):
        if self.noise_mode == 'device':
            return torch.randn([n, 1, h, w], device=device)
        return torch.randn([n, 1, h, w]).to(device)

    def run(self, x):
        n, c, h, w = x.n, x.c, x.h, x.w
        dev = x.t.device
        style1, mr1 = ops.instance_stats(x, self.instance_norm_1.eps)
        w1 = ops.dense(style1, self.inver_mod1.weight, self.inver_mod1.bias)
        xn, _ = ops.instance_norm(x, mr1, planes=self.planes)
        y1 = ops.conv(xn, self.conv_1.packed(self.planes), c, ops.CONV_3X3, noise=self._noise(n, h, w, dev),
                      noise_batched=True, noise_w=self.noise_weight_1.detach().view(-1),
                      bias=self.bias_1.detach().view(-1), slope=0.2, out_f32b=True)['f32b']
        style2, mr2 = ops.instance_stats(y1, self.instance_norm_2.eps)
        w2 = ops.dense(style2, self.inver_mod2.weight, self.inver_mod2.bias)
        nw2, b2 = self.noise_weight_2.detach().view(-1), self.bias_2.detach().view(-1)
        if self.has_last_conv:
            if self.fused_scale:
                xb = ops.instance_norm_blur(y1, mr2, s2d=True, planes=self.planes)                    # :69-71
                y2 = ops.conv(xb, self.conv_2.packed_down4(self.planes), self.outputs, ops.CONV_DOWN4X4S2,
                              noise=self._noise(n, h // 2, w // 2, dev), noise_batched=True, noise_w=nw2, bias=b2,
                              slope=0.2, out_f32b=True)['f32b']                                       # :72-75
                y2_pool = False
            else:
                xb = ops.instance_norm_blur(y1, mr2, s2d=False, planes=self.planes)
                y2 = ops.conv(xb, self.conv_2.packed(self.planes), self.outputs, ops.CONV_3X3,
                              noise=self._noise(n, h, w, dev), noise_batched=True, noise_w=nw2, bias=b2, slope=0.2,
                              out_f32b=True)['f32b']
                y2_pool = True                                                                        # :76-77
            if self.inputs != self.outputs:
                rp = ops.avgpool_to_act(x, planes=self.planes)                                        # :78
                out = ops.conv(rp, self.conv_3.packed(self.planes), self.outputs, ops.CONV_1X1,
                               bias=self.conv_3.scaled_bias(), blend_src=y2, blend_pool=y2_pool, blend_a=0.111,
                               blend_b=0.889, out_f32b=True)['f32b']                                  # :80-84
            else:
                out = ops.blend(y2, x, 0.111, 0.889, pool=3 if y2_pool else 2)
        else:
            _, y1n = ops.instance_norm(y1, mr2, out_act=False, out_f32b=True)
            if self.inputs != self.outputs:
                out = ops.conv(ops.f32b_to_act(x, self.planes), self.conv_3.packed(self.planes), self.outputs,
                               ops.CONV_1X1, bias=self.conv_3.scaled_bias(), blend_src=y1n, blend_pool=False,
                               blend_a=0.111, blend_b=0.889, out_f32b=True)['f32b']
            else:
                out = ops.blend(y1n, x, 0.111, 0.889, pool=False)
        return out, w1, w2

    def _forward_autograd(self, x):
        """Differentiable form of the block (E_Blur.py:50-85) on NCHW tensors."""
        n, c, h, w = x.shape
        dev = x.device
        w1 = F.linear(_mean_std(x), self.inver_mod1.weight, self.inver_mod1.bias)
        res = x
        y = tc.conv2d(F.instance_norm(x, eps=self.instance_norm_1.eps), self.conv_1.weight, self.planes)
        y = F.leaky_relu(torch.addcmul(y, self.noise_weight_1, self._noise(n, h, w, dev)) + self.bias_1, 0.2)
        w2 = F.linear(_mean_std(y), self.inver_mod2.weight, self.inver_mod2.bias)
        y = F.instance_norm(y, eps=self.instance_norm_2.eps)
        if self.has_last_conv:
            y = tc.lib_conv2d(y, self.blur.weight, groups=self.blur.groups, padding=1)                  # :71
            if self.fused_scale:                                                                   # :72, lreq.py:144-156
                k = F.pad(self.conv_2.weight, (1, 1, 1, 1))
                k = (k[:, :, 1:, 1:] + k[:, :, :-1, 1:] + k[:, :, 1:, :-1] + k[:, :, :-1, :-1]) * 0.25
                y = tc.lib_conv2d(y, k, stride=2, padding=1)
            else:
                y = tc.conv2d(y, self.conv_2.weight, self.planes)
            nh, nw = y.shape[2], y.shape[3]
            y = F.leaky_relu(torch.addcmul(y, self.noise_weight_2, self._noise(n, nh, nw, dev)) + self.bias_2, 0.2)
            if not self.fused_scale:
                y = F.avg_pool2d(y, 2, 2)
            res = F.avg_pool2d(res, 2, 2)
        if self.inputs != self.outputs:
            res = tc.conv2d(res, self.conv_3.weight, self.planes) + self.conv_3.bias.view(1, -1, 1, 1)
        return 0.111 * y + 0.889 * res, w1, w2

    def forward(self, x):
        if _wants_grad(self, x):
            return self._forward_autograd(x.float())
        ln._guard('E_Blur.BEBlock', x, self.conv_1.weight)
        out, w1, w2 = self.run(ops.nchw_to_f32b(x.float()))
        return out.to_nchw(), w1, w2


class BE(nn.Module):
    def __init__(self, startf=16, maxf=512, layer_count=9, latent_size=512, channels=3):
        super().__init__()
        self.maxf = maxf
        self.startf = startf
        self.latent_size = latent_size
        self.layer_to_resolution = [0 for _ in range(layer_count)]
        self.decode_block = nn.ModuleList()
        self.layer_count = layer_count
        inputs = startf
        outputs = startf * 2
        resolution = 1024
        self.FromRGB = FromRGB(channels, inputs)
        for i in range(layer_count):
            has_last_conv = i + 1 != layer_count
            fused_scale = resolution >= 128
            self.decode_block.append(BEBlock(inputs, outputs, latent_size, has_last_conv, fused_scale=fused_scale))
            inputs = min(maxf, inputs * 2)
            outputs = min(maxf, outputs * 2)
            self.layer_to_resolution[i] = resolution
            resolution /= 2

    def set_noise_mode(self, mode):
        assert mode in ('reference', 'device')
        for b in self.decode_block:
            b.noise_mode = mode

    def _forward_autograd(self, x, block_num):
        c = self.FromRGB.from_rgb
        if not c.implicit_lreq:
            raise NotImplementedError('training path: explicit lreq scaling is not used by the reference (lreq.py:23-24)')
        f = F.leaky_relu(tc.lib_conv2d(x, c.weight, c.bias), 0.2)
        w = torch.tensor(0)
        for i in range(9 - block_num, self.layer_count):
            f, w1, w2 = self.decode_block[i]._forward_autograd(f)
            w_ = torch.cat((w2.view(f.shape[0], 1, 512), w1.view(f.shape[0], 1, 512)), dim=1)
            w = w_ if i == (9 - block_num) else torch.cat((w_, w), dim=1)
        return f, w

    def forward(self, x, block_num=9):
        if _wants_grad(self, x):
            if FUSED_TRAIN and self.startf % 16 == 0:
                if not x.is_cuda:
                    raise ops.DgeError('E_Blur.BE: dge_b200 runs on a B200 only; there is no CPU fallback')
                if not self.FromRGB.from_rgb.implicit_lreq:
                    raise NotImplementedError('training path: explicit lreq scaling is not used by the reference')
                from dge_b200 import train_e
                return train_e.encoder_forward(self, x, block_num)
            return self._forward_autograd(x.float(), block_num)
        ln._guard('E_Blur.BE', x, self.FromRGB.from_rgb.weight)
        f = self.FromRGB.run(x)
        w = torch.tensor(0)
        for i in range(9 - block_num, self.layer_count):
            f, w1, w2 = self.decode_block[i].run(f)
            w_ = torch.cat((w2.view(f.n, 1, 512), w1.view(f.n, 1, 512)), dim=1)
            w = w_ if i == (9 - block_num) else torch.cat((w_, w), dim=1)
        return f.to_nchw(), w
