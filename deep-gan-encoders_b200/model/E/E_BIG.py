"""B200-native BigGAN encoder -- drop-in for the reference's `model/E/E_BIG.py` (BigGANBatchNorm :33-82, FromRGB
:84-92, BEBlock :94-169, BE :172-227).  Same classes / constructor arguments / state_dict keys (spectral-norm
`scale.weight_orig/_u/_v`, ...).  `forward(x, cond_vector) -> (c_v [N,256], z [N,128])`.

Block (reference order kept): CBN1 (no activation; running stats are the constant 0/1 buffers, eps 1e-12)
-> conv_1 -> noise, bias, lrelu -> CBN2 -> conv_2 -> noise, bias, lrelu [-> lrelu AGAIN when channels change, :163]
-> + residual (conv_3(CBN3(x)) when channels change) -> 2x2 avg-pool.  `truncation` is hard-wired to 0.4 (:222).

Training: as in `model/E/E.py` -- a call that must be recorded for backward runs ONE fused autograd node per block
(`dge_b200/train_big.py`, `FUSED_TRAIN`): the conditional-BN affines enter the node as [N, C] coefficient tensors, so the
trainable spectral-norm `scale` / `offset` layers (and their power iteration) stay small torch graphs around it.
`_forward_autograd` -- separate torch nodes with the convs on the tcgen05 kernels -- is kept as the cross-check.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401

import model.utils.lreq as ln
from model.E.E import _wants_grad
from model.biggan_generator import BigGANBatchNorm, snlinear  # noqa: F401  (identical classes upstream)
from dge_b200 import autograd as tc
from dge_b200 import ops

DEFAULT_PLANES = 2
FUSED_TRAIN = True      # one fused autograd node per block (dge_b200/train_big.py); False: separate torch nodes


class FromRGB(nn.Module):
    def __init__(self, channels, outputs):
        super().__init__()
        self.from_rgb = torch.nn.Conv2d(channels, outputs, 1, 1, 0)

    def run(self, x):
        return ops.from_rgb(x.float(), self.from_rgb.weight.detach(), self.from_rgb.bias.detach(), slope=0.2)

    def forward(self, x):
        ln._guard('E_BIG.FromRGB', x, self.from_rgb.weight)
        return self.run(x).to_nchw()


class BEBlock(nn.Module):
    def __init__(self, inputs, outputs, latent_size, has_second_conv=True, fused_scale=True):
        super().__init__()
        self.has_second_conv = has_second_conv
        self.noise_weight_1 = nn.Parameter(torch.zeros(1, inputs, 1, 1))
        self.bias_1 = nn.Parameter(torch.zeros(1, inputs, 1, 1))
        self.batch_norm_1 = BigGANBatchNorm(inputs, condition_vector_dim=256, n_stats=51, eps=1e-12, conditional=True)
        self.conv_1 = ln.Conv2d(inputs, inputs, 3, 1, 1, bias=False)
        self.noise_weight_2 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.bias_2 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.batch_norm_2 = BigGANBatchNorm(inputs, condition_vector_dim=256, n_stats=51, eps=1e-12, conditional=True)
        if has_second_conv:
            if fused_scale:
                raise NotImplementedError('fused_scale=True is never used by E_BIG.BE (:189)')
            self.conv_2 = ln.Conv2d(inputs, outputs, 3, 1, 1, bias=False)
        self.fused_scale = fused_scale
        self.inputs = inputs
        self.outputs = outputs
        if self.inputs != self.outputs:
            self.batch_norm_3 = BigGANBatchNorm(inputs, condition_vector_dim=256, n_stats=51, eps=1e-12,
                                                conditional=True)
            self.conv_3 = ln.Conv2d(inputs, outputs, 1, 1, 0)
        self.planes = DEFAULT_PLANES
        self.noise_mode = 'reference'

    def _noise(self, n, h, w, device):
        if self.noise_mode == 'device':
            return torch.randn([n, 1, h, w], device=device)
        return torch.randn([n, 1, h, w]).to(device)

    def run(self, x, cond_vector, truncation=0.4):
        n, c, h, w = x.n, x.c, x.h, x.w
        dev = x.t.device
        a, b = self.batch_norm_1.coeffs(truncation, cond_vector, n)
        xn, _ = ops.affine_act(x, a, b, relu=False, planes=self.planes)                                    # :134
        y1 = ops.conv(xn, self.conv_1.packed(self.planes), c, ops.CONV_3X3, noise=self._noise(n, h, w, dev),
                      noise_batched=True, noise_w=self.noise_weight_1.detach().view(-1),
                      bias=self.bias_1.detach().view(-1), slope=0.2, out_f32b=True)['f32b']                # :135-138
        if not self.has_second_conv:
            return y1
        a, b = self.batch_norm_2.coeffs(truncation, cond_vector, n)
        y1n, _ = ops.affine_act(y1, a, b, relu=False, planes=self.planes)                                  # :150
        if self.inputs != self.outputs:
            a, b = self.batch_norm_3.coeffs(truncation, cond_vector, n)
            rn, _ = ops.affine_act(x, a, b, relu=False, planes=self.planes)                                # :160
            res = ops.conv(rn, self.conv_3.packed(self.planes), self.outputs, ops.CONV_1X1,
                           bias=self.conv_3.scaled_bias(), out_f32b=True)['f32b']                          # :161
            slope = 0.2 * 0.2          # lrelu applied twice (:158,163): negative slope 0.04
        else:
            res, slope = x, 0.2
        y2 = ops.conv(y1n, self.conv_2.packed(self.planes), self.outputs, ops.CONV_3X3,
                      noise=self._noise(n, h, w, dev), noise_batched=True,
                      noise_w=self.noise_weight_2.detach().view(-1), bias=self.bias_2.detach().view(-1), slope=slope,
                      blend_src=res, blend_a=1.0, blend_b=1.0, out_f32b=True)['f32b']                      # :151-164
        return ops.blend(y2, y2, 1.0, 0.0, pool=3)                                                         # :165-166

    def _forward_autograd(self, x, cond_vector, truncation=0.4):
        """Differentiable form of the block (E_BIG.py:129-169) on NCHW tensors."""
        n, c, h, w = x.shape
        dev = x.device
        res = x
        y = self.batch_norm_1._forward_autograd(x, truncation, cond_vector, frozen=False)                 # :134
        y = tc.conv2d(y, self.conv_1.weight, self.planes)                                                 # :135
        y = F.leaky_relu(torch.addcmul(y, self.noise_weight_1, self._noise(n, h, w, dev)) + self.bias_1, 0.2)
        if not self.has_second_conv:
            return y
        y = self.batch_norm_2._forward_autograd(y, truncation, cond_vector, frozen=False)                 # :150
        y = tc.conv2d(y, self.conv_2.weight, self.planes)                                                 # :151
        y = F.leaky_relu(torch.addcmul(y, self.noise_weight_2, self._noise(n, h, w, dev)) + self.bias_2, 0.2)
        if self.inputs != self.outputs:
            res = self.batch_norm_3._forward_autograd(res, truncation, cond_vector, frozen=False)         # :160
            res = tc.conv2d(res, self.conv_3.weight, self.planes) + self.conv_3.bias.view(1, -1, 1, 1)    # :161
            y = F.leaky_relu(y, 0.2)                                                                      # :163
        return F.avg_pool2d(y + res, 2, 2)                                                                # :164-166

    def forward(self, x, cond_vector, truncation=0.4):
        if _wants_grad(self, x):
            return self._forward_autograd(x.float(), cond_vector.float(), truncation), 0, 0
        ln._guard('E_BIG.BEBlock', x, cond_vector, self.conv_1.weight)
        return self.run(ops.nchw_to_f32b(x.float()), cond_vector.float().contiguous(), truncation).to_nchw(), 0, 0


class BE(nn.Module):
    def __init__(self, startf=16, maxf=512, layer_count=9, latent_size=512, channels=3, pggan=False, biggan=False):
        super().__init__()
        self.maxf = maxf
        self.startf = startf
        self.latent_size = latent_size
        self.decode_block = nn.ModuleList()
        self.layer_count = layer_count
        inputs = startf
        outputs = startf * 2
        self.FromRGB = FromRGB(channels, inputs)
        for i in range(layer_count):
            has_second_conv = i + 1 != layer_count
            self.decode_block.append(BEBlock(inputs, outputs, latent_size, has_second_conv, fused_scale=False))
            inputs = min(maxf, inputs * 2)
            outputs = min(maxf, outputs * 2)
        self.biggan = biggan
        if biggan:
            self.new_final_1 = ln.Linear(8192, 256, gain=1)
            self.new_final_2 = ln.Linear(256, 128, gain=1)

    def set_noise_mode(self, mode):
        assert mode in ('reference', 'device')
        for b in self.decode_block:
            b.noise_mode = mode

    def features(self, x, cond_vector, block_num=9):
        """Extension: the [N, C, 4, 4] feature map before the two heads (parity tests of small configurations)."""
        ln._guard('E_BIG.BE', x, cond_vector, self.FromRGB.from_rgb.weight)
        cv = cond_vector.float().contiguous()
        f = self.FromRGB.run(x)
        for i in range(9 - block_num, self.layer_count):
            f = self.decode_block[i].run(f, cv, truncation=0.4)
        return f.to_nchw()

    def _features_autograd(self, x, cond_vector, block_num=9):
        if FUSED_TRAIN and self.startf % 16 == 0:
            from dge_b200 import train_big
            return train_big.ebig_features(self, x, cond_vector, block_num)
        cv = cond_vector.float()
        c = self.FromRGB.from_rgb
        f = F.leaky_relu(tc.lib_conv2d(x.float(), c.weight, c.bias), 0.2)                                      # :84-92
        for i in range(9 - block_num, self.layer_count):
            f = self.decode_block[i]._forward_autograd(f, cv, truncation=0.4)
        return f

    def forward(self, x, cond_vector, block_num=9):
        if _wants_grad(self, x):
            x = self._features_autograd(x, cond_vector, block_num)
            if self.biggan:
                c_v = F.linear(x.reshape(x.shape[0], -1), self.new_final_1.weight, self.new_final_1.bias)
                z = F.linear(c_v, self.new_final_2.weight, self.new_final_2.bias)
            return c_v, z
        x = self.features(x, cond_vector, block_num)
        if self.biggan:
            c_v = self.new_final_1(x.view(x.shape[0], -1))
            z = self.new_final_2(c_v)
        return c_v, z            # (UnboundLocalError without biggan=True, exactly as upstream :223-227)
