"""B200-native PGGAN encoder -- drop-in for the reference's `model/E/E_PG.py` (FromRGB :30-37, BEBlock :39-108,
BE :111-164).  Same classes / constructor arguments / state_dict keys.

Reference quirks kept: no `inver_mod` heads; `instance_norm_2` is built with `outputs` channels but applied to
`inputs` channels (affine=False, so only a warning upstream, :53,93); the residual is `IN3_affine(conv_3(x))`,
added BEFORE the leaky-ReLU, pooling comes last (:95-103); `BE.forward` computes `new_final(...)`, discards it and
returns `(tensor(0), tensor(0))` (:161-164).  `BE.features(x)` is an extension returning that discarded tensor
(used by the parity tests).
"""
import torch
import torch.nn as nn

import model.utils.lreq as ln
from model.utils.net import FromRGB  # same 1x1 conv + lrelu as E_PG.FromRGB (:30-37)
from dge_b200 import ops

DEFAULT_PLANES = 2


class BEBlock(nn.Module):
    def __init__(self, inputs, outputs, latent_size, has_second_conv=True, fused_scale=True):
        super().__init__()
        self.has_second_conv = has_second_conv
        self.noise_weight_1 = nn.Parameter(torch.zeros(1, inputs, 1, 1))
        self.bias_1 = nn.Parameter(torch.zeros(1, inputs, 1, 1))
        self.instance_norm_1 = nn.InstanceNorm2d(inputs, affine=False, eps=1e-8)
        self.conv_1 = ln.Conv2d(inputs, inputs, 3, 1, 1, bias=False)
        self.noise_weight_2 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.bias_2 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.instance_norm_2 = nn.InstanceNorm2d(outputs, affine=False, eps=1e-8)
        if has_second_conv:
            if fused_scale:
                raise NotImplementedError('fused_scale=True is never used by E_PG.BE (:127)')
            self.conv_2 = ln.Conv2d(inputs, outputs, 3, 1, 1, bias=False)
        self.fused_scale = fused_scale
        self.inputs = inputs
        self.outputs = outputs
        if self.inputs != self.outputs:
            self.conv_3 = ln.Conv2d(inputs, outputs, 1, 1, 0)
            self.instance_norm_3 = nn.InstanceNorm2d(outputs, affine=True, eps=1e-8)
        self.planes = DEFAULT_PLANES
        self.noise_mode = 'reference'

    def _noise(self, n, h, w, device):
        if self.noise_mode == 'device':
            return torch.randn([n, 1, h, w], device=device)
        return torch.randn([n, 1, h, w]).to(device)

    def run(self, x):
        n, c, h, w = x.n, x.c, x.h, x.w
        dev = x.t.device
        _, mr1 = ops.instance_stats(x, self.instance_norm_1.eps)
        xn, _ = ops.instance_norm(x, mr1, planes=self.planes)
        y1 = ops.conv(xn, self.conv_1.packed(self.planes), c, ops.CONV_3X3, noise=self._noise(n, h, w, dev),
                      noise_batched=True, noise_w=self.noise_weight_1.detach().view(-1),
                      bias=self.bias_1.detach().view(-1), slope=0.2, out_f32b=True)['f32b']          # :84-88
        if not self.has_second_conv:
            return y1
        _, mr2 = ops.instance_stats(y1, self.instance_norm_2.eps)
        y1n, _ = ops.instance_norm(y1, mr2, planes=self.planes)                                       # :97
        if self.inputs != self.outputs:
            r = ops.conv(ops.f32b_to_act(x, self.planes), self.conv_3.packed(self.planes), self.outputs,
                         ops.CONV_1X1, bias=self.conv_3.scaled_bias(), out_f32b=True)['f32b']         # :100
            _, mr3 = ops.instance_stats(r, self.instance_norm_3.eps)
            _, res = ops.instance_norm(r, mr3, out_act=False, out_f32b=True, gamma=self.instance_norm_3.weight,
                                       beta=self.instance_norm_3.bias)                                # :101
        else:
            res = x
        y2 = ops.conv(y1n, self.conv_2.packed(self.planes), self.outputs, ops.CONV_3X3,
                      noise=self._noise(n, h, w, dev), noise_batched=True,
                      noise_w=self.noise_weight_2.detach().view(-1), bias=self.bias_2.detach().view(-1),
                      preact_add=res, slope=0.2, out_f32b=True)['f32b']                               # :98-103
        return ops.blend(y2, y2, 1.0, 0.0, pool=True)                                                 # :104-105

    def forward(self, x):
        ln._guard('E_PG.BEBlock', x, self.conv_1.weight)
        return self.run(ops.nchw_to_f32b(x.float())).to_nchw(), 0, 0


class BE(nn.Module):
    def __init__(self, startf=16, maxf=512, layer_count=9, latent_size=512, channels=3, pggan=False):
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
        self.pggan = pggan
        if pggan:
            self.new_final = ln.Linear(512 * 16, latent_size, gain=1)

    def set_noise_mode(self, mode):
        assert mode in ('reference', 'device')
        for b in self.decode_block:
            b.noise_mode = mode

    def features(self, x, block_num=9):
        """Extension: the tensor the reference computes and then throws away (E_PG.py:153-163)."""
        ln._guard('E_PG.BE', x, self.FromRGB.from_rgb.weight)
        f = self.FromRGB.run(x)
        for i in range(9 - block_num, self.layer_count):
            f = self.decode_block[i].run(f)
        out = f.to_nchw()
        if self.pggan:
            out = self.new_final(out.view(out.shape[0], -1))
        return out

    def forward(self, x, block_num=9):
        self.features(x, block_num)
        return torch.tensor(0), torch.tensor(0)
