"""B200-native case-1 encoder -- drop-in for the reference's `model/E/E.py` (BEBlock :16-85, BE :88-135).

Same classes, constructor arguments, `forward` signature/returns and state_dict keys (SURVEY
Appendix A).  Per block the reference issues ~25 ATen kernels; here it is
  stats -> GEMV -> IN-apply(+bf16 split) -> tcgen05 conv [noise, bias, lrelu fused]      (x2)
  -> tcgen05 1x1 residual conv whose epilogue does the 2x2 pools and the 0.111/0.889 blend.
Noise: the reference draws `torch.randn([N,1,H,W])` on the CPU inside each conv stage (E.py:60,73);
`noise_mode = 'reference'` (default) does exactly that (same RNG stream => bit-identical noise),
`'device'` draws on the GPU instead (no H2D copy; different stream).

Training (`loss.backward()` in E_align_s2.py:205, embedding_img.py:100-128): when autograd is recording and a
parameter requires grad, `forward` records ONE fused autograd node per block (dge_b200/train_e.py): the forward of a
node is the same kernel chain as inference, its backward is 11 dge_b200 launches (tcgen05 data / weight gradients and the
fused point-wise backward kernels of csrc/train_bwd.cu).  Same noise draws, same return values.
`_forward_autograd` is the same computation as ~25 separate torch nodes per block with only the convs on the tcgen05
kernels: it is kept as the cross-check of the fused path (tests) and is selected with `FUSED_TRAIN = False`.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import model.utils.lreq as ln
from model.utils.net import FromRGB
from dge_b200 import autograd as tc
from dge_b200 import ops

from dge_b200 import train_e

DEFAULT_PLANES = 2
FUSED_TRAIN = True     # False: the unfused torch-node graph (`_forward_autograd`)


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
        if has_last_conv:
            if fused_scale:
                # E.py:32-33 -- never taken by E.BE (fused_scale is hard-wired False, :106)
                raise NotImplementedError('fused_scale=True (strided conv) belongs to model/E/E_Blur.py')
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
        return torch.randn([n, 1, h, w]).to(device)   # E.py:60 -- CPU draw, then H2D

    def run(self, x, stats=None):
        """x: F32B [N, inputs, H, W] -> (F32B out, w1, w2).  `stats` = (style, mean_rstd) of x when its producer
        already computed them (FromRGB)."""
        n, c, h, w = x.n, x.c, x.h, x.w
        dev = x.t.device
        eps = self.instance_norm_1.eps
        style1, mr1 = stats if stats is not None else ops.instance_stats(x, eps)   # E.py:51-53 + IN stats
        w1 = ops.dense(style1, self.inver_mod1.weight, self.inver_mod1.bias)      # :54
        rp = None
        if self.has_last_conv and self.inputs != self.outputs:
            xn, rp = ops.instance_norm_pool(x, mr1, planes=self.planes)            # :58 and :78 in one pass over x
        else:
            xn, _ = ops.instance_norm(x, mr1, planes=self.planes)                  # :58
        y1 = ops.conv(xn, self.conv_1.packed(self.planes), c, ops.CONV_3X3, noise=self._noise(n, h, w, dev),
                      noise_batched=True, noise_w=self.noise_weight_1.detach().view(-1),
                      bias=self.bias_1.detach().view(-1), slope=0.2, out_f32b=True)['f32b']      # :59-62
        style2, mr2 = ops.instance_stats(y1, self.instance_norm_2.eps)             # :64-66
        w2 = ops.dense(style2, self.inver_mod2.weight, self.inver_mod2.bias)      # :67
        if self.has_last_conv:
            y1n, _ = ops.instance_norm(y1, mr2, planes=self.planes)                # :69
            # conv_2's output is only ever consumed through avg_pool2d (:76-77, 84): the 2x2 mean is taken in the
            # conv epilogue, so the full-resolution tensor is never written or re-read
            y2p = ops.conv(y1n, self.conv_2.packed(self.planes), self.outputs, ops.CONV_3X3,
                           noise=self._noise(n, h, w, dev), noise_batched=True,
                           noise_w=self.noise_weight_2.detach().view(-1), bias=self.bias_2.detach().view(-1),
                           slope=0.2, out_f32b_pool=True)['f32b_pool']             # :72-75 + :76-77
            if self.inputs != self.outputs:
                out = ops.conv(rp, self.conv_3.packed(self.planes), self.outputs, ops.CONV_1X1,
                               bias=self.conv_3.scaled_bias(), blend_src=y2p, blend_pool=False, blend_a=0.111,
                               blend_b=0.889, out_f32b=True)['f32b']               # :81-84
            else:
                out = ops.blend(y2p, x, 0.111, 0.889, pool=2)
        else:
            _, y1n = ops.instance_norm(y1, mr2, out_act=False, out_f32b=True)      # :69
            if self.inputs != self.outputs:
                out = ops.conv(ops.f32b_to_act(x, self.planes), self.conv_3.packed(self.planes), self.outputs,
                               ops.CONV_1X1, bias=self.conv_3.scaled_bias(), blend_src=y1n, blend_pool=False,
                               blend_a=0.111, blend_b=0.889, out_f32b=True)['f32b']
            else:
                out = ops.blend(y1n, x, 0.111, 0.889, pool=False)
        return out, w1, w2

    def _forward_autograd(self, x):
        """Differentiable form of the block (E.py:50-85) on NCHW tensors; convs on the tensor-core kernels."""
        n, c, h, w = x.shape
        dev = x.device
        w1 = F.linear(_mean_std(x), self.inver_mod1.weight, self.inver_mod1.bias)                 # :51-54
        res = x
        y = tc.conv2d(F.instance_norm(x, eps=self.instance_norm_1.eps), self.conv_1.weight, self.planes)   # :58-59
        y = F.leaky_relu(torch.addcmul(y, self.noise_weight_1, self._noise(n, h, w, dev)) + self.bias_1, 0.2)  # :60-62
        w2 = F.linear(_mean_std(y), self.inver_mod2.weight, self.inver_mod2.bias)                 # :64-67
        y = F.instance_norm(y, eps=self.instance_norm_2.eps)                                       # :69
        if self.has_last_conv:
            y = tc.conv2d(y, self.conv_2.weight, self.planes)                                      # :72
            y = F.leaky_relu(torch.addcmul(y, self.noise_weight_2, self._noise(n, h, w, dev)) + self.bias_2, 0.2)
            y = F.avg_pool2d(y, 2, 2)                                                              # :76-77
            res = F.avg_pool2d(res, 2, 2)                                                          # :78
        if self.inputs != self.outputs:
            res = tc.conv2d(res, self.conv_3.weight, self.planes) + self.conv_3.bias.view(1, -1, 1, 1)   # :81-82
        return 0.111 * y + 0.889 * res, w1, w2                                                     # :84

    def forward(self, x):
        """Reference signature: NCHW in -> (NCHW out, w1, w2)."""
        if _wants_grad(self, x):
            if FUSED_TRAIN:
                out_t, w1, w2 = train_e.block_forward(self, train_e.nchw_to_f32b(x.float()))
                return train_e.f32b_to_nchw(out_t), w1, w2
            return self._forward_autograd(x.float())
        ln._guard('BEBlock', x, self.conv_1.weight)
        out, w1, w2 = self.run(ops.nchw_to_f32b(x.float()))
        return out.to_nchw(), w1, w2


def _mean_std(x):
    """[N, 2C] = per-channel mean || biased std over (H, W), no epsilon (E.py:51-53, 64-66)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    std = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True).sqrt()
    return torch.cat((mean, std), dim=1).flatten(1)


def _wants_grad(module, x):
    """True when this call must be recorded for backward: autograd is on and the input or a parameter needs grad."""
    if not torch.is_grad_enabled():
        return False
    if not x.is_cuda:
        raise ops.DgeError(f'{type(module).__name__}: dge_b200 runs on a B200 only (got a {x.device} tensor); '
                           'no CPU fallback')
    return x.requires_grad or any(p.requires_grad for p in module.parameters())


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
            block = BEBlock(inputs, outputs, latent_size, has_last_conv, fused_scale=False)
            inputs = min(maxf, inputs * 2)
            outputs = min(maxf, outputs * 2)
            self.layer_to_resolution[i] = resolution
            resolution /= 2
            self.decode_block.append(block)

    def set_noise_mode(self, mode):
        assert mode in ('reference', 'device')
        for b in self.decode_block:
            b.noise_mode = mode

    def _forward_autograd(self, x, block_num):
        """Training path: same data flow as `forward`, recorded for backward (see the module docstring)."""
        c = self.FromRGB.from_rgb
        if not c.implicit_lreq:
            raise NotImplementedError('training path: explicit lreq scaling is not used by the reference (lreq.py:23-24)')
        f = F.leaky_relu(tc.lib_conv2d(x, c.weight, c.bias), 0.2)          # net.py:231-240 (3 input channels: point-wise)
        w = torch.tensor(0)
        for i in range(9 - block_num, self.layer_count):
            f, w1, w2 = self.decode_block[i]._forward_autograd(f)
            w_ = torch.cat((w2.view(f.shape[0], 1, 512), w1.view(f.shape[0], 1, 512)), dim=1)      # E.py:131
            w = w_ if i == (9 - block_num) else torch.cat((w_, w), dim=1)
        return f, w

    def forward(self, x, block_num=9):
        if _wants_grad(self, x):
            if FUSED_TRAIN and self.FromRGB.from_rgb.implicit_lreq:
                return train_e.encoder_forward(self, x, block_num)
            return self._forward_autograd(x.float(), block_num)
        ln._guard('BE', x, self.FromRGB.from_rgb.weight)
        first = 9 - block_num
        eps0 = self.decode_block[first].instance_norm_1.eps if first < self.layer_count else 1e-8
        f, style0, mr0 = self.FromRGB.run_with_stats(x, eps0)
        w = torch.tensor(0)
        for i in range(first, self.layer_count):
            f, w1, w2 = self.decode_block[i].run(f, stats=(style0, mr0) if i == first else None)
            w_ = torch.cat((w2.view(f.n, 1, 512), w1.view(f.n, 1, 512)), dim=1)   # E.py:131 (512 is hard-coded)
            w = w_ if i == (9 - block_num) else torch.cat((w_, w), dim=1)
        return f.to_nchw(), w
