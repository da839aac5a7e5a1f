"""B200-native StyleGAN1 generator -- drop-in for the hot-path classes of the reference's `model/stylegan1/net.py`:
DecodeBlock (:110-169), ToRGB (:244-253), Generator (:256-362, `decode` :331-336), Mapping (:441-466),
plus pixel_norm / style_mod / upscale2d / downscale2d / Blur (:28-58).

Same constructor arguments, state_dict keys (`const`, `decode_block.{i}.*` incl. the `blur.weight`
buffer and the `[in, out, 3, 3]` transposed-conv weights of the fused-scale blocks, `to_rgb.{i}.to_rgb.*`), and the
reference's quirks: the first block runs on `const` with batch 1, so its first noise draw is `[1,1,4,4]` and shared by
the batch; noise is `torch.randn` on the CPU per stage (:148,160).

Per block:  [nearest x2 -> tcgen05 3x3 conv | tcgen05 stride-2 transposed conv]  -> one kernel (2x2 box for the
4-tap `transform_kernel`, 3x3 Blur, noise, bias, lrelu) -> stats -> one kernel (instance norm + style_mod + bf16 split
[+ upsample for the next block]) -> tcgen05 3x3 conv with noise/bias/lrelu epilogue -> stats -> norm+style_mod.
`decode2`/`decode3`/`forward_double` (blend / blob-removal experiments) are not on the inversion path.

Training (E_align_s2.py:158: `Gs.forward(w2, lod)` with w2 from the encoder under autograd): when `styles` requires
grad, `Generator.forward` records a differentiable graph w.r.t. `styles` (`_decode_autograd`; the generator is frozen,
its parameters are constants): the stride-1 3x3 convs run forward and backward on the tcgen05 kernels
(dge_b200.autograd.conv2d), the stride-2 transposed convs of the fused-scale blocks and the point-wise steps are torch
CUDA ops in this build.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

import model.utils.lreq as ln
from dge_b200 import autograd as tc
from dge_b200 import ops
from dge_b200 import train_g1

FUSED_TRAIN = True     # False: `_decode_autograd` (separate torch nodes), the cross-check of the fused node

DEFAULT_PLANES = 2


def pixel_norm(x, epsilon=1e-8):
    ln._guard('pixel_norm', x)
    return ops.pixel_norm(x.float(), epsilon)


def upscale2d(x, factor=2):
    ln._guard('upscale2d', x)
    if factor != 2:
        raise NotImplementedError('upscale2d: factor 2 only')
    return ops.upsample_nearest_nchw(x.float())


def downscale2d(x, factor=2):
    from model.utils.net import downscale2d as _d
    return _d(x, factor)


class Blur(nn.Module):
    """Holds the reference's `weight` buffer ([C,1,3,3] of [1,2,1]^2/16); the filtering is fused (dge_sg1_post)."""

    def __init__(self, channels):
        super().__init__()
        f = np.array([1, 2, 1], dtype=np.float32)
        f = f[:, np.newaxis] * f[np.newaxis, :]
        f /= np.sum(f)
        self.register_buffer('weight', torch.Tensor(f).view(1, 1, 3, 3).repeat(channels, 1, 1, 1))
        self.groups = channels

    def forward(self, x):
        ln._guard('Blur', x)
        n, c, h, w = x.shape
        return ops.sg1_post(ops.nchw_to_f32b(x.float()), 0, n, c, h, w, slope=1.0).to_nchw()


class DecodeBlock(nn.Module):
    def __init__(self, inputs, outputs, latent_size, has_first_conv=True, fused_scale=True):
        super().__init__()
        self.has_first_conv = has_first_conv
        self.inputs = inputs
        self.outputs = outputs
        self.fused_scale = fused_scale
        if has_first_conv:
            if fused_scale:
                self.conv_1 = ln.ConvTranspose2d(inputs, outputs, 3, 2, 1, bias=False, transform_kernel=True)
            else:
                self.conv_1 = ln.Conv2d(inputs, outputs, 3, 1, 1, bias=False)
        self.blur = Blur(outputs)
        self.noise_weight_1 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.bias_1 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.instance_norm_1 = nn.InstanceNorm2d(outputs, affine=False, eps=1e-8)
        self.style_1 = ln.Linear(latent_size, 2 * outputs, gain=1)
        self.conv_2 = ln.Conv2d(outputs, outputs, 3, 1, 1, bias=False)
        self.noise_weight_2 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.bias_2 = nn.Parameter(torch.zeros(1, outputs, 1, 1))
        self.instance_norm_2 = nn.InstanceNorm2d(outputs, affine=False, eps=1e-8)
        self.style_2 = ln.Linear(latent_size, 2 * outputs, gain=1)
        self.planes = DEFAULT_PLANES
        self.noise_mode = 'reference'

    def _noise(self, n, h, w, device):
        if self.noise_mode == 'device':
            return torch.randn([n, 1, h, w], device=device)
        return torch.randn([n, 1, h, w]).to(device)

    def _conv1_packed(self):
        c = self.conv_1
        key = (ops.weight_key(c.weight), self.planes)
        if getattr(self, '_c1_key', None) != key:
            scale = 1.0 if c.implicit_lreq else c.std
            if self.fused_scale:     # weight is [in, out, k, k] and used un-flipped by conv_transpose2d (lreq.py:127-140)
                w = c.weight.detach().permute(1, 0, 2, 3).contiguous()
            else:
                w = c.weight.detach()
            self._c1 = ops.pack_conv_weight(w, scale=scale, planes=self.planes)
            self._c1_key = key
        return self._c1

    def run(self, x, s1, s2, next_up=1, last=False):
        """x: for the first block the F32B `const` (batch 1); otherwise the ACT operand of conv_1 (already
        nearest-upsampled for the non-fused blocks).  Returns the ACT operand for the next block (upsampled if
        `next_up == 2`), or the F32B feature map if `last`."""
        n = s1.shape[0]
        nw1, b1 = self.noise_weight_1.detach().view(-1), self.bias_1.detach().view(-1)
        nw2, b2 = self.noise_weight_2.detach().view(-1), self.bias_2.detach().view(-1)
        if not self.has_first_conv:
            dev, h, w, nb = x.t.device, x.h, x.w, x.n
            y = ops.sg1_post(x, 2, nb, self.outputs, h, w, noise=self._noise(nb, h, w, dev), noise_w=nw1, bias=b1)
        elif self.fused_scale:
            dev, h, w, nb = x.t.device, 2 * x.h, 2 * x.w, x.n
            raw = ops.conv(x, self._conv1_packed(), self.outputs, ops.CONV_UP3X3)['raw_up']
            y = ops.sg1_post(raw, 1, nb, self.outputs, h, w, noise=self._noise(nb, h, w, dev), noise_w=nw1, bias=b1)
        else:
            dev, h, w, nb = x.t.device, x.h, x.w, x.n
            c1 = ops.conv(x, self._conv1_packed(), self.outputs, ops.CONV_3X3, out_f32b=True)['f32b']
            y = ops.sg1_post(c1, 0, nb, self.outputs, h, w, noise=self._noise(nb, h, w, dev), noise_w=nw1, bias=b1)
        _, mr1 = ops.instance_stats(y, self.instance_norm_1.eps)
        st1 = ops.dense(s1.float().contiguous(), self.style_1.weight, self.style_1.bias)
        xa, _ = ops.instance_norm_style(y, mr1, st1, n, planes=self.planes)                    # :154-156
        y2 = ops.conv(xa, self.conv_2.packed(self.planes), self.outputs, ops.CONV_3X3, noise=self._noise(n, h, w, dev),
                      noise_batched=True, noise_w=nw2, bias=b2, slope=0.2, out_f32b=True)['f32b']   # :158-164
        _, mr2 = ops.instance_stats(y2, self.instance_norm_2.eps)
        st2 = ops.dense(s2.float().contiguous(), self.style_2.weight, self.style_2.bias)
        act, f = ops.instance_norm_style(y2, mr2, st2, n, up=next_up, planes=self.planes, out_act=not last,
                                         out_f32b=last)                                       # :165-167
        return f if last else act

    def _forward_autograd(self, x, s1, s2):
        """Differentiable w.r.t. x, s1, s2 (parameters are constants): net.py:141-169 on NCHW tensors."""
        def wgt(conv):
            return conv.weight.detach() if conv.implicit_lreq else conv.weight.detach() * conv.std

        def style_mod(v, lin, s):                                                          # :32-34
            lb = None if lin.bias is None else (lin.bias.detach() if lin.implicit_lreq else lin.bias.detach() * lin.lrmul)
            st = F.linear(s, lin.weight.detach() if lin.implicit_lreq else lin.weight.detach() * lin.std, lb)
            st = st.view(st.shape[0], 2, v.shape[1], 1, 1)
            return torch.addcmul(st[:, 1], v, st[:, 0] + 1)

        if self.has_first_conv:
            if self.fused_scale:                                                           # lreq.py:127-140
                w = F.pad(wgt(self.conv_1), (1, 1, 1, 1))
                w = w[:, :, 1:, 1:] + w[:, :, :-1, 1:] + w[:, :, 1:, :-1] + w[:, :, :-1, :-1]
                x = tc.lib_conv_transpose2d(x, w, stride=2, padding=1)
            else:                                                                          # upscale2d :37-43, conv :145
                x = tc.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), wgt(self.conv_1), self.planes)
            x = tc.lib_conv2d(x, self.blur.weight, groups=self.blur.groups, padding=1)          # :48-58
        n, _, h, w_ = x.shape
        x = torch.addcmul(x, self.noise_weight_1.detach(), self._noise(n, h, w_, x.device))    # :148 (batch 1 in block 0)
        x = F.instance_norm(F.leaky_relu(x + self.bias_1.detach(), 0.2), eps=self.instance_norm_1.eps)
        x = style_mod(x, self.style_1, s1)                                                 # :154-156
        x = tc.conv2d(x, wgt(self.conv_2), self.planes)                                    # :158
        n = x.shape[0]
        x = torch.addcmul(x, self.noise_weight_2.detach(), self._noise(n, h, w_, x.device))    # :160
        x = F.instance_norm(F.leaky_relu(x + self.bias_2.detach(), 0.2), eps=self.instance_norm_2.eps)
        return style_mod(x, self.style_2, s2)                                              # :165-167

    def forward(self, x, s1, s2):
        """Reference signature: NCHW in -> NCHW out."""
        ln._guard('DecodeBlock', x, s1, s2, self.conv_2.weight)
        f = ops.nchw_to_f32b(x.float())
        if self.has_first_conv:
            if self.fused_scale:
                xin = ops.f32b_to_act(f, self.planes)
            else:
                xin = ops.f32b_to_act(ops.nchw_to_f32b(ops.upsample_nearest_nchw(x.float())), self.planes)
        else:
            xin = f
        return self.run(xin, s1, s2, last=True).to_nchw()


class FromRGB(nn.Module):
    def __init__(self, channels, outputs):
        super().__init__()
        self.from_rgb = ln.Conv2d(channels, outputs, 1, 1, 0)

    def forward(self, x):
        ln._guard('FromRGB', x, self.from_rgb.weight)
        return ops.from_rgb(x.float(), self.from_rgb.weight, self.from_rgb.scaled_bias(), slope=0.2).to_nchw()


class ToRGB(nn.Module):
    def __init__(self, inputs, channels):
        super().__init__()
        self.inputs = inputs
        self.channels = channels
        self.to_rgb = ln.Conv2d(inputs, channels, 1, 1, 0, gain=1)

    def run(self, f):
        return ops.to_rgb_f32b(f, self.to_rgb.weight, self.to_rgb.scaled_bias())

    def forward(self, x):
        ln._guard('ToRGB', x, self.to_rgb.weight)
        return self.run(ops.nchw_to_f32b(x.float()))


class Generator(nn.Module):
    def __init__(self, startf=32, maxf=256, layer_count=3, latent_size=128, channels=3):
        super().__init__()
        self.maxf = maxf
        self.startf = startf
        self.layer_count = layer_count
        self.channels = channels
        self.latent_size = latent_size
        mul = 2 ** (self.layer_count - 1)
        inputs = min(self.maxf, startf * mul)
        self.const = Parameter(torch.Tensor(1, inputs, 4, 4))
        self.zeros = torch.zeros(1, 1, 1, 1)
        init.ones_(self.const)
        self.layer_to_resolution = [0 for _ in range(layer_count)]
        resolution = 2
        self.style_sizes = []
        to_rgb = nn.ModuleList()
        self.decode_block = nn.ModuleList()
        for i in range(self.layer_count):
            outputs = min(self.maxf, startf * mul)
            has_first_conv = i != 0
            fused_scale = resolution * 2 >= 128
            block = DecodeBlock(inputs, outputs, latent_size, has_first_conv, fused_scale=fused_scale)
            resolution *= 2
            self.layer_to_resolution[i] = resolution
            self.style_sizes += [2 * (inputs if has_first_conv else outputs), 2 * outputs]
            to_rgb.append(ToRGB(outputs, channels))
            self.decode_block.append(block)
            inputs = outputs
            mul //= 2
        self.to_rgb = to_rgb

    def set_noise_mode(self, mode):
        assert mode in ('reference', 'device')
        for b in self.decode_block:
            b.noise_mode = mode

    def _decode_autograd(self, styles, lod):
        """Training path of `decode` (:331-336): recorded for backward w.r.t. `styles`; see the module docstring."""
        styles = styles.float()
        x = self.const.detach().float()
        for i in range(lod + 1):
            x = self.decode_block[i]._forward_autograd(x, styles[:, 2 * i + 0], styles[:, 2 * i + 1])
        c = self.to_rgb[lod].to_rgb
        w = c.weight.detach() if c.implicit_lreq else c.weight.detach() * c.std
        return tc.lib_conv2d(x, w, c.scaled_bias())                                             # :244-253

    def decode(self, styles, lod, noise=0):
        if torch.is_grad_enabled() and styles.requires_grad:
            if not styles.is_cuda:
                raise ops.DgeError('Generator.decode: dge_b200 runs on a B200 only; there is no CPU fallback')
            if FUSED_TRAIN and all(b.conv_2.implicit_lreq for b in self.decode_block):
                return train_g1.decode(self, styles, lod)            # one fused node (dge_b200/train_g1.py)
            return self._decode_autograd(styles, lod)
        ln._guard('Generator.decode', styles, self.const)
        x = ops.nchw_to_f32b(self.const.detach().float())
        for i in range(lod + 1):
            blk = self.decode_block[i]
            last = i == lod
            nxt = self.decode_block[i + 1] if not last else None
            next_up = 2 if (nxt is not None and not nxt.fused_scale) else 1
            x = blk.run(x, styles[:, 2 * i + 0], styles[:, 2 * i + 1], next_up=next_up, last=last)
        return self.to_rgb[lod].run(x)

    def forward(self, styles, lod, blend=1, remove_blob=False):
        if remove_blob or blend != 1:
            raise NotImplementedError('dge_b200 implements Generator.decode (blend == 1, remove_blob == False), the '
                                      'path the inversion scripts call (E_align_s2.py:109,158)')
        return self.decode(styles, lod, 1)


def _staged(module, name, t, dev):
    """Device copy of a tensor the caller left on the CPU, cached per version.  The training scripts never move the
    mapping network: `Gm` and `z` stay on the CPU and only the result is `.cuda()`-ed (E_align_s2.py:33-44, 108).  The
    drop-in keeps that call working by staging the (tiny) operands onto the current CUDA device and running the same
    kernels there -- it is not a CPU path: without a CUDA device `ops.lib()` raises."""
    if t is None or t.device == dev:
        return t
    cache = module.__dict__.setdefault('_dge_staged', {})
    key = (ops.weight_key(t), str(dev))
    hit = cache.get(name)
    if hit is None or hit[0] != key:
        hit = (key, t.detach().to(dev))
        cache[name] = hit
    return hit[1]


def _compute_device(t):
    if t.is_cuda:
        return t.device
    ops.lib()                                   # raises DgeError when there is no B200 to run on
    return torch.device('cuda', torch.cuda.current_device())


class MappingBlock(nn.Module):
    def __init__(self, inputs, output, lrmul=0.01):
        super().__init__()
        self.fc = ln.Linear(inputs, output, lrmul=lrmul)

    def forward(self, x):
        dev = _compute_device(x)
        x = x.to(dev)
        fc = self.fc
        w, b = _staged(self, 'w', fc.weight, dev), _staged(self, 'b', fc.bias, dev)
        ln._guard('MappingBlock', x, w)
        if torch.is_grad_enabled() and fc.weight.requires_grad:
            raise NotImplementedError('MappingBlock: the mapping network is forward-only in this build; use torch.no_grad()')
        if fc.implicit_lreq:
            return ops.dense(x.float(), w, b, slope=0.2)
        return ops.dense(x.float(), w, b, wscale=fc.std, bscale=fc.lrmul, slope=0.2)


class Mapping(nn.Module):
    def __init__(self, num_layers=18, mapping_layers=8, latent_size=512, dlatent_size=512, mapping_fmaps=512,
                 trunc_tensor=None):
        super().__init__()
        inputs = latent_size
        self.mapping_layers = mapping_layers
        self.num_layers = num_layers
        for i in range(mapping_layers):
            outputs = dlatent_size if i == mapping_layers - 1 else mapping_fmaps
            setattr(self, "block_%d" % (i + 1), MappingBlock(inputs, outputs, lrmul=0.01))
            inputs = outputs
        self.register_buffer('buffer1', trunc_tensor)

    def forward(self, z, coefs_m=0):
        dev = _compute_device(z)                # z / parameters / coefs may all live on the CPU (see `_staged`)
        x = pixel_norm(z.to(dev))
        for i in range(self.mapping_layers):
            x = getattr(self, "block_%d" % (i + 1))(x)
        x = x.view(x.shape[0], 1, x.shape[1]).repeat(1, self.num_layers, 1)
        if self.buffer1 is not None:
            coefs = coefs_m.to(dev) if torch.is_tensor(coefs_m) else coefs_m
            x = torch.lerp(_staged(self, 'buffer1', self.buffer1.data, dev), x, coefs)   # avg + (styles - avg) * coefs (:464-465)
        return x
