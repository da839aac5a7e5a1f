"""B200-native PGGAN generator -- drop-in for the reference's `model/pggan/pggan_generator.py`
(PGGANGenerator :28-204, PixelNormLayer :207-216, UpsamplingLayer :219-233, ConvBlock :236-339).

Same constructor/forward signatures, result dict ('z', 'label', 'image'), state_dict keys (`lod`,
`layer{i}.weight/bias`, `output{k}.weight/bias`) and reference quirks: z is pixel-normed twice (:160 + layer0),
every ConvBlock (ToRGB included) pixel-norms its input (:320), one `print(x.shape)` per resolution (:196).
Per block the chain pixel-norm -> nearest x2 -> conv*wscale -> bias -> lrelu is two kernels: a pixel-norm(+upsample)
producer that writes the bf16 hi/lo conv operand, and the tcgen05 conv with the bias/lrelu epilogue.
"""
import numpy as np
import torch
import torch.nn as nn

from dge_b200 import ops

__all__ = ['PGGANGenerator']

_RESOLUTIONS_ALLOWED = [8, 16, 32, 64, 128, 256, 512, 1024]
_INIT_RES = 4
_WSCALE_GAIN = np.sqrt(2.0)
DEFAULT_PLANES = 2


def _guard(name, *tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ops.DgeError(f'{name}: dge_b200 runs on a B200 only (got a {t.device} tensor); no CPU fallback')
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(f'{name}: dge_b200 kernels are forward-only in this build; use torch.no_grad()')


class PGGANGenerator(nn.Module):
    def __init__(self, resolution, z_space_dim=512, image_channels=3, final_tanh=False, label_size=0,
                 fused_scale=False, use_wscale=True, fmaps_base=16 << 10, fmaps_max=512):
        super().__init__()
        if resolution not in _RESOLUTIONS_ALLOWED:
            raise ValueError(f'Invalid resolution: `{resolution}`!\n'
                             f'Resolutions allowed: {_RESOLUTIONS_ALLOWED}.')
        if label_size or fused_scale or final_tanh or not use_wscale:
            raise NotImplementedError('dge_b200 PGGAN: label_size=0, fused_scale=False, final_tanh=False, use_wscale')
        self.init_res = _INIT_RES
        self.init_res_log2 = int(np.log2(self.init_res))
        self.resolution = resolution
        self.final_res_log2 = int(np.log2(self.resolution))
        self.z_space_dim = z_space_dim
        self.image_channels = image_channels
        self.final_tanh = final_tanh
        self.label_size = label_size
        self.fused_scale = fused_scale
        self.use_wscale = use_wscale
        self.fmaps_base = fmaps_base
        self.fmaps_max = fmaps_max
        self.num_layers = (self.final_res_log2 - self.init_res_log2 + 1) * 2
        self.register_buffer('lod', torch.zeros(()))
        self.pth_to_tf_var_mapping = {'lod': 'lod'}
        for res_log2 in range(self.init_res_log2, self.final_res_log2 + 1):
            res = 2 ** res_log2
            block_idx = res_log2 - self.init_res_log2
            if res == self.init_res:
                self.add_module(f'layer{2 * block_idx}',
                                ConvBlock(in_channels=z_space_dim + label_size, out_channels=self.get_nf(res),
                                          kernel_size=self.init_res, padding=self.init_res - 1, use_wscale=use_wscale))
                tf = 'Dense'
            else:
                self.add_module(f'layer{2 * block_idx}',
                                ConvBlock(in_channels=self.get_nf(res // 2), out_channels=self.get_nf(res),
                                          upsample=True, fused_scale=fused_scale, use_wscale=use_wscale))
                tf = 'Conv0_up' if fused_scale else 'Conv0'
            self.pth_to_tf_var_mapping[f'layer{2 * block_idx}.weight'] = f'{res}x{res}/{tf}/weight'
            self.pth_to_tf_var_mapping[f'layer{2 * block_idx}.bias'] = f'{res}x{res}/{tf}/bias'
            self.add_module(f'layer{2 * block_idx + 1}',
                            ConvBlock(in_channels=self.get_nf(res), out_channels=self.get_nf(res),
                                      use_wscale=use_wscale))
            tf = 'Conv' if res == self.init_res else 'Conv1'
            self.pth_to_tf_var_mapping[f'layer{2 * block_idx + 1}.weight'] = f'{res}x{res}/{tf}/weight'
            self.pth_to_tf_var_mapping[f'layer{2 * block_idx + 1}.bias'] = f'{res}x{res}/{tf}/bias'
            self.add_module(f'output{block_idx}',
                            ConvBlock(in_channels=self.get_nf(res), out_channels=image_channels, kernel_size=1,
                                      padding=0, use_wscale=use_wscale, wscale_gain=1.0, activation_type='linear'))
            self.pth_to_tf_var_mapping[f'output{block_idx}.weight'] = f'ToRGB_lod{self.final_res_log2 - res_log2}/weight'
            self.pth_to_tf_var_mapping[f'output{block_idx}.bias'] = f'ToRGB_lod{self.final_res_log2 - res_log2}/bias'
        self.upsample = UpsamplingLayer()
        self.final_activate = nn.Identity()

    def get_nf(self, res):
        return min(self.fmaps_base // res, self.fmaps_max)

    def forward(self, z, label=None, lod=None, **_unused_kwargs):
        if z.ndim != 2 or z.shape[1] != self.z_space_dim:
            raise ValueError(f'Input latent code should be with shape [batch_size, latent_dim], where '
                             f'`latent_dim` equals to {self.z_space_dim}!\nBut `{z.shape}` is received!')
        _guard('PGGANGenerator', z)
        z = self.layer0.pixel_norm(z)                                   # :160
        lod = self.lod.cpu().tolist() if lod is None else lod           # :175
        if lod + self.init_res_log2 > self.final_res_log2:
            raise ValueError(f'Maximum level-of-detail (lod) is {self.final_res_log2 - self.init_res_log2}, '
                             f'but `{lod}` is received!')
        x = z                                                           # [N, C] stands for [N, C, 1, 1]
        image = None
        for res_log2 in range(self.init_res_log2, self.final_res_log2 + 1):
            current_lod = self.final_res_log2 - res_log2
            if lod < current_lod + 1:
                block_idx = res_log2 - self.init_res_log2
                x = getattr(self, f'layer{2 * block_idx}').run(x)
                x = getattr(self, f'layer{2 * block_idx + 1}').run(x)
            if current_lod - 1 < lod <= current_lod:
                image = getattr(self, f'output{block_idx}').run(x)
            elif current_lod < lod < current_lod + 1:
                alpha = np.ceil(lod) - lod
                image = ops.axpby(getattr(self, f'output{block_idx}').run(x), self.upsample(image), alpha, 1 - alpha)
            elif lod >= current_lod + 1:
                image = self.upsample(image)
            print(torch.Size((x.n, x.c, x.h, x.w)))                     # :196 (reference side effect)
        image = self.final_activate(image)
        return {'z': z, 'label': label, 'image': image}


class PixelNormLayer(nn.Module):
    def __init__(self, epsilon=1e-8):
        super().__init__()
        self.eps = epsilon

    def forward(self, x):
        _guard('PixelNormLayer', x)
        if x.ndim == 2:
            return ops.pixel_norm(x.float(), self.eps)
        f = ops.nchw_to_f32b(x.float())
        return ops.pixelnorm_to_act(f, 1, self.eps).to_nchw()


class UpsamplingLayer(nn.Module):
    def __init__(self, scale_factor=2):
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, x):
        if self.scale_factor <= 1:
            return x
        if self.scale_factor != 2:
            raise NotImplementedError('nearest upsampling x2 only')
        _guard('UpsamplingLayer', x)
        return ops.upsample_nearest_nchw(x.float())


class ConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, add_bias=True, upsample=False,
                 fused_scale=False, use_wscale=True, wscale_gain=_WSCALE_GAIN, activation_type='lrelu'):
        super().__init__()
        if upsample and fused_scale:
            raise NotImplementedError('fused_scale=True (conv2d_transpose) is not used by the inversion scripts')
        self.pixel_norm = PixelNormLayer()
        self.upsample = UpsamplingLayer() if upsample else nn.Identity()
        self.up = 2 if upsample else 1
        self.use_conv2d_transpose = False
        self.in_c, self.out_c, self.ksize = in_channels, out_channels, kernel_size
        self.stride, self.padding = stride, padding
        fan_in = kernel_size * kernel_size * in_channels
        wscale = wscale_gain / np.sqrt(fan_in)
        if use_wscale:
            self.weight = nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
            self.wscale = wscale
        else:
            self.weight = nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size) * wscale)
            self.wscale = 1.0
        self.bias = nn.Parameter(torch.zeros(out_channels)) if add_bias else None
        if activation_type == 'linear':
            self.slope = 1.0
        elif activation_type == 'lrelu':
            self.slope = 0.2
        else:
            raise NotImplementedError(f'Not implemented activation function: `{activation_type}`!')
        self.planes = DEFAULT_PLANES
        self._key, self._prep = None, None

    def _prepared(self):
        key = (ops.weight_key(self.weight), self.planes)
        if key != self._key:
            d = {}
            w = self.weight.detach()
            if self.ksize == 3 and self.padding == 1:
                d['wpk'] = ops.pack_conv_weight(w, scale=self.wscale, planes=self.planes)
            elif self.ksize == 1:
                d['w1x1'] = (w.view(self.out_c, self.in_c) * self.wscale).contiguous()
            else:
                # k x k conv with padding k-1 on a 1x1 input == dense: out[o,y,x] = sum_i z[i] W[o,i,k-1-y,k-1-x]  (:319-336)
                k = self.ksize
                d['wdense'] = w.flip(2, 3).permute(0, 2, 3, 1).reshape(self.out_c * k * k, self.in_c).contiguous()
                d['bdense'] = None if self.bias is None else self.bias.detach().repeat_interleave(k * k).contiguous()
            self._prep, self._key = d, key
        return self._prep

    def run(self, x):
        """x: F32B feature map, or [N, C] for the 1x1 'image' fed to layer0.  Returns F32B (NCHW image for k=1)."""
        p = self._prepared()
        if isinstance(x, torch.Tensor):            # layer0: latent vector
            assert x.ndim == 2 and self.padding == self.ksize - 1
            z = ops.pixel_norm(x.float(), self.pixel_norm.eps)
            y = ops.dense(z, p['wdense'], p['bdense'], wscale=self.wscale, slope=self.slope)
            k = self.ksize
            return ops.nchw_to_f32b(y.view(x.shape[0], self.out_c, k, k))
        if self.ksize == 1:
            return ops.pixelnorm_to_rgb(x, p['w1x1'], self.bias, self.pixel_norm.eps)
        xa = ops.pixelnorm_to_act(x, self.up, self.pixel_norm.eps, self.planes)
        return ops.conv(xa, p['wpk'], self.out_c, ops.CONV_3X3, bias=None if self.bias is None else self.bias.detach(),
                        slope=self.slope, out_f32b=True)['f32b']

    def forward(self, x):
        """Reference signature: NCHW in -> NCHW out."""
        _guard('ConvBlock', x, self.weight)
        if x.shape[2] == 1 and x.shape[3] == 1 and self.padding == self.ksize - 1:
            out = self.run(x.reshape(x.shape[0], -1))
        else:
            out = self.run(ops.nchw_to_f32b(x.float()))
        return out if isinstance(out, torch.Tensor) else out.to_nchw()
