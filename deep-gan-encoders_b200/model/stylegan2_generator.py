"""B200-native StyleGAN2 generator -- drop-in for the reference's `model/stylegan2_generator.py`.

Same import path, class names, constructor signatures, `forward` signatures / result dicts and
`state_dict()` keys+shapes, so `E_align_s2.py` / checkpoints written for
the reference load unchanged.  The arithmetic is NOT PyTorch: every dense layer, modulated
convolution, FIR up-sampling and ToRGB accumulation runs in the hand-written sm_100a kernels behind
the C ABI (include/dge_b200.h); PyTorch only owns device memory and the stream.

Inference (`torch.no_grad()`, or no input that requires grad) runs the fused forward-only kernels.  Training
(E_align_s2.py:160: `generator.synthesis(w2)['image']` with w2 produced by the encoder under autograd):
`SynthesisModule.forward` records ONE autograd node for the whole pass (dge_b200/train_g.py): the same forward kernel
chain, and a backward made of dge_b200 kernels only (fused point-wise backward + tcgen05 data-gradient convs, the x2
layers through the stride-2 conv over a space-to-depth map).  The generator is frozen in every training script (only
`E.parameters()` reach LREQAdam, E_align_s2.py:97), so its parameters enter as constants and its never-read `.grad`
is not accumulated (a one-time warning says so when they have requires_grad=True).  `_forward_autograd` is the same
computation as separate torch nodes (cross-check of the fused path; `FUSED_TRAIN = False` selects it).  The mapping
network stays forward-only.  No CPU fallback: CPU tensors raise DgeError.

Reference behaviours kept on purpose: result-dict keys, train-mode `w_avg` EMA + style mixing with
the same RNG calls (:177-191), `randomize_noise=True` drawing `torch.randn(N,1,res,res)` on the CPU
per layer in layer order (:912-913), ValueErrors on bad shapes (:99-105, :247-251, :493-498).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import warnings

from dge_b200 import autograd as tc
from dge_b200 import ops
from dge_b200 import train_g

FUSED_TRAIN = True     # False: the unfused torch-node graph (`_forward_autograd`)
_warned_frozen = False

__all__ = ['StyleGAN2Generator']

_RESOLUTIONS_ALLOWED = [8, 16, 32, 64, 128, 256, 512, 1024]
_INIT_RES = 4
_ARCHITECTURES_ALLOWED = ['resnet', 'skip', 'origin']
_WSCALE_GAIN = 1.0
_SQRT2 = float(np.sqrt(2.0))

# planes=2: bf16x3 split precision (fp32-equivalent, the parity mode); planes=1: plain bf16 fast mode.
DEFAULT_PLANES = 2


def _check_no_grad(*tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            'this stand-alone dge_b200 StyleGAN2 module is forward-only: wrap the call in torch.no_grad().  The '
            'training path is `generator.synthesis(wp)` (gradient w.r.t. wp, as E_align_s2.py:160 uses it)')


def _require_cuda(t, what):
    if not t.is_cuda:
        raise ops.DgeError(f'{what}: dge_b200 runs on a B200 only (got a {t.device} tensor); there is no CPU fallback')


class _Cache:
    """Derived tensors (packed weights etc.) keyed by the versions of their source parameters."""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, tensors, build):
        key = ops.weight_key(*tensors)
        if key != self.key:
            self.val = build()
            self.key = key
        return self.val


class StyleGAN2Generator(nn.Module):
    """Reference: model/stylegan2_generator.py:35-196."""

    def __init__(self, resolution, z_space_dim=512, w_space_dim=512, label_size=0, mapping_layers=8,
                 mapping_fmaps=512, mapping_lr_mul=0.01, repeat_w=True, image_channels=3, final_tanh=False,
                 const_input=True, architecture='skip', fused_modulate=True, demodulate=True, use_wscale=True,
                 fmaps_base=32 << 10, fmaps_max=512):
        super().__init__()
        if resolution not in _RESOLUTIONS_ALLOWED:
            raise ValueError(f'Invalid resolution: `{resolution}`!\n'
                             f'Resolutions allowed: {_RESOLUTIONS_ALLOWED}.')
        if architecture not in _ARCHITECTURES_ALLOWED:
            raise ValueError(f'Invalid architecture: `{architecture}`!\n'
                             f'Architectures allowed: {_ARCHITECTURES_ALLOWED}.')
        if architecture != 'skip':
            # the reference's 'resnet' branch cannot run (undefined self.scale_factor, :729)
            raise NotImplementedError("only architecture='skip' (the reference default) is implemented")
        if label_size != 0 or not const_input or not repeat_w or not use_wscale or final_tanh:
            raise NotImplementedError('dge_b200 implements the configuration the inversion scripts use: '
                                      'label_size=0, const_input, repeat_w, use_wscale, final_tanh=False')
        self.init_res = _INIT_RES
        self.resolution = resolution
        self.z_space_dim = z_space_dim
        self.w_space_dim = w_space_dim
        self.label_size = label_size
        self.mapping_layers = mapping_layers
        self.mapping_fmaps = mapping_fmaps
        self.mapping_lr_mul = mapping_lr_mul
        self.repeat_w = repeat_w
        self.image_channels = image_channels
        self.final_tanh = final_tanh
        self.const_input = const_input
        self.architecture = architecture
        self.fused_modulate = fused_modulate
        self.demodulate = demodulate
        self.use_wscale = use_wscale
        self.fmaps_base = fmaps_base
        self.fmaps_max = fmaps_max
        self.num_layers = int(np.log2(self.resolution // self.init_res * 2)) * 2
        self.mapping_space_dim = self.w_space_dim

        self.mapping = MappingModule(input_space_dim=z_space_dim, hidden_space_dim=mapping_fmaps,
                                     final_space_dim=self.mapping_space_dim, label_size=label_size,
                                     num_layers=mapping_layers, use_wscale=use_wscale, lr_mul=mapping_lr_mul)
        self.truncation = TruncationModule(w_space_dim=w_space_dim, num_layers=self.num_layers, repeat_w=repeat_w)
        self.synthesis = SynthesisModule(resolution=resolution, init_resolution=self.init_res,
                                         w_space_dim=w_space_dim, image_channels=image_channels,
                                         final_tanh=final_tanh, const_input=const_input, architecture=architecture,
                                         fused_modulate=fused_modulate, demodulate=demodulate,
                                         use_wscale=use_wscale, fmaps_base=fmaps_base, fmaps_max=fmaps_max)
        self.pth_to_tf_var_mapping = {}
        for sub in ('mapping', 'truncation', 'synthesis'):
            for key, val in getattr(self, sub).pth_to_tf_var_mapping.items():
                self.pth_to_tf_var_mapping[f'{sub}.{key}'] = val

    def forward(self, z, label=None, w_moving_decay=0.995, style_mixing_prob=0.9, trunc_psi=None,
                trunc_layers=None, randomize_noise=False, **_unused_kwargs):
        mapping_results = self.mapping(z, label)
        w = mapping_results['w']
        if self.training and w_moving_decay < 1:           # :177-182
            batch_w_avg = w.mean(dim=0)
            self.truncation.w_avg.copy_(self.truncation.w_avg * w_moving_decay + batch_w_avg * (1 - w_moving_decay))
        if self.training and style_mixing_prob > 0:        # :184-191
            new_z = torch.randn_like(z)
            new_w = self.mapping(new_z, label)['w']
            if np.random.uniform() < style_mixing_prob:
                mixing_cutoff = np.random.randint(1, self.num_layers)
                w = self.truncation(w)
                new_w = self.truncation(new_w)
                w[:, :mixing_cutoff] = new_w[:, :mixing_cutoff]
        wp = self.truncation(w, trunc_psi, trunc_layers)
        synthesis_results = self.synthesis(wp, randomize_noise)
        return {**mapping_results, **synthesis_results}


class MappingModule(nn.Module):
    """Reference: model/stylegan2_generator.py:199-278.  pixel-norm + `num_layers` dense kernels."""

    def __init__(self, input_space_dim=512, hidden_space_dim=512, final_space_dim=512, label_size=0, num_layers=8,
                 normalize_input=True, use_wscale=True, lr_mul=0.01):
        super().__init__()
        self.input_space_dim = input_space_dim
        self.hidden_space_dim = hidden_space_dim
        self.final_space_dim = final_space_dim
        self.label_size = label_size
        self.num_layers = num_layers
        self.normalize_input = normalize_input
        self.use_wscale = use_wscale
        self.lr_mul = lr_mul
        self.norm = PixelNormLayer() if normalize_input else nn.Identity()
        self.pth_to_tf_var_mapping = {}
        for i in range(num_layers):
            in_channels = input_space_dim if i == 0 else hidden_space_dim
            out_channels = final_space_dim if i == (num_layers - 1) else hidden_space_dim
            self.add_module(f'dense{i}', DenseBlock(in_channels=in_channels, out_channels=out_channels,
                                                    use_wscale=use_wscale, lr_mul=lr_mul))
            self.pth_to_tf_var_mapping[f'dense{i}.weight'] = f'Dense{i}/weight'
            self.pth_to_tf_var_mapping[f'dense{i}.bias'] = f'Dense{i}/bias'

    def forward(self, z, label=None):
        if z.ndim != 2 or z.shape[1] != self.input_space_dim:
            raise ValueError(f'Input latent code should be with shape [batch_size, input_dim], where '
                             f'`input_dim` equals to {self.input_space_dim}!\nBut `{z.shape}` is received!')
        z = self.norm(z)
        w = z
        for i in range(self.num_layers):
            w = getattr(self, f'dense{i}')(w)
        return {'z': z, 'label': label, 'w': w}


class TruncationModule(nn.Module):
    """Reference: model/stylegan2_generator.py:281-333 (a [N,L,D] lerp towards w_avg; host-side plumbing)."""

    def __init__(self, w_space_dim, num_layers, repeat_w=True):
        super().__init__()
        self.num_layers = num_layers
        self.w_space_dim = w_space_dim
        self.repeat_w = repeat_w
        self.register_buffer('w_avg', torch.zeros(w_space_dim))
        self.pth_to_tf_var_mapping = {'w_avg': 'dlatent_avg'}

    def forward(self, w, trunc_psi=None, trunc_layers=None):
        if w.ndim == 2:
            if self.repeat_w and w.shape[1] == self.w_space_dim:
                wp = w.view(-1, 1, self.w_space_dim).repeat(1, self.num_layers, 1)
            else:
                assert w.shape[1] == self.w_space_dim * self.num_layers
                wp = w.view(-1, self.num_layers, self.w_space_dim)
        else:
            wp = w
        assert wp.ndim == 3
        assert wp.shape[1:] == (self.num_layers, self.w_space_dim)
        trunc_psi = 1.0 if trunc_psi is None else trunc_psi
        trunc_layers = 0 if trunc_layers is None else trunc_layers
        if trunc_psi < 1.0 and trunc_layers > 0:
            layer_idx = np.arange(self.num_layers).reshape(1, -1, 1)
            coefs = np.ones_like(layer_idx, dtype=np.float32)
            coefs[layer_idx < trunc_layers] *= trunc_psi
            coefs = torch.from_numpy(coefs).to(wp)
            w_avg = self.w_avg.view(1, -1, self.w_space_dim)
            wp = w_avg + (wp - w_avg) * coefs
        return wp


class PixelNormLayer(nn.Module):
    """Reference: model/stylegan2_generator.py:543-553."""

    def __init__(self, dim=1, epsilon=1e-8):
        super().__init__()
        self.dim = dim
        self.eps = epsilon

    def forward(self, x):
        _require_cuda(x, 'PixelNormLayer')
        _check_no_grad(x)
        if x.ndim != 2 or self.dim != 1:
            raise NotImplementedError('PixelNormLayer kernel handles [N, D] latents (the mapping-network use)')
        return ops.pixel_norm(x.float(), self.eps)


class UpsamplingLayer(nn.Module):
    """Holds the FIR `kernel` buffer of the reference layer (state_dict compatibility,
    model/stylegan2_generator.py:556-615); the filtering itself is fused into dge kernels."""

    def __init__(self, scale_factor=2, kernel=(1, 3, 3, 1), extra_padding=0, kernel_gain=None):
        super().__init__()
        assert scale_factor >= 1
        self.scale_factor = scale_factor
        k = np.array(kernel, dtype=np.float32)
        k = np.outer(k, k)
        k = k / np.sum(k)
        k = k * (scale_factor ** 2 if kernel_gain is None else kernel_gain ** 2)
        self.register_buffer('kernel', torch.from_numpy(k[np.newaxis, np.newaxis]))
        if tuple(kernel) != (1, 3, 3, 1):
            raise NotImplementedError('dge_b200 fuses the (1,3,3,1) FIR only')

    def forward(self, x):
        """Skip-branch x2 upsample of an NCHW image (scale_factor == 2)."""
        if self.scale_factor != 2:
            raise NotImplementedError('standalone filtering is fused into ModulateConvBlock')
        _require_cuda(x, 'UpsamplingLayer')
        _check_no_grad(x)
        n, c, h, w = x.shape
        return ops.rgb_init(x.float().contiguous(), None, n, c, 2 * h, 2 * w, x.device)


class InputBlock(nn.Module):
    """Reference: model/stylegan2_generator.py:618-632."""

    def __init__(self, init_resolution, channels):
        super().__init__()
        self.const = nn.Parameter(torch.randn(1, channels, init_resolution, init_resolution))

    def forward(self, w):
        return self.const.repeat(w.shape[0], 1, 1, 1)


class DenseBlock(nn.Module):
    """Reference: model/stylegan2_generator.py:925-996.  One `dge_dense` launch."""

    def __init__(self, in_channels, out_channels, add_bias=True, additional_bias=0, use_wscale=True,
                 wscale_gain=_WSCALE_GAIN, lr_mul=1.0, activation_type='lrelu'):
        super().__init__()
        wscale = wscale_gain / np.sqrt(in_channels)
        if use_wscale:
            self.weight = nn.Parameter(torch.randn(out_channels, in_channels) / lr_mul)
            self.wscale = wscale * lr_mul
        else:
            self.weight = nn.Parameter(torch.randn(out_channels, in_channels) * wscale / lr_mul)
            self.wscale = lr_mul
        self.bias = nn.Parameter(torch.zeros(out_channels)) if add_bias else None
        self.bscale = lr_mul
        self.additional_bias = additional_bias
        if activation_type == 'linear':
            self.slope, self.activate_scale = 1.0, 1.0
        elif activation_type == 'lrelu':
            self.slope, self.activate_scale = 0.2, _SQRT2
        else:
            raise NotImplementedError(f'Not implemented activation function: `{activation_type}`!')

    def forward(self, x):
        _require_cuda(x, 'DenseBlock')
        _check_no_grad(x, self.weight)
        if x.ndim != 2:
            x = x.view(x.shape[0], -1)
        return ops.dense(x.float(), self.weight, self.bias, wscale=self.wscale, bscale=self.bscale,
                         add_bias=self.additional_bias, slope=self.slope, gain=self.activate_scale)


class ModulateConvBlock(nn.Module):
    """Reference: model/stylegan2_generator.py:742-922.

    y = d[n,o] * conv(x * s[n,i], W*wscale) (+FIR for the x2 layers) + noise*strength + bias -> lrelu*sqrt2,
    computed by `dge_conv_forward` (tcgen05) with the epilogue fused; see csrc/conv_mma.cu.
    """

    def __init__(self, in_channels, out_channels, resolution, w_space_dim, kernel_size=3, add_bias=True,
                 scale_factor=1, filtering_kernel=(1, 3, 3, 1), fused_modulate=True, demodulate=True,
                 use_wscale=True, wscale_gain=_WSCALE_GAIN, lr_mul=1.0, add_noise=True, activation_type='lrelu',
                 epsilon=1e-8):
        super().__init__()
        self.res = resolution
        self.in_c = in_channels
        self.out_c = out_channels
        self.ksize = kernel_size
        self.eps = epsilon
        if scale_factor > 1:
            if scale_factor != 2 or kernel_size != 3:
                raise NotImplementedError('dge_b200 implements the x2, 3x3 up-sampling layer')
            self.use_conv2d_transpose = True
            self.filter = UpsamplingLayer(scale_factor=1, kernel=filtering_kernel,
                                          extra_padding=scale_factor - kernel_size, kernel_gain=scale_factor)
        else:
            self.use_conv2d_transpose = False
            assert kernel_size % 2 == 1
        fan_in = kernel_size * kernel_size * in_channels
        wscale = wscale_gain / np.sqrt(fan_in)
        if use_wscale:
            self.weight = nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size) / lr_mul)
            self.wscale = wscale * lr_mul
        else:
            self.weight = nn.Parameter(
                torch.randn(out_channels, in_channels, kernel_size, kernel_size) * wscale / lr_mul)
            self.wscale = lr_mul
        self.style = DenseBlock(in_channels=w_space_dim, out_channels=in_channels, additional_bias=1.0,
                                use_wscale=use_wscale, activation_type='linear')
        self.fused_modulate = fused_modulate
        self.demodulate = demodulate
        self.bias = nn.Parameter(torch.zeros(out_channels)) if add_bias else None
        self.bscale = lr_mul
        if activation_type == 'linear':
            self.slope, self.activate_scale = 1.0, 1.0
        elif activation_type == 'lrelu':
            self.slope, self.activate_scale = 0.2, _SQRT2
        else:
            raise NotImplementedError(f'Not implemented activation function: `{activation_type}`!')
        self.add_noise = add_noise
        if self.add_noise:
            self.register_buffer('noise', torch.randn(1, 1, self.res, self.res))
            self.noise_strength = nn.Parameter(torch.zeros(()))
        self.planes = DEFAULT_PLANES
        self._cache = _Cache()

    # ---- derived, cached per parameter version ------------------------------------------------
    def _prepared(self):
        srcs = [self.weight] + ([self.bias] if self.bias is not None else []) + \
               ([self.noise_strength] if self.add_noise else [])

        def build():
            d = {}
            mma_ok = self.out_c % 16 == 0 and self.in_c % 16 == 0
            if mma_ok:
                d['wpk'] = ops.pack_conv_weight(self.weight, scale=self.wscale, flip=self.use_conv2d_transpose,
                                                planes=self.planes)
            if self.demodulate:
                d['w2'] = ops.weight_sqsum(self.weight, scale=self.wscale)
            d['bias'] = None if self.bias is None else (self.bias.detach() * self.bscale).contiguous()
            d['strength'] = float(self.noise_strength.detach().item()) if self.add_noise else 0.0
            return d

        return self._cache.get(srcs, build)

    def _noise(self, batch, randomize_noise, device):
        """-> (tensor or None, batched flag); reference :911-916 (randn on the CPU, then .to(x))."""
        if not self.add_noise:
            return None, False
        if randomize_noise:
            return torch.randn(batch, 1, self.res, self.res).to(device).contiguous(), True
        return self.noise, False

    def run(self, xa, style, randomize_noise=False, next_style=None, want_act=True, want_nchw=False, rgb=None,
            dm=None):
        """Chained form: `xa` is an ACT tensor already multiplied by this layer's style; `dm` = precomputed
        demodulation coefficients (dge_sg2_prep), else they are computed here."""
        p = self._prepared()
        n = xa.n
        if dm is None:
            dm = ops.demod(p['w2'], style, self.eps) if self.demodulate else None
        noise, batched = self._noise(n, randomize_noise, xa.t.device)
        if self.use_conv2d_transpose:
            raw = ops.conv(xa, p['wpk'], self.out_c, ops.CONV_UP3X3)['raw_up']
            return ops.up_fir_epilogue(raw, n, self.out_c, 2 * xa.h, 2 * xa.w, demod=dm, noise=noise,
                                       noise_batched=batched, noise_scalar=p['strength'], bias=p['bias'],
                                       slope=self.slope, gain=self.activate_scale, out_scale=next_style,
                                       planes=self.planes, out_act=want_act, out_nchw=want_nchw)
        kind = ops.CONV_3X3 if self.ksize == 3 else ops.CONV_1X1
        rgb_w, rgb_out = rgb if rgb is not None else (None, None)
        return ops.conv(xa, p['wpk'], self.out_c, kind, demod=dm, noise=noise, noise_batched=batched,
                        noise_scalar=p['strength'], bias=p['bias'], slope=self.slope, gain=self.activate_scale,
                        out_act=want_act, out_scale=next_style, out_nchw=want_nchw, rgb_w=rgb_w, rgb_out=rgb_out)

    def _forward_autograd(self, x, w_latent, randomize_noise=False):
        """Differentiable w.r.t. x and w_latent (parameters are constants): :855-922 in the shared-weight form
        y = d[n,o] * conv(x * s[n,i], W) of SURVEY Appendix E-1.  NCHW in, (NCHW out, style)."""
        n = x.shape[0]
        st = self.style
        sb = None if st.bias is None else st.bias.detach() * st.bscale
        style = F.linear(w_latent, st.weight.detach() * st.wscale, sb) + st.additional_bias       # :862, 990-996
        wgt = self.weight.detach() * self.wscale                                                   # :858
        xs = x * style.view(n, self.in_c, 1, 1)                                                    # :877
        if self.use_conv2d_transpose:
            y = tc.lib_conv_transpose2d(xs, wgt.flip(2, 3).transpose(0, 1), stride=2)                   # :879-895
            y = _fir4(y, self.filter.kernel, (1, 1, 1, 1))                                         # :896
        elif self.ksize == 3 and self.in_c % 16 == 0 and self.out_c % 16 == 0:
            y = tc.conv2d(xs, wgt, self.planes)                                                    # :897-904 (tcgen05)
        else:
            y = tc.lib_conv2d(xs, wgt, padding=self.ksize // 2)                                         # ToRGB (3 outputs)
        if self.demodulate:
            d = torch.rsqrt((wgt.square().sum(dim=(2, 3))[None] * style.square()[:, None]).sum(dim=2) + self.eps)
            y = y * d.view(n, self.out_c, 1, 1)                                                    # :867-870, 908-909
        if self.add_noise:
            noise, _ = self._noise(n, randomize_noise, x.device)
            y = y + noise * self.noise_strength.detach()                                           # :911-916
        if self.bias is not None:
            y = y + (self.bias.detach() * self.bscale).view(1, -1, 1, 1)                           # :918-920
        if self.slope != 1.0:
            y = F.leaky_relu(y, self.slope) * self.activate_scale                                  # :921
        return y, style

    def rgb_weights(self, style):
        """ToRGB layers (k=1, no demod): per-sample [N][3][Cin] weights with style and wscale folded in."""
        return ops.rgb_weights(self.weight, style, self.wscale)

    def forward(self, x, w, randomize_noise=False):
        """Stand-alone reference signature: NCHW in, (NCHW out, style)."""
        _require_cuda(x, 'ModulateConvBlock')
        _check_no_grad(x, w, self.weight)
        if x.shape[1] != self.in_c:
            raise ValueError(f'expected {self.in_c} input channels, got {x.shape[1]}')
        style = self.style(w)
        if self.ksize == 1 and not self.demodulate and not self.add_noise and self.slope == 1.0:
            p = self._prepared()
            return ops.to_rgb_nchw(x.float(), self.rgb_weights(style), p['bias']), style
        if self.out_c % 16 or self.in_c % 16:
            raise NotImplementedError('tcgen05 conv needs channel counts that are multiples of 16')
        xa = ops.nchw_to_act(x.float(), scale=style, planes=self.planes)
        out = self.run(xa, style, randomize_noise, want_act=False, want_nchw=True)
        return out['nchw'], style


def _fir4(x, kernel, pad):
    """Depth-wise 4x4 FIR of an NCHW tensor after zero padding `pad` (left, right, top, bottom) -- UpsamplingLayer
    :592-615 (training path only; the inference kernels fuse it)."""
    n, c, h, w = x.shape
    y = tc.lib_conv2d(F.pad(x.reshape(n * c, 1, h, w), pad), kernel.to(x))
    return y.view(n, c, y.shape[2], y.shape[3])


def _fir_up2(img, kernel):
    """Skip-branch x2 upsample (:556-615, scale_factor=2): zero insertion, pad (2,1,2,1), 4x4 FIR."""
    n, c, h, w = img.shape
    z = img.new_zeros(n, c, 2 * h, 2 * w)
    z[:, :, ::2, ::2] = img
    return _fir4(z, kernel, (2, 1, 2, 1))


class SynthesisModule(nn.Module):
    """Reference: model/stylegan2_generator.py:336-539 (architecture 'skip')."""

    def __init__(self, resolution=1024, init_resolution=4, w_space_dim=512, image_channels=3, final_tanh=False,
                 const_input=True, architecture='skip', fused_modulate=True, demodulate=True, use_wscale=True,
                 fmaps_base=32 << 10, fmaps_max=512):
        super().__init__()
        self.init_res = init_resolution
        self.init_res_log2 = int(np.log2(self.init_res))
        self.resolution = resolution
        self.final_res_log2 = int(np.log2(self.resolution))
        self.w_space_dim = w_space_dim
        self.image_channels = image_channels
        self.final_tanh = final_tanh
        self.const_input = const_input
        self.architecture = architecture
        self.fused_modulate = fused_modulate
        self.demodulate = demodulate
        self.use_wscale = use_wscale
        self.fmaps_base = fmaps_base
        self.fmaps_max = fmaps_max
        self.num_layers = (self.final_res_log2 - self.init_res_log2 + 1) * 2
        self.pth_to_tf_var_mapping = {}
        common = dict(w_space_dim=w_space_dim, fused_modulate=fused_modulate, use_wscale=use_wscale)
        for res_log2 in range(self.init_res_log2, self.final_res_log2 + 1):
            res = 2 ** res_log2
            block_idx = res_log2 - self.init_res_log2
            if res == self.init_res:
                self.add_module('early_layer', InputBlock(init_resolution=self.init_res, channels=self.get_nf(res)))
                self.pth_to_tf_var_mapping['early_layer.const'] = f'{res}x{res}/Const/const'
            else:
                name = f'layer{2 * block_idx - 1}'
                self.add_module(name, ModulateConvBlock(in_channels=self.get_nf(res // 2),
                                                        out_channels=self.get_nf(res), resolution=res,
                                                        scale_factor=2, demodulate=demodulate, **common))
                self._tf_names(name, f'{res}x{res}/Conv0_up', f'noise{2 * block_idx - 1}')
            name = f'layer{2 * block_idx}'
            self.add_module(name, ModulateConvBlock(in_channels=self.get_nf(res), out_channels=self.get_nf(res),
                                                    resolution=res, demodulate=demodulate, **common))
            self._tf_names(name, f'{res}x{res}/' + ('Conv' if res == self.init_res else 'Conv1'),
                           f'noise{2 * block_idx}')
            name = f'output{block_idx}'
            self.add_module(name, ModulateConvBlock(in_channels=self.get_nf(res), out_channels=image_channels,
                                                    resolution=res, kernel_size=1, demodulate=False,
                                                    add_noise=False, activation_type='linear', **common))
            self._tf_names(name, f'{res}x{res}/ToRGB', None)
        self.upsample = UpsamplingLayer()
        self.final_activate = nn.Identity()

    def _tf_names(self, name, tf, noise):
        m = self.pth_to_tf_var_mapping
        m[f'{name}.weight'] = f'{tf}/weight'
        m[f'{name}.bias'] = f'{tf}/bias'
        m[f'{name}.style.weight'] = f'{tf}/mod_weight'
        m[f'{name}.style.bias'] = f'{tf}/mod_bias'
        if noise is not None:
            m[f'{name}.noise_strength'] = f'{tf}/noise_strength'
            m[f'{name}.noise'] = noise

    def get_nf(self, res):
        return min(self.fmaps_base // res, self.fmaps_max)

    def _prep(self, wp32, layers, outputs, with_handle=False):
        """-> (styles, demods, rgb_styles, rgb_weights), all views into one arena filled by `dge_sg2_prep`.
        The item table (device array of `dge_sg2_prep_item`) is rebuilt only when a parameter or the batch size
        changes.  with_handle: a fifth element {'cache': the cached item table, 'arena': this pass's arena} for
        `dge_sg2_prep_bwd` (training, dge_b200/train_g.py)."""
        import ctypes
        from dge_b200._lib import Sg2PrepItem
        n, dev = wp32.shape[0], wp32.device
        srcs = []
        for m in list(layers) + list(outputs):
            srcs += [m.weight, m.style.weight] + ([m.style.bias] if m.style.bias is not None else [])
        if not hasattr(self, '_prep_cache'):
            self._prep_cache = _Cache()

        def build():
            items = (Sg2PrepItem * (len(layers) + len(outputs)))()
            keep, views, off = [], [], 0
            for j, m in enumerate(list(layers) + list(outputs)):
                it = items[j]
                is_rgb = j >= len(layers)
                st = m.style
                it.st_w = st.weight.data_ptr()
                it.st_b = st.bias.data_ptr() if st.bias is not None else None
                it.wp_index = (2 * (j - len(layers)) + 1) if is_rgb else j
                it.cin, it.cout, it.nch = m.in_c, m.out_c, (m.out_c if is_rgb else 0)
                it.st_wscale, it.st_bscale, it.st_add_bias = st.wscale, st.bscale, st.additional_bias
                it.eps = m.eps
                it.style_off = off
                v = {'style': (off, (n, m.in_c))}
                off += n * m.in_c
                it.demod_off = it.rgbw_off = -1
                if is_rgb:
                    w = m.weight.detach().contiguous().view(m.out_c, -1)
                    keep.append(w)
                    it.rgb_w, it.rgb_scale = w.data_ptr(), m.wscale
                    it.rgbw_off = off
                    v['rgbw'] = (off, (n, m.out_c, m.in_c))
                    off += n * m.out_c * m.in_c
                elif m.demodulate:
                    w2 = m._prepared()['w2']
                    keep.append(w2)
                    it.w2 = w2.data_ptr()
                    it.demod_off = off
                    v['demod'] = (off, (n, m.out_c))
                    off += n * m.out_c
                views.append(v)
            raw = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).to(dev)
            return {'items': raw, 'n_items': len(items), 'views': views, 'floats': off, 'keep': keep, 'n': n}

        c = self._prep_cache.get(srcs, build)
        if c['n'] != n:
            self._prep_cache.key = None
            c = self._prep_cache.get(srcs, build)
        arena = torch.empty(c['floats'], dtype=torch.float32, device=dev)
        ops.sg2_prep(c['items'], c['n_items'], wp32, arena, n, self.num_layers, self.w_space_dim)

        def view(spec):
            o, shape = spec
            numel = 1
            for d in shape:
                numel *= d
            return arena[o:o + numel].view(*shape)

        nl = len(layers)
        styles = [view(c['views'][i]['style']) for i in range(nl)]
        demods = [view(c['views'][i]['demod']) if 'demod' in c['views'][i] else None for i in range(nl)]
        rgb_styles = [view(c['views'][nl + k]['style']) for k in range(len(outputs))]
        rgb_ws = [view(c['views'][nl + k]['rgbw']) for k in range(len(outputs))]
        if with_handle:
            return styles, demods, rgb_styles, rgb_ws, {'cache': c, 'arena': arena}
        return styles, demods, rgb_styles, rgb_ws

    def _forward_autograd(self, wp, randomize_noise=False):
        """Training path: the same result dict, recorded for backward w.r.t. `wp` (see the module docstring)."""
        n, nl = wp.shape[0], self.num_layers
        wp32 = wp.float()
        results = {'wp': wp}
        x = self.early_layer.const.detach().float().expand(n, -1, -1, -1)                          # :630-632
        image = None
        for i in range(nl - 1):
            x, style = getattr(self, f'layer{i}')._forward_autograd(x, wp32[:, i], randomize_noise)
            results[f'style{i:02d}'] = style
            if i % 2 == 0:                                                                         # :511-522
                rgb, style = getattr(self, f'output{i // 2}')._forward_autograd(x, wp32[:, i + 1])
                results[f'output_style{i // 2}'] = style
                image = rgb if image is None else rgb + _fir_up2(image, self.upsample.kernel)
        results['image'] = self.final_activate(image)
        return results

    def forward(self, wp, randomize_noise=False):
        if wp.ndim != 3 or wp.shape[1:] != (self.num_layers, self.w_space_dim):
            raise ValueError(f'Input tensor should be with shape [batch_size, num_layers, w_space_dim], where '
                             f'`num_layers` equals to {self.num_layers}, and `w_space_dim` equals to '
                             f'{self.w_space_dim}!\nBut `{wp.shape}` is received!')
        _require_cuda(wp, 'SynthesisModule')
        if torch.is_grad_enabled() and wp.requires_grad:
            global _warned_frozen
            if not _warned_frozen and any(p.requires_grad for p in self.parameters()):
                _warned_frozen = True
                warnings.warn('dge_b200 SynthesisModule: the generator is treated as FROZEN in the training path -- '
                              'gradients flow to `wp` only; its parameters receive no .grad (the inversion scripts '
                              'never read it).', stacklevel=2)
            if FUSED_TRAIN:
                return train_g.synthesis_forward(self, wp, randomize_noise)
            return self._forward_autograd(wp, randomize_noise)
        n, dev = wp.shape[0], wp.device
        wp32 = wp.float()
        results = {'wp': wp}
        nl = self.num_layers
        layers = [getattr(self, f'layer{i}') for i in range(nl - 1)]
        outputs = [getattr(self, f'output{k}') for k in range(nl // 2)]
        # every per-layer scalar -- style affines (layer i reads wp[:, i], ToRGB k reads wp[:, 2k+1], :511-517), the
        # demodulation coefficients and the ToRGB weights -- comes out of ONE launch into one arena
        styles, demods, rgb_styles, rgb_ws = self._prep(wp32.contiguous(), layers, outputs)
        for i, st in enumerate(styles):
            results[f'style{i:02d}'] = st
        for k, st in enumerate(rgb_styles):
            results[f'output_style{k}'] = st

        planes = layers[0].planes
        # InputBlock: const.repeat(N) (:630-632), pre-multiplied by layer0's style
        xa = ops.nchw_to_act(self.early_layer.const.detach().float(), scale=styles[0], planes=planes, batch=n)
        image = None
        for i in range(nl - 1):
            layer = layers[i]
            nxt = styles[i + 1] if i + 1 < nl - 1 else None
            if i % 2 == 1:
                xa = layer.run(xa, styles[i], randomize_noise, next_style=nxt, dm=demods[i])['act']
                continue
            k = i // 2
            out_l = outputs[k]
            res = layer.res
            # image_k = bias + up2(image_{k-1}); the conv epilogue adds the ToRGB contribution (:515-522)
            image = ops.rgb_init(image, out_l._prepared()['bias'], n, self.image_channels, res, res, dev)
            r = layer.run(xa, styles[i], randomize_noise, next_style=nxt, want_act=nxt is not None,
                          rgb=(rgb_ws[k], image), dm=demods[i])
            xa = r.get('act')
        results['image'] = self.final_activate(image)
        return results
