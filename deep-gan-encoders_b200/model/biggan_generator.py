"""B200-native BigGAN-deep generator -- drop-in for the reference's `model/biggan_generator.py`
(snconv2d/snlinear :49-56, SelfAttn :58-97, BigGANBatchNorm :100-150, GenBlock :153-203, Generator :205-256,
BigGAN :258-304).

Same class names / constructor arguments / forward signatures and the same state_dict keys: the spectral-norm layers
are real `nn.utils.spectral_norm` wrappers (weight_orig / weight_u / weight_v / bias), so HF `biggan-deep-*` weights
load unchanged and train-mode power iteration behaves as upstream.  The normalised weights are read
through the wrapper's own hook (weight prep, cached per parameter version in eval mode); every activation op is a
dge_b200 kernel:
  conditional BN (+ReLU, + nearest x2) -> bf16 hi/lo operand : `dge_cbn_coeffs` + `dge_affine_act`
  1x1 / 3x3 convs, bias, residual add incl. the channel-drop + nearest-up skip : `dge_conv_forward`
  self-attention: q.k^T and attn.v are 1x1 tcgen05 convs whose "weights" are the per-sample key / value maps
  (the ACT layout of a [C, HW] map IS the packed-weight layout), softmax over keys = `dge_channel_softmax_to_act`.
`truncation` follows the reference's Python arithmetic (`math.modf(truncation / step)`), tensors included.

Training (E_align_s2.py:162, mtype 4: `generator(w2, conditions, truncation)` with w2 from the encoder under autograd):
when `z` requires grad, `BigGAN.forward` records a differentiable graph w.r.t. `z`; the generator is frozen (its effective
spectral-norm weights enter as constants).  `FUSED_TRAIN`: one autograd node per GenBlock and one for the RGB tail
(`dge_b200/train_big.py`): forward = the inference kernel chain, backward = 4 x [data-gradient conv -> `dge_affine_relu_bwd`];
the [N, C] conditional-BN coefficients and the attention block stay torch graphs.  Channel counts that are not multiples
of 16 (toy widths) and `FUSED_TRAIN = False` take the graph of separate torch nodes with tcgen05 convs (`_forward_autograd`).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from dge_b200 import autograd as tc
from dge_b200 import ops
from model.utils.biggan_config import BigGANConfig  # noqa: F401

DEFAULT_PLANES = 2
FUSED_TRAIN = True      # one fused autograd node per GenBlock (dge_b200/train_big.py); False: separate torch nodes


def snconv2d(eps=1e-12, **kwargs):
    return nn.utils.spectral_norm(nn.Conv2d(**kwargs), eps=eps)


def snlinear(eps=1e-12, **kwargs):
    return nn.utils.spectral_norm(nn.Linear(**kwargs), eps=eps)


def sn_embedding(eps=1e-12, **kwargs):
    return nn.utils.spectral_norm(nn.Embedding(**kwargs), eps=eps)


def _guard(name, *tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ops.DgeError(f'{name}: dge_b200 runs on a B200 only (got a {t.device} tensor); no CPU fallback')
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(f'{name}: dge_b200 kernels are forward-only in this build; use torch.no_grad()')


def _conv_autograd(x, w, b, planes):
    """Differentiable stride-1 'same' conv of the training path: tensor cores when the channel counts allow it."""
    k = w.shape[-1]
    if w.shape[0] % 16 == 0 and w.shape[1] % 16 == 0:
        y = tc.conv2d(x, w, planes)
    else:
        y = tc.lib_conv2d(x, w, padding=k // 2)
    return y if b is None else y + b.view(1, -1, 1, 1)


def sn_weight(module):
    """Effective weight of a spectral-norm wrapped layer: runs the wrapper's forward-pre-hook exactly as a forward would
    (one power iteration in train mode, none in eval), i.e. W_orig / (u^T W v)."""
    for hook in module._forward_pre_hooks.values():
        hook(module, None)
    return module.weight


class _Prep:
    """Cache of derived tensors keyed by the versions of the source tensors (eval mode only)."""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, module, tensors, build):
        key = ops.weight_key(*tensors)
        if module.training or key != self.key:
            self.val, self.key = build(), key
        return self.val


def _sn_sources(layer):
    return [layer.weight_orig, layer.weight_u, layer.weight_v]


class SelfAttn(nn.Module):
    def __init__(self, in_channels, eps=1e-12):
        super().__init__()
        self.in_channels = in_channels
        self.snconv1x1_theta = snconv2d(in_channels=in_channels, out_channels=in_channels // 8, kernel_size=1,
                                        bias=False, eps=eps)
        self.snconv1x1_phi = snconv2d(in_channels=in_channels, out_channels=in_channels // 8, kernel_size=1,
                                      bias=False, eps=eps)
        self.snconv1x1_g = snconv2d(in_channels=in_channels, out_channels=in_channels // 2, kernel_size=1,
                                    bias=False, eps=eps)
        self.snconv1x1_o_conv = snconv2d(in_channels=in_channels // 2, out_channels=in_channels, kernel_size=1,
                                         bias=False, eps=eps)
        self.maxpool = nn.MaxPool2d(2, stride=2, padding=0)
        self.softmax = nn.Softmax(dim=-1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.planes = DEFAULT_PLANES
        self._prep = _Prep()

    def _prepared(self):
        layers = [self.snconv1x1_theta, self.snconv1x1_phi, self.snconv1x1_g, self.snconv1x1_o_conv]
        srcs = [t for l in layers for t in _sn_sources(l)] + [self.gamma]

        def build():
            d = {n: ops.pack_conv_weight(sn_weight(l).detach(), planes=self.planes)
                 for n, l in zip(('theta', 'phi', 'g', 'o'), layers)}
            d['gamma'] = float(self.gamma.detach().item())
            return d

        return self._prep.get(self, srcs, build)

    def run(self, x):
        """x: F32B [N, C, H, W] -> F32B."""
        p = self._prepared()
        n, ch, h, w = x.n, x.c, x.h, x.w
        dev = x.t.device
        xa = ops.f32b_to_act(x, self.planes)
        theta = ops.conv(xa, p['theta'], ch // 8, ops.CONV_1X1, out_act=True)['act']                  # queries (:78-79)
        phi = ops.maxpool2(ops.conv(xa, p['phi'], ch // 8, ops.CONV_1X1, out_f32b=True)['f32b'])      # keys (:81-83)
        g = ops.maxpool2(ops.conv(xa, p['g'], ch // 2, ops.CONV_1X1, out_f32b=True)['f32b'])          # values (:88-92)
        nk = (h // 2) * (w // 2)
        # keys as per-sample "weights" [Cout = nk][Cin = ch/8]: the ACT layout of phi is exactly WPK with one tap
        phi_w = ops.f32b_to_act(phi, self.planes).t.view(n, 1, ch // 64, self.planes, nk, 8)
        g_nchw = g.to_nchw().view(n, ch // 2, nk, 1, 1)                                               # [Cout][Cin = nk]
        attn_g = ops.F32B(n, ch // 2, h, w, dev)
        for b in range(n):
            q_b = ops.Act.__new__(ops.Act)
            q_b.n, q_b.c, q_b.h, q_b.w, q_b.planes, q_b.t = 1, ch // 8, h, w, self.planes, theta.t[b:b + 1]
            logits = ops.conv(q_b, phi_w[b], nk, ops.CONV_1X1, out_f32b=True)['f32b']                 # bmm (:85)
            attn = ops.channel_softmax_to_act(logits, self.planes)                                    # softmax (:86)
            gw = ops.pack_conv_weight(g_nchw[b], planes=self.planes)
            ops.conv(attn, gw, ch // 2, ops.CONV_1X1, out_f32b_into=attn_g.t[b:b + 1])                # bmm (:94)
        ga = ops.f32b_to_act(attn_g, self.planes)
        return ops.conv(ga, p['o'], ch, ops.CONV_1X1, gain=p['gamma'], blend_src=x, blend_a=1.0, blend_b=1.0,
                        out_f32b=True)['f32b']                                                        # :95-97

    def _forward_autograd(self, x):
        """:75-97 on NCHW tensors, recorded for backward (frozen weights)."""
        n, ch, h, w = x.shape
        wt = [sn_weight(l).detach() for l in (self.snconv1x1_theta, self.snconv1x1_phi, self.snconv1x1_g,
                                              self.snconv1x1_o_conv)]
        theta = _conv_autograd(x, wt[0], None, self.planes).view(n, ch // 8, h * w)
        phi = F.max_pool2d(_conv_autograd(x, wt[1], None, self.planes), 2, 2).view(n, ch // 8, h * w // 4)
        attn = torch.softmax(torch.bmm(theta.transpose(1, 2), phi), dim=-1)
        g = F.max_pool2d(_conv_autograd(x, wt[2], None, self.planes), 2, 2).view(n, ch // 2, h * w // 4)
        attn_g = torch.bmm(g, attn.transpose(1, 2)).view(n, ch // 2, h, w)
        return x + self.gamma.detach() * _conv_autograd(attn_g, wt[3], None, self.planes)

    def forward(self, x):
        _guard('SelfAttn', x)
        return self.run(ops.nchw_to_f32b(x.float())).to_nchw()


class BigGANBatchNorm(nn.Module):
    def __init__(self, num_features, condition_vector_dim=None, n_stats=51, eps=1e-4, conditional=True):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.conditional = conditional
        self.register_buffer('running_means', torch.zeros(n_stats, num_features))
        self.register_buffer('running_vars', torch.ones(n_stats, num_features))
        self.step_size = 1.0 / (n_stats - 1)
        if conditional:
            assert condition_vector_dim is not None
            self.scale = snlinear(in_features=condition_vector_dim, out_features=num_features, bias=False, eps=eps)
            self.offset = snlinear(in_features=condition_vector_dim, out_features=num_features, bias=False, eps=eps)
        else:
            self.weight = torch.nn.Parameter(torch.Tensor(num_features))
            self.bias = torch.nn.Parameter(torch.Tensor(num_features))

    def stats(self, truncation):
        """Pre-computed statistics for this truncation, with the reference's interpolation quirk (:129-136)."""
        coef, start_idx = math.modf(truncation / self.step_size)
        start_idx = int(start_idx)
        if coef != 0.0:
            mean = self.running_means[start_idx] * coef + self.running_means[start_idx + 1] * (1 - coef)
            var = self.running_vars[start_idx] * coef + self.running_vars[start_idx + 1] * (1 - coef)
        else:
            mean, var = self.running_means[start_idx], self.running_vars[start_idx]
        return mean, var

    def coeffs(self, truncation, condition_vector, n):
        """Per-(n, c) affine (A, B) with y = x*A + B  ==  (x - mean)/sqrt(var+eps)*weight + bias  (:138-150)."""
        mean, var = self.stats(truncation)
        if self.conditional:
            # effective spectral-norm weights: cached per parameter version in eval mode (the wrapper's hook costs three
            # tiny library launches per layer per forward; train mode still runs it -- it updates u)
            if not hasattr(self, '_prep'):
                self._prep = _Prep()
            ws, wo = self._prep.get(self, _sn_sources(self.scale) + _sn_sources(self.offset),
                                    lambda: (sn_weight(self.scale).detach().clone(), sn_weight(self.offset).detach().clone()))
            s = ops.dense(condition_vector.float(), ws, None)
            o = ops.dense(condition_vector.float(), wo, None)
            return ops.cbn_coeffs(mean, var, self.eps, n, scale=s, offset=o)
        return ops.cbn_coeffs(mean, var, self.eps, n, weight=self.weight, bias=self.bias)

    def coeffs_autograd(self, truncation, condition_vector, n, frozen=True):
        """`coeffs` as differentiable torch ops -> (A, B) fp32 [n, c]: the path the gradient takes from a fused block node
        back to the condition vector (frozen generator) or to the scale / offset layers (the BigGAN encoder)."""
        mean, var = self.stats(truncation)
        if self.conditional:
            if frozen and not self.training:
                if not hasattr(self, '_prep'):
                    self._prep = _Prep()
                ws, wo = self._prep.get(self, _sn_sources(self.scale) + _sn_sources(self.offset),
                                        lambda: (sn_weight(self.scale).detach().clone(),
                                                 sn_weight(self.offset).detach().clone()))
            else:
                keep = (lambda t: t.detach()) if frozen else (lambda t: t)
                ws, wo = keep(sn_weight(self.scale)), keep(sn_weight(self.offset))
            cv = condition_vector.float()
            weight = 1 + F.linear(cv, ws)
            bias = F.linear(cv, wo)
        else:
            keep = (lambda t: t.detach()) if frozen else (lambda t: t)
            weight = keep(self.weight).unsqueeze(0).expand(n, -1)
            bias = keep(self.bias).unsqueeze(0).expand(n, -1)
        a = weight / torch.sqrt(var + self.eps)
        b = bias - mean * a
        return a.contiguous(), b.contiguous()

    def _forward_autograd(self, x, truncation, condition_vector=None, frozen=True):
        """:138-150 recorded for backward.  `frozen`: the layer's own parameters are constants (generator) or
        trainable (the BigGAN encoder re-uses this class, E_BIG.py:33-82)."""
        mean, var = self.stats(truncation)
        keep = (lambda t: t.detach()) if frozen else (lambda t: t)
        if self.conditional:
            weight = 1 + F.linear(condition_vector, keep(sn_weight(self.scale)))[:, :, None, None]
            bias = F.linear(condition_vector, keep(sn_weight(self.offset)))[:, :, None, None]
            return (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + self.eps) * weight + bias
        return F.batch_norm(x, mean, var, keep(self.weight), keep(self.bias), False, 0.0, self.eps)

    def forward(self, x, truncation, condition_vector=None):
        _guard('BigGANBatchNorm', x, condition_vector)
        a, b = self.coeffs(truncation, condition_vector, x.shape[0])
        _, f = ops.affine_act(ops.nchw_to_f32b(x.float()), a, b, relu=False, out_act=False, out_f32b=True)
        return f.to_nchw()


class GenBlock(nn.Module):
    def __init__(self, in_size, out_size, condition_vector_dim, reduction_factor=4, up_sample=False, n_stats=51,
                 eps=1e-12):
        super().__init__()
        self.up_sample = up_sample
        self.drop_channels = (in_size != out_size)
        middle_size = in_size // reduction_factor
        self.in_size, self.out_size, self.middle_size = in_size, out_size, middle_size
        self.bn_0 = BigGANBatchNorm(in_size, condition_vector_dim, n_stats=n_stats, eps=eps, conditional=True)
        self.conv_0 = snconv2d(in_channels=in_size, out_channels=middle_size, kernel_size=1, eps=eps)
        self.bn_1 = BigGANBatchNorm(middle_size, condition_vector_dim, n_stats=n_stats, eps=eps, conditional=True)
        self.conv_1 = snconv2d(in_channels=middle_size, out_channels=middle_size, kernel_size=3, padding=1, eps=eps)
        self.bn_2 = BigGANBatchNorm(middle_size, condition_vector_dim, n_stats=n_stats, eps=eps, conditional=True)
        self.conv_2 = snconv2d(in_channels=middle_size, out_channels=middle_size, kernel_size=3, padding=1, eps=eps)
        self.bn_3 = BigGANBatchNorm(middle_size, condition_vector_dim, n_stats=n_stats, eps=eps, conditional=True)
        self.conv_3 = snconv2d(in_channels=middle_size, out_channels=out_size, kernel_size=1, eps=eps)
        self.relu = nn.ReLU()
        self.planes = DEFAULT_PLANES
        self._prep = _Prep()

    def _prepared(self):
        convs = [self.conv_0, self.conv_1, self.conv_2, self.conv_3]
        srcs = [t for c in convs for t in _sn_sources(c) + [c.bias]]

        def build():
            eff = [sn_weight(c).detach() for c in convs]
            return {'w': [ops.pack_conv_weight(w, planes=self.planes) for w in eff], 'eff': eff,
                    'b': [c.bias.detach() for c in convs]}

        return self._prep.get(self, srcs, build)

    def _prepared_k(self, kns):
        """`_prepared` through the kernel namespace `kns` (dge_b200.ops, or the CPU tests' emulation: uncached)."""
        if kns is ops:
            return self._prepared()
        convs = [self.conv_0, self.conv_1, self.conv_2, self.conv_3]
        eff = [sn_weight(c).detach() for c in convs]
        return {'w': [kns.pack_conv_weight(w, planes=self.planes) for w in eff], 'eff': eff,
                'b': [c.bias.detach() for c in convs]}

    def run(self, x, cond_vector, truncation):
        """x: F32B -> F32B."""
        p = self._prepared()
        n = x.n
        a, b = self.bn_0.coeffs(truncation, cond_vector, n)
        t, _ = ops.affine_act(x, a, b, relu=True, planes=self.planes)                                          # :178-179
        t = ops.conv(t, p['w'][0], self.middle_size, ops.CONV_1X1, bias=p['b'][0], out_f32b=True)['f32b']      # :180
        a, b = self.bn_1.coeffs(truncation, cond_vector, n)
        t, _ = ops.affine_act(t, a, b, relu=True, up=2 if self.up_sample else 1, planes=self.planes)           # :182-185
        t = ops.conv(t, p['w'][1], self.middle_size, ops.CONV_3X3, bias=p['b'][1], out_f32b=True)['f32b']      # :186
        a, b = self.bn_2.coeffs(truncation, cond_vector, n)
        t, _ = ops.affine_act(t, a, b, relu=True, planes=self.planes)
        t = ops.conv(t, p['w'][2], self.middle_size, ops.CONV_3X3, bias=p['b'][2], out_f32b=True)['f32b']      # :188-190
        a, b = self.bn_3.coeffs(truncation, cond_vector, n)
        t, _ = ops.affine_act(t, a, b, relu=True, planes=self.planes)
        # conv_3 + skip: x0[:, :out] (channel drop) nearest-upsampled if needed, added in the epilogue (:192-203)
        return ops.conv(t, p['w'][3], self.out_size, ops.CONV_1X1, bias=p['b'][3], preact_add=x,
                        preact_up=2 if self.up_sample else 1, out_f32b=True)['f32b']

    def _forward_autograd(self, x, cond_vector, truncation):
        """:175-203 recorded for backward (frozen weights)."""
        def conv(layer, v):
            return _conv_autograd(v, sn_weight(layer).detach(), layer.bias.detach(), self.planes)

        x0 = x
        x = conv(self.conv_0, F.relu(self.bn_0._forward_autograd(x, truncation, cond_vector)))
        x = F.relu(self.bn_1._forward_autograd(x, truncation, cond_vector))
        if self.up_sample:
            x = F.interpolate(x, scale_factor=2, mode='nearest')
        x = conv(self.conv_1, x)
        x = conv(self.conv_2, F.relu(self.bn_2._forward_autograd(x, truncation, cond_vector)))
        x = conv(self.conv_3, F.relu(self.bn_3._forward_autograd(x, truncation, cond_vector)))
        if self.drop_channels:
            x0 = x0[:, :x0.shape[1] // 2]                                                 # :195-197
        if self.up_sample:
            x0 = F.interpolate(x0, scale_factor=2, mode='nearest')
        return x + x0

    def forward(self, x, cond_vector, truncation):
        _guard('GenBlock', x, cond_vector)
        return self.run(ops.nchw_to_f32b(x.float()), cond_vector, truncation).to_nchw()


class Generator(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        ch = config.channel_width
        condition_vector_dim = config.z_dim * 2
        self.gen_z = snlinear(in_features=condition_vector_dim, out_features=4 * 4 * 16 * ch, eps=config.eps)
        layers = []
        for i, layer in enumerate(config.layers):
            if i == config.attention_layer_position:
                layers.append(SelfAttn(ch * layer[1], eps=config.eps))
            layers.append(GenBlock(ch * layer[1], ch * layer[2], condition_vector_dim, up_sample=layer[0],
                                   n_stats=config.n_stats, eps=config.eps))
        self.layers = nn.ModuleList(layers)
        self.bn = BigGANBatchNorm(ch, n_stats=config.n_stats, eps=config.eps, conditional=False)
        self.relu = nn.ReLU()
        self.conv_to_rgb = snconv2d(in_channels=ch, out_channels=ch, kernel_size=3, padding=1, eps=config.eps)
        self.tanh = nn.Tanh()
        self.planes = DEFAULT_PLANES
        self._prep = _Prep()

    def _frozen(self, layer):
        """Effective (spectral-norm) weight of a frozen layer as a constant; cached in eval mode."""
        if self.training:
            return sn_weight(layer).detach()
        cache = self.__dict__.setdefault('_frozen_cache', {})
        key = ops.weight_key(*_sn_sources(layer))
        hit = cache.get(id(layer))
        if hit is None or hit[0] != key:
            hit = (key, sn_weight(layer).detach().clone())
            cache[id(layer)] = hit
        return hit[1]

    def _rgb_prepared_k(self, kns):
        """Forward / data-gradient operands of conv_to_rgb: only channels [:3] of its output are kept (:253), so the
        first 16 output channels are packed.  Cached in eval mode for the real kernel namespace."""
        def build():
            w = sn_weight(self.conv_to_rgb).detach()[:16].contiguous()
            return {'w': kns.pack_conv_weight(w, planes=self.planes), 'wd': kns.pack_conv_weight_dgrad(w, planes=self.planes),
                    'b': self.conv_to_rgb.bias.detach()[:16].contiguous()}

        if kns is not ops:
            return build()
        if not hasattr(self, '_prep_rgb'):
            self._prep_rgb = _Prep()
        return self._prep_rgb.get(self, _sn_sources(self.conv_to_rgb) + [self.conv_to_rgb.bias], build)

    def _forward_autograd(self, cond_vector, truncation):
        """:232-256 recorded for backward w.r.t. the condition vector (frozen weights)."""
        # (the fused nodes keep every map in the 16-channel-blocked layouts; narrower toy widths take the torch graph below)
        if FUSED_TRAIN and all(s % 16 == 0 for l in self.layers if isinstance(l, GenBlock)
                               for s in (l.in_size, l.out_size, l.middle_size)):
            from dge_b200 import train_big
            return train_big.generator_forward(self, cond_vector, truncation)
        ch = self.config.channel_width
        x = F.linear(cond_vector, sn_weight(self.gen_z).detach(), self.gen_z.bias.detach())
        x = x.view(-1, 4, 4, 16 * ch).permute(0, 3, 1, 2).contiguous()
        for layer in self.layers:
            x = layer._forward_autograd(x, cond_vector, truncation) if isinstance(layer, GenBlock) \
                else layer._forward_autograd(x)
        x = F.relu(self.bn._forward_autograd(x, truncation))
        x = _conv_autograd(x, sn_weight(self.conv_to_rgb).detach(), self.conv_to_rgb.bias.detach(), self.planes)
        return torch.tanh(x[:, :3])

    def forward(self, cond_vector, truncation):
        if torch.is_grad_enabled() and cond_vector.requires_grad:
            if not cond_vector.is_cuda:
                raise ops.DgeError('BigGAN.Generator: dge_b200 runs on a B200 only; there is no CPU fallback')
            return self._forward_autograd(cond_vector.float(), truncation)
        _guard('BigGAN.Generator', cond_vector)
        ch = self.config.channel_width
        n = cond_vector.shape[0]
        cv = cond_vector.float().contiguous()
        if not hasattr(self, '_prep_z'):
            self._prep_z = _Prep()
        wz = self._prep_z.get(self, _sn_sources(self.gen_z), lambda: sn_weight(self.gen_z).detach().clone())
        z = ops.dense(cv, wz, self.gen_z.bias)                                             # :233
        z = z.view(-1, 4, 4, 16 * ch).permute(0, 3, 1, 2).contiguous()                    # TF NHWC -> NCHW (:237-239)
        x = ops.nchw_to_f32b(z)
        for layer in self.layers:
            x = layer.run(x, cv, truncation) if isinstance(layer, GenBlock) else layer.run(x)
        a, b = self.bn.coeffs(truncation, None, n)                                         # :247
        t, _ = ops.affine_act(x, a, b, relu=True, planes=self.planes)                      # :249

        def build():
            # only channels [:3] of the ch-wide RGB conv are kept (:253); pack the first 16 output channels
            w = sn_weight(self.conv_to_rgb).detach()
            return {'w': ops.pack_conv_weight(w[:16].contiguous(), planes=self.planes),
                    'b': self.conv_to_rgb.bias.detach()[:16].contiguous()}

        p = self._prep.get(self, _sn_sources(self.conv_to_rgb) + [self.conv_to_rgb.bias], build)
        rgb16 = ops.conv(t, p['w'], 16, ops.CONV_3X3, bias=p['b'], out_nchw=True)['nchw']  # :251
        return ops.tanh_slice_nchw(rgb16, 3)                                               # :253-255


class BigGAN(nn.Module):
    """BigGAN Generator (reference :258-304; `from_pretrained` needs the network and is not provided)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = nn.Linear(config.num_classes, config.z_dim, bias=False)
        self.generator = Generator(config)

    def forward(self, z, class_label, truncation):
        assert 0 < truncation <= 1
        if torch.is_grad_enabled() and z.requires_grad:
            if not z.is_cuda:
                raise ops.DgeError('BigGAN: dge_b200 runs on a B200 only; there is no CPU fallback')
            embed = F.linear(class_label.float(), self.embeddings.weight.detach())          # :299
            cond_vector = torch.cat((z.float(), embed), dim=1)                              # :301
            return self.generator._forward_autograd(cond_vector, truncation), cond_vector
        _guard('BigGAN', z, class_label)
        embed = ops.dense(class_label.float(), self.embeddings.weight, None)               # :299
        cond_vector = torch.cat((z.float(), embed), dim=1)                                  # :301
        z = self.generator(cond_vector, truncation)
        return z, cond_vector
