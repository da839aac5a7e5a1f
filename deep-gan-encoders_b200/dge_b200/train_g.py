"""Fused training path of the frozen StyleGAN2 synthesis network (model/stylegan2_generator.py:492-539 under
`loss.backward()`: E_align_s2.py:160 feeds the encoder's `w2` through `generator.synthesis` and back-propagates the
image loss into the encoder, so the generator needs d image / d wp and nothing else -- its parameters are constants).

ONE autograd node for the whole pass.  Forward = the inference kernel chain (style / demodulation / ToRGB tables in one
launch, tcgen05 modulated convs with the fused epilogue, x2 layers as transposed conv + TMA-staged FIR); what it keeps
for the backward is only what it produced anyway -- the ACT operand of every layer (y_{l-1} * s_l, bf16 hi + lo) -- plus
the last layer's output.  Backward, per layer from the top (csrc/train_bwd.cu):

    sg2_layer_bwd    lrelu' * gain * demod on (dxs * s_next + ToRGB^T d_image) -> operand of the data-gradient conv, and in
                     the same pass the three reductions the style gradient needs (sum dxs*y, sum d_image*y, sum d_pre*pre)
    plain layer      dge_conv_forward with the transposed + flipped weights (tcgen05)
    x2 layer         up_fir_bwd_s2d (FIR transpose, written space-to-depth) + dge_conv_forward(DOWN4X4S2): the stride-2
                     3x3 conv that inverts the transposed conv, as 9 of 16 taps over the space-to-depth map
    rgb_up_bwd       transpose of the skip image's x2 up-sampling

and finally ONE launch (sg2_prep_bwd, the transpose of the forward's sg2_prep) that turns the per-layer sums into d wp
(style affine :872-877, demodulation :867-870, ToRGB modulation :462-474): every layer's sums land in one arena whose
offsets the kernel's item table knows.  `K` is the kernel namespace (see train_e.py).
"""
import torch
from torch.autograd.function import once_differentiable

from . import graphs
from . import ops

K = ops


def _consts(layer, planes):
    """Per-layer constants of a frozen layer, cached like ModulateConvBlock._prepared but built through `K`:
    packed forward weights, demodulation Gram diagonal W2[o][i], scaled bias, noise strength."""
    srcs = [layer.weight] + ([layer.bias] if layer.bias is not None else []) + \
           ([layer.noise_strength] if layer.add_noise else [])
    key = (K.weight_key(*srcs), planes, K is ops)
    hit = layer.__dict__.get('_train_consts')
    if hit is not None and hit[0] == key:
        return hit[1]
    w = layer.weight.detach()
    d = {}
    if layer.out_c % 16 == 0 and layer.in_c % 16 == 0:
        d['wpk'] = K.pack_conv_weight(w, scale=layer.wscale, flip=layer.use_conv2d_transpose, planes=planes)
    if layer.demodulate:
        d['w2'] = ((w * layer.wscale) ** 2).sum(dim=(2, 3))
    d['bias'] = None if layer.bias is None else (layer.bias.detach() * layer.bscale).contiguous()
    d['strength'] = float(layer.noise_strength.detach().item()) if layer.add_noise else 0.0
    layer.__dict__['_train_consts'] = (key, d)
    return d


def _dgrad_operands(layer, planes):
    """Cached data-gradient operand of a (frozen) layer: plain 3x3 -> transposed + flipped WPK; x2 layer -> the 16-tap
    WPK of DGE_CONV_DOWN4X4S2 holding Wr[i][o][ky][kx] = W[o][i][2-ky][2-kx] in taps (ky+1, kx+1), zeros elsewhere."""
    key = (K.weight_key(layer.weight), planes, K is ops)
    hit = layer.__dict__.get('_dgrad_cache')
    if hit is not None and hit[0] == key:
        return hit[1]
    w = layer.weight.detach()
    if layer.use_conv2d_transpose:
        w4 = w.new_zeros((layer.in_c, layer.out_c, 4, 4))
        w4[:, :, 1:, 1:] = (w * layer.wscale).flip(2, 3).transpose(0, 1)
        op = K.pack_conv_weight(w4.contiguous(), planes=planes)
    else:
        op = K.pack_conv_weight_dgrad(w, scale=layer.wscale, planes=planes)
    layer.__dict__['_dgrad_cache'] = (key, op)
    return op


def _run_forward(S, wp32, randomize_noise):
    """The forward kernel chain -> (outputs: (image, *styles, *rgb_styles), saved: what the backward reads)."""
    n, dev = wp32.shape[0], wp32.device
    nl = S.num_layers
    layers = [getattr(S, f'layer{i}') for i in range(nl - 1)]
    outputs = [getattr(S, f'output{k}') for k in range(nl // 2)]
    styles, demods, rgb_styles, rgb_ws, prep = K.sg2_prep_all(S, wp32.contiguous(), layers, outputs)
    planes = layers[0].planes
    const = S.early_layer.const.detach().float()
    xa = K.nchw_to_act(const, scale=styles[0], planes=planes, batch=n)                        # :630-632
    xas, noises = [], []
    image = None
    y_last = None
    for i, layer in enumerate(layers):
        p = _consts(layer, planes)
        nxt = styles[i + 1] if i + 1 < nl - 1 else None
        noise, batched = layer._noise(n, randomize_noise, dev)
        xas.append(xa)
        noises.append((noise, batched))
        if layer.use_conv2d_transpose:                                                        # :879-896
            raw = K.conv(xa, p['wpk'], layer.out_c, K.CONV_UP3X3)['raw_up']
            xa = K.up_fir_epilogue(raw, n, layer.out_c, 2 * xa.h, 2 * xa.w, demod=demods[i], noise=noise,
                                   noise_batched=batched, noise_scalar=p['strength'], bias=p['bias'],
                                   slope=layer.slope, gain=layer.activate_scale, out_scale=nxt,
                                   planes=planes)['act']
            continue
        k = i // 2
        image = K.rgb_init(image, _consts(outputs[k], planes)['bias'], n, S.image_channels, layer.res, layer.res, dev)
        r = K.conv(xa, p['wpk'], layer.out_c, K.CONV_3X3, demod=demods[i], noise=noise, noise_batched=batched,
                   noise_scalar=p['strength'], bias=p['bias'], slope=layer.slope, gain=layer.activate_scale,
                   out_act=True, out_scale=nxt, rgb_w=rgb_ws[k], rgb_out=image)             # :897-921, 515-522
        xa = r['act']
        if nxt is None:
            y_last = xa
    saved = {'geom': (n, nl, planes), 'noises': noises, 'acts': xas + [y_last],
             'tabs': (styles, demods, rgb_styles, rgb_ws, prep)}
    return (image,) + tuple(styles) + tuple(rgb_styles), saved


def _run_backward(S, saved, d_image):
    """d image -> d wp [n, num_layers, w_space_dim] from what `_run_forward` saved."""
    n, nl, planes = saved['geom']
    styles, demods, rgb_styles, rgb_ws, prep = saved['tabs']
    layers = [getattr(S, f'layer{i}') for i in range(nl - 1)]
    outputs = [getattr(S, f'output{k}') for k in range(nl // 2)]
    acts = saved['acts']
    nk = nl // 2
    d_imgs = [None] * nk
    d_imgs[nk - 1] = d_image.contiguous().float()
    for k in range(nk - 1, 0, -1):                                                            # :519-522 transposed
        d_imgs[k - 1] = K.rgb_up_bwd(d_imgs[k])
    # one arena for every reduction the style gradient needs: [n, C0] (layer 0's style through x_0 = const * s_0),
    # then per layer the [n, out_c, 5] sums of sg2_layer_bwd: 0 = d style of the NEXT layer (through x * s),
    # 1..3 = d of the ToRGB weights, 4 = demod * d demod
    const_off, offs, total = 0, [], n * layers[0].in_c
    for layer in layers:
        offs.append(total)
        total += n * layer.out_c * 5
    arena = torch.empty(total, dtype=torch.float32, device=d_image.device)
    dxs = None
    for i in range(nl - 2, -1, -1):
        layer = layers[i]
        last = i == nl - 2
        p = _consts(layer, planes)
        k = i // 2 if i % 2 == 0 else None
        noise, batched = saved['noises'][i]
        up = layer.use_conv2d_transpose
        dconv, _ = K.sg2_layer_bwd(
            acts[i + 1], None if last else styles[i + 1], dxs, None if k is None else d_imgs[k],
            None if k is None else rgb_ws[k], noise, batched, p['strength'], p['bias'], demods[i],
            layer.activate_scale, layer.slope, out_kind='f32b' if up else 'act', planes=planes,
            sums=arena[offs[i]:offs[i] + n * layer.out_c * 5].view(n, layer.out_c, 5))
        if up:
            s2d = K.up_fir_bwd_s2d(dconv, planes)
            dxs = K.conv(s2d, _dgrad_operands(layer, planes), layer.in_c, K.CONV_DOWN4X4S2, out_f32b=True,
                         out_hw=(acts[i].h, acts[i].w))['f32b']
        else:
            dxs = K.conv(dconv, _dgrad_operands(layer, planes), layer.in_c, K.CONV_3X3, out_f32b=True)['f32b']
        del dconv
    const = S.early_layer.const.detach().float()
    torch.sum(dxs.to_nchw() * const, dim=(2, 3), out=arena[:n * layers[0].in_c].view(n, layers[0].in_c))
    return K.sg2_prep_bwd(S, prep, layers, outputs, arena, offs, const_off, n)


# CUDA-graph replay of the node: dge_b200/graphs.py (opt-in).  slot = 'synthesis'; key = wp shape, planes, weights epoch.
def _graph_key(S, wp32):
    srcs = S.__dict__.get('_train_graph_srcs')
    if srcs is None:
        srcs = S.__dict__['_train_graph_srcs'] = list(S.parameters()) + list(S.buffers())
    return (tuple(wp32.shape), wp32.device.index, S.layer0.planes, K.weight_key(*srcs))


class _SynthesisFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wp32, S, randomize_noise):
        ctx.S = S
        # CPU noise draws (randomize_noise, :912-913) cannot be captured
        outs, ctx.handle = graphs.forward(S, 'synthesis', _graph_key(S, wp32) if graphs.GRAPHS else None, (wp32,),
                                          lambda wp: _run_forward(S, wp, randomize_noise), 'train_g (synthesis)',
                                          enabled=K is ops and wp32.is_cuda and not randomize_noise)
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, d_image, *_unused):
        S = ctx.S
        d_wp = graphs.backward(ctx.handle, (d_image,), lambda saved, g: _run_backward(S, saved, g), 'train_g (synthesis)')
        return d_wp, None, None


def synthesis_forward(S, wp, randomize_noise=False):
    """`SynthesisModule.forward` (:492-539) recorded for backward w.r.t. `wp` -> the same results dict."""
    nl = S.num_layers
    out = _SynthesisFn.apply(wp.float(), S, randomize_noise)
    results = {'wp': wp}
    for i in range(nl - 1):
        results[f'style{i:02d}'] = out[1 + i]
    for k in range(nl // 2):
        results[f'output_style{k}'] = out[nl + k]
    results['image'] = S.final_activate(out[0])
    return results
