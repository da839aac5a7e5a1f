"""Fused LPIPS-VGG16 distance (the third-party `lpips.LPIPS(net='vgg')` the scripts put into `space_loss`,
training_utils.py:93): ONE autograd node for scaling layer -> 13 VGG convs -> 5 taps -> distance.

Both image batches run through the VGG stack as one batch of 2N.  Every conv is dge_conv_forward (tcgen05, split precision)
with bias and ReLU in its epilogue, writing the next conv's ACT operand directly (and, at the five taps, the F32B feature
map the distance reads); the 3-channel first conv runs on a zero-padded 16-channel operand; max-pool writes the next
operand.  Per tap the distance is one reduction kernel (channel unit-normalisation, squared difference, `lin` weights,
spatial mean).  Backward: the distance kernel's backward form, then per conv `relu_pool_bwd` (ReLU mask + arg-max routing
of the pooled gradient, straight into the ACT operand) + dge_conv_forward on the data-gradient weights -- only for the
half of the batch whose image requires grad.  The VGG / lin weights are frozen (`requires_grad=False`, as in the package).
PARITY UNPINNED (no reference vectors offline): checked against deep-gan-encoders_b200/lpips's unfused graph and
oracle/lpips.py.
"""
import torch
from torch.autograd.function import once_differentiable

from . import graphs
from . import ops

K = ops

# (slice k, conv index in torchvision vgg16.features); a 2x2 max-pool precedes every slice but the first
_VGG = [[0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28]]


def _operands(model, planes):
    """Packed forward / data-gradient weights of the frozen VGG stack (cached until a parameter changes)."""
    convs = [getattr(model.net, f'slice{k + 1}')[j] for k, idxs in enumerate(_VGG) for j in range(len(idxs))]
    key = (K.weight_key(*[c.weight for c in convs]), planes)
    hit = model.__dict__.get('_dge_vgg_ops')
    if hit is not None and hit[0] == key:
        return hit[1]
    fwd, bwd, bias = [], [], []
    for i, c in enumerate(convs):
        w = c.weight.detach()
        if i == 0:                      # 3 -> 64: zero-pad the input channels to the 16 the tensor-core path needs
            w = torch.cat((w, w.new_zeros(w.shape[0], 13, 3, 3)), dim=1).contiguous()
        fwd.append(K.pack_conv_weight(w, planes=planes))
        bwd.append(K.pack_conv_weight_dgrad(w, planes=planes))
        bias.append(c.bias.detach().contiguous())
    lin = [model.lins[k].model[1].weight.detach().reshape(-1).contiguous() for k in range(5)]
    sl = model.scaling_layer
    val = (fwd, bwd, bias, lin, (sl.shift.flatten().tolist(), sl.scale.flatten().tolist()))
    model.__dict__['_dge_vgg_ops'] = (key, val)
    return val


def _forward_chain(model, in0, in1, planes, keep_all=True):
    """The fused forward: -> (distance [N], saved activations (every conv's output; only the five taps if not keep_all))."""
    n = in0.shape[0]
    fwd, bwd, bias, lin, (shift, scale) = _operands(model, planes)
    x = torch.cat((in0.detach().float(), in1.detach().float()), dim=0).contiguous()
    act = K.lpips_input(x, shift, scale, planes)
    out = torch.zeros(n, dtype=torch.float32, device=x.device)
    saved = []                      # saved[i]: the activated output of conv i (Act, or F32B at a tap)
    i = 0
    for k, idxs in enumerate(_VGG):
        for j in range(len(idxs)):
            last = j == len(idxs) - 1
            cout = bias[i].numel()
            r = K.conv(act, fwd[i], cout, K.CONV_3X3, bias=bias[i], slope=0.0, out_act=not last, out_f32b=last)
            if last:
                f = r['f32b']
                saved.append(f)
                K.lpips_dist(f, lin[k], out)
                if k < 4:
                    act = K.maxpool_to_act(f, planes)
            else:
                act = r['act']
                if keep_all:
                    saved.append(act)
            i += 1
    return out, saved


def _backward_chain(model, planes, n, saved, go, want):
    """d distance [n] -> (d in0 | None, d in1 | None) from the saved conv outputs."""
    fwd, bwd, bias, lin, _ = _operands(model, planes)
    go = go.reshape(n).contiguous().float()
    sl = model.scaling_layer
    grads = [None, None]
    # tap gradients for both halves in one launch per tap
    tap_idx = [sum(len(v) for v in _VGG[:k + 1]) - 1 for k in range(5)]
    tap_g = [K.lpips_dist_bwd(saved[ti], lin[k], go, want[0], want[1]) for k, ti in enumerate(tap_idx)]
    for half in (0, 1):
        if not want[half]:
            continue

        def view(t):                 # this half of a saved [2n, ...] tensor (batch-major layouts: a contiguous slice)
            if isinstance(t, K.F32B):
                return K.F32B.wrap(t.t[half * n:(half + 1) * n], n, t.c, t.h, t.w)
            return K.Act.wrap(t.t[half * n:(half + 1) * n], n, t.c, t.h, t.w, t.planes)

        g_in = None                  # gradient w.r.t. the INPUT of the conv above (F32B), flowing down
        i = len(saved) - 1
        for k in range(4, -1, -1):
            for j in range(len(_VGG[k]) - 1, -1, -1):
                y = view(saved[i])
                if j == len(_VGG[k]) - 1:       # tap: its own distance gradient + what came back through the pool
                    d = K.relu_pool_bwd(y, g_same=tap_g[k][half], g_pool=g_in, planes=planes)
                else:
                    d = K.relu_pool_bwd(y, g_same=g_in, planes=planes)
                cin = 16 if i == 0 else (saved[i - 1].c)
                g_in = K.conv(d, bwd[i], cin, K.CONV_3X3, out_f32b=True)['f32b']
                i -= 1
        g = g_in.to_nchw()[:, :3] / sl.scale.to(g_in.t.device)                      # ScalingLayer: (x - shift) / scale
        grads[half] = g.contiguous()
    return grads[0], grads[1]


def _graph_key(model):
    srcs = model.__dict__.get('_dge_graph_srcs')
    if srcs is None:
        srcs = model.__dict__['_dge_graph_srcs'] = list(model.parameters()) + list(model.buffers())
    return K.weight_key(*srcs)


class _LpipsFn(torch.autograd.Function):
    """CUDA-graph replay (dge_b200/graphs.py, opt-in): one slot per (input shapes, which image needs a gradient) -- the three
    `space_loss` calls of an iteration pool to 256x256, 256x192 and 176x176 (training_utils.py:81-84), so each has its own."""

    @staticmethod
    def forward(ctx, in0, in1, model, planes):
        n = in0.shape[0]
        want = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        ctx.model, ctx.planes, ctx.n, ctx.want = model, planes, n, want
        slot = ('lpips', tuple(in0.shape), tuple(in1.shape), in0.dtype, in1.dtype, in0.device.index, planes, want)

        def fwd(a, b):
            out, saved = _forward_chain(model, a, b, planes)
            return (out,), saved

        (out,), ctx.handle = graphs.forward(model, slot, _graph_key(model) if graphs.GRAPHS else None, (in0, in1), fwd,
                                            'train_lpips', enabled=K is ops and in0.is_cuda)
        return out.view(n, 1, 1, 1)

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        model, planes, n, want = ctx.model, ctx.planes, ctx.n, ctx.want
        g0, g1 = graphs.backward(ctx.handle, (go,), lambda saved, g: _backward_chain(model, planes, n, saved, g, want),
                                 'train_lpips')
        return g0, g1, None, None


class _LpipsLinOnlyFn(torch.autograd.Function):
    """The distance of two images that carry NO gradient, as a function of the module's five `lin` weights only.  The
    package's result requires grad through those weights even then, and some of the scripts call backward() on exactly
    such losses (E_mis_align_cropping_s1.py:171-193, E_align_cropping_s1.py:185-206, embedding_img.py:96-112).  Forward = the
    fused chain; backward = d out[n] / d w_k[c] = mean_p (a_c/|a| - b_c/|b|)^2 from the five saved taps (torch reductions:
    nobody's optimiser holds these weights, the path only has to exist and be right)."""

    @staticmethod
    def forward(ctx, model, in0, in1, planes, *lin_w):
        n = in0.shape[0]
        need = tuple(ctx.needs_input_grad[4:])
        ctx.n, ctx.shapes, ctx.need = n, [w.shape for w in lin_w], need
        slot = ('lpips-lin', tuple(in0.shape), tuple(in1.shape), in0.dtype, in1.dtype, in0.device.index, planes, need)

        def fwd(a, b):
            out, taps = _forward_chain(model, a, b, planes, keep_all=False)
            return (out,), taps

        (out,), ctx.handle = graphs.forward(model, slot, _graph_key(model) if graphs.GRAPHS else None, (in0, in1), fwd,
                                            'train_lpips (lin only)', enabled=K is ops and in0.is_cuda)
        return out.view(n, 1, 1, 1)

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        n, shapes, need = ctx.n, ctx.shapes, ctx.need

        def bwd(taps, go):
            go = go.reshape(n, 1).float()
            grads = []
            for tap, shape, nd in zip(taps, shapes, need):
                if not nd:
                    grads.append(None)
                    continue
                f = tap.to_nchw()
                a, b = f[:n], f[n:]
                na = a.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10         # lpips.normalize_tensor
                nb = b.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10
                s = (a / na - b / nb).pow(2).mean(dim=(2, 3))                 # [n, C]
                grads.append((go * s).sum(dim=0).reshape(shape))
            return tuple(grads)

        grads = graphs.backward(ctx.handle, (go,), bwd, 'train_lpips (lin only)')
        return (None, None, None, None, *grads)


def distance_lin_only(model, in0, in1, planes=2):
    """[N, 1, 1, 1] distance of gradient-free images, differentiable w.r.t. the module's `lin` weights."""
    return _LpipsLinOnlyFn.apply(model, in0, in1, planes, *[model.lins[k].model[1].weight for k in range(5)])


def distance(model, in0, in1, planes=2):
    """[N, 1, 1, 1] LPIPS distance with the fused node; `model` is the lpips.LPIPS module (frozen VGG / lin weights)."""
    return _LpipsFn.apply(in0, in1, model, planes)
